// sm_100a device code, part 2: one ray against the scene — instance cull, HDDA over the tree levels,
// leaf step.  ONE copy of this code exists per kernel (the frame kernel calls it from a single site
// for primary, shadow and bounce rays alike): the first kernels inlined it three times and spent
// half of their issue slots waiting for instruction fetch (profiles/).
//
// Traversal semantics are DESIGN.md "Traversal semantics" (the CPU checker restates the same spec:
// trace_instance / trace_closest / trace_any); every float expression below has its twin there.
//
// Scheduling inside a warp (the part that is NOT in the oracle, and does not change results):
//   phase 1, cull   every lane walks the instance BVH and collects the instances whose padded box
//                   its ray touches, sorted by entry distance, in shared memory.
//   phase 2, walk   every lane then owns a little program: NEXT (enter its next candidate instance)
//                   -> WALK (leave a cell / descend the masks) ... -> LEAF (hit.rint's 4x4x4 DDA) -> ...
//                   Each trip of the loop the warp votes and executes the ONE operation most lanes are
//                   waiting for.  A lane that finishes an instance moves on to its next candidate by
//                   itself; nobody waits for the slowest lane of an instance, and the expensive leaf
//                   step always runs with as many lanes as are ready for it.
#pragma once
#include "trace_math.cuh"

namespace dust {

struct Counters {
    uint32_t primary_rays, bounce_rays, shadow_rays, mask_words, child_ptrs, hits_shaded;
    // scheduling diagnostics of the STATS instantiation (DUST_DEBUG_STATS=1 prints them; not part of the ABI):
    //  0..2 trips that ran WALK / LEAF / NEXT   3..5 lanes served by them   6 lanes idle (ray finished) summed over trips
    //  7 cells left with edge 1   8 edge 4   9 edge 64 (leaf units)   10 leaf steps   11 leaf DDA iterations
    //  12 instances entered   13 BVH nodes expanded   14 candidates   15 trace_scene calls (per warp)
    uint32_t dbg[16];
    uint32_t mask_words_strict;      // see DCounters
};

#define DUST_DIR_EPS 9.094947017729282e-13f   // 2^-40

#ifndef DUST_THREADS
#define DUST_THREADS 128                       // block size of every tracing kernel
#endif
#ifndef DUST_SCRATCH_WORDS
#define DUST_SCRATCH_WORDS 24                  // shared words per lane: candidates grow up, the BVH stack grows down
#endif
#define DUST_MAX_BVH_DEPTH (DUST_SCRATCH_WORDS - 4)
#ifndef DUST_WALK_REPS
#define DUST_WALK_REPS 2                       // cells a lane may walk per voted WALK operation (persistent tracer)
#endif
#ifndef DUST_FAST_VOTE_LANES_TS
#define DUST_FAST_VOTE_LANES_TS 16               // trace_scene: WALK runs without the full vote when at least this many lanes wait for it
#endif
#ifndef DUST_WALK_REPS_TS
#define DUST_WALK_REPS_TS 3                    // the same inside trace_scene (one thread per ray: coherent queues)
#endif

enum : int { M_NEXT = 0, M_WALK = 1, M_LEAF = 2, M_DONE = 3 };

struct Hit {
    bool hit;
    float t;
    uint32_t inst, leaf, pos, voxel;     // pos = leaf origin in leaf units: x | y << 10 | z << 20
};

// object-space ray of an instance (hit.rint:140-141) with the zero-direction guard (deviation (1))
__device__ __forceinline__ void object_ray(const DInstance &in, f3 ow, f3 dw, f3 &o, f3 &d)
{
    const float4 r0 = __ldg(reinterpret_cast<const float4 *>(in.w2o));
    const float4 r1 = __ldg(reinterpret_cast<const float4 *>(in.w2o) + 1);
    const float4 r2 = __ldg(reinterpret_cast<const float4 *>(in.w2o) + 2);
    o.x = ((r0.x * ow.x + r0.y * ow.y) + r0.z * ow.z) + r0.w;
    o.y = ((r1.x * ow.x + r1.y * ow.y) + r1.z * ow.z) + r1.w;
    o.z = ((r2.x * ow.x + r2.y * ow.y) + r2.z * ow.z) + r2.w;
    d.x = (r0.x * dw.x + r0.y * dw.y) + r0.z * dw.z;
    d.y = (r1.x * dw.x + r1.y * dw.y) + r1.z * dw.z;
    d.z = (r2.x * dw.x + r2.y * dw.y) + r2.z * dw.z;
    if (fabsf(d.x) < DUST_DIR_EPS) d.x = copysignf(DUST_DIR_EPS, d.x);
    if (fabsf(d.y) < DUST_DIR_EPS) d.y = copysignf(DUST_DIR_EPS, d.y);
    if (fabsf(d.z) < DUST_DIR_EPS) d.z = copysignf(DUST_DIR_EPS, d.z);
}

// conservative slab test of a (padded) world box; returns entry t or +inf
__device__ __forceinline__ float box_entry(float lx, float ly, float lz, float hx, float hy, float hz, f3 o, f3 inv, float tmin, float tmax)
{
    float ax = (lx - o.x) * inv.x, bx = (hx - o.x) * inv.x;
    float ay = (ly - o.y) * inv.y, by = (hy - o.y) * inv.y;
    float az = (lz - o.z) * inv.z, bz = (hz - o.z) * inv.z;
    float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), tmin));
    float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
    return t0 <= t1 ? t0 : CUDART_INF_F;
}

// 1/d for the box tests only: a (near-)zero component gets a huge FINITE inverse so (lo - o) * inv is never 0 * inf.  The bound is
// 1e24 (a component below 1e-24 moves the ray by less than the boxes' padding over any distance that matters): with it the wide
// node's decode, fma(2^23 + q, step * inv, ..), cannot overflow for grid steps below 2^24 world units.
__device__ __forceinline__ f3 cull_inverse(f3 rd)
{
    return mk3(fabsf(rd.x) < 1e-24f ? copysignf(1e24f, rd.x) : 1.0f / rd.x,
               fabsf(rd.y) < 1e-24f ? copysignf(1e24f, rd.y) : 1.0f / rd.y,
               fabsf(rd.z) < 1e-24f ? copysignf(1e24f, rd.z) : 1.0f / rd.z);
}

// ---------------------------------------------------------------------------------------- the cull
// Candidates = instances whose padded box the ray touches, kept sorted by a 32-bit key
// (entry t truncated downwards | instance) in this lane's shared-memory column sm[0 .. nc); the BVH
// stack grows down from the top of the same column, sm[sp .. WORDS).  The result of a ray does not
// depend on the visiting order (lexicographic min of (t, instance)), only the pruning does.

__device__ __forceinline__ uint32_t cand_key(float te, uint32_t instance, uint32_t id_mask)
{
    return (__float_as_uint(fmaxf(te, 0.0f)) & ~id_mask) | instance;
}

// start a ray: scene box test; returns false when the ray misses the whole scene
__device__ __forceinline__ bool cull_begin(const DSceneView &sc, f3 ro, f3 winv, float tmin, float tmax, uint32_t id_mask,
                                           uint32_t *sm, int &sp, int &pend, int &nc)
{
    sp = DUST_SCRATCH_WORDS; pend = -1; nc = 0;
    if (!sc.n_inst) return false;
    const float te = box_entry(sc.root_lo[0], sc.root_lo[1], sc.root_lo[2], sc.root_hi[0], sc.root_hi[1], sc.root_hi[2], ro, winv, tmin, tmax);
    if (te == CUDART_INF_F) return false;
    if (sc.root_ref >= 0) pend = sc.root_ref;
    else { sm[0] = cand_key(te, (uint32_t)~sc.root_ref, id_mask); nc = 1; }
    return true;
}

// Per-tile candidate lists (wf_tile_cull, kernels.cuh): the rays of a 32x32 tile share the camera origin, so ONE conservative frustum
// test per tile and instance replaces every ray's BVH walk (26 % of the primary pass' instructions, profiles/r02_history.md); a ray then
// runs the same exact box test on the few listed instances.  The candidate set is the one the BVH walk finds (box tests only cull).
#define DUST_TILE_LIST_MAX 15                  // entries per tile (+ 1 count word = 64 bytes); a longer list marks the tile "walk the BVH"
#define DUST_TILE_LIST_WORDS (DUST_TILE_LIST_MAX + 1)

__device__ __forceinline__ void cull_from_list(const DSceneView &sc, const uint32_t *list, f3 ro, f3 winv, float tmin, float tmax, uint32_t id_mask,
                                               uint32_t *sm, int &nc)
{
    const uint32_t n = __ldg(&list[0]);
    for (uint32_t e = 0; e < n; e++) {
        const uint32_t inst = __ldg(&list[1u + e]);
        const float4 lo = __ldg(&sc.inst_boxes[2u * inst]), hi = __ldg(&sc.inst_boxes[2u * inst + 1u]);
        const float te = box_entry(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, ro, winv, tmin, tmax);
        if (te == CUDART_INF_F) continue;
        const uint32_t key = cand_key(te, inst, id_mask);
        int k = nc++;
        while (k > 0) {
            const uint32_t prev = sm[(k - 1) * DUST_THREADS];
            if (prev <= key) break;
            sm[k * DUST_THREADS] = prev; --k;
        }
        sm[k * DUST_THREADS] = key;
    }
}

#ifdef DUST_BINARY_BVH
// room for one more expansion (2 candidates + 1 stack entry)?
__device__ __forceinline__ bool cull_room(int sp, int nc) { return nc + 3 <= sp; }

struct CullRay { f3 ro, winv; };
__device__ __forceinline__ CullRay cull_ray(const DSceneView &, f3 ro, f3 winv) { CullRay r; r.ro = ro; r.winv = winv; return r; }

// expand BVH node `pend` (both children: one 64-byte node); instances go to the candidate list, inner
// nodes to `pend` / the stack
template <bool STATS>
__device__ __forceinline__ void cull_expand(const DSceneView &sc, const CullRay &cr, float tmin, float tmax, uint32_t id_mask,
                                            uint32_t *sm, int &sp, int &pend, int &nc, Counters &st)
{
    const f3 ro = cr.ro, winv = cr.winv;
    const float4 *n = reinterpret_cast<const float4 *>(sc.bvh + pend);
    const float4 a = __ldg(n), b = __ldg(n + 1), c = __ldg(n + 2), e = __ldg(n + 3);
    if (STATS) st.dbg[13]++;
    float t0 = box_entry(a.x, a.y, a.z, b.x, b.y, b.z, ro, winv, tmin, tmax);
    float t1 = box_entry(c.x, c.y, c.z, e.x, e.y, e.z, ro, winv, tmin, tmax);
    int c0 = __float_as_int(a.w), c1 = __float_as_int(b.w);
    // near child first: the candidate set does not depend on the visiting order, but candidates then arrive almost
    // sorted and the insertion below rarely moves anything (without the swap wf_cull was 8 % slower)
    if (t1 < t0) { const float tf = t0; t0 = t1; t1 = tf; const int ci = c0; c0 = c1; c1 = ci; }
    int next = -1;
#pragma unroll
    for (int s = 0; s < 2; s++) {
        const float te = s ? t1 : t0;
        const int ch = s ? c1 : c0;
        if (te == CUDART_INF_F) continue;
        if (ch >= 0) {
            if (next < 0) next = ch; else sm[--sp * DUST_THREADS] = (uint32_t)ch;
        } else {                                                      // instance: insertion sort by key
            const uint32_t key = cand_key(te, (uint32_t)~ch, id_mask);
            int k = nc++;
            if (STATS) st.dbg[14]++;
            while (k > 0) {
                const uint32_t prev = sm[(k - 1) * DUST_THREADS];
                if (prev <= key) break;
                sm[k * DUST_THREADS] = prev; --k;
            }
            sm[k * DUST_THREADS] = key;
        }
    }
    if (next < 0 && sp < DUST_SCRATCH_WORDS) next = (int)sm[sp++ * DUST_THREADS];
    pend = next;
}
#else
// room for one more expansion (up to DUST_BVH_WIDTH candidates + 1 stack entry)?
__device__ __forceinline__ bool cull_room(int sp, int nc) { return nc + DUST_BVH_WIDTH + 1 <= sp; }

// what a ray needs to test quantised boxes (device_layout.h "WIDE BVH"), computed once per ray:
// t(q) = fma(2^23 + q, s, a) with s = step / d (exact scaling by a power of two) and a = (org - o) / d - 2^23 s
struct CullRay {
    float sx, sy, sz, ax, ay, az;
    uint32_t nx, ny, nz;             // byte-permute selectors of the NEAR plane: {q.b0, q.b1, 0x00, 0x4B} = float 2^23 + lo for d > 0, .. + hi otherwise; far = near ^ 0x22
};

__device__ __forceinline__ CullRay cull_ray(const DSceneView &sc, f3 ro, f3 winv)
{
    CullRay r;
    r.sx = sc.grid_step[0] * winv.x; r.sy = sc.grid_step[1] * winv.y; r.sz = sc.grid_step[2] * winv.z;
    r.ax = __fmaf_rn(-8388608.0f, r.sx, (sc.grid_org[0] - ro.x) * winv.x);
    r.ay = __fmaf_rn(-8388608.0f, r.sy, (sc.grid_org[1] - ro.y) * winv.y);
    r.az = __fmaf_rn(-8388608.0f, r.sz, (sc.grid_org[2] - ro.z) * winv.z);
    r.nx = winv.x > 0.0f ? 0x7410u : 0x7432u; r.ny = winv.y > 0.0f ? 0x7410u : 0x7432u; r.nz = winv.z > 0.0f ? 0x7410u : 0x7432u;
    return r;
}

__device__ __forceinline__ float cull_plane(uint32_t q, uint32_t sel, float s, float a)
{
    return __fmaf_rn(__uint_as_float(__byte_perm(q, 0x4B000000u, sel)), s, a);
}

// expand wide node `pend`: instances go to the candidate list, the first inner child that is hit becomes `pend`, the others
// are ONE stack entry (node << 8 | mask of children still to visit), so the stack is as deep as the tree (the host keeps
// that depth within the scratch column: dust_abi.cu build_wide).  The quantised boxes contain the padded instance boxes, so an
// entry time from here is a lower bound like box_entry's.
template <bool STATS>
__device__ __forceinline__ void cull_expand(const DSceneView &sc, const CullRay &cr, float tmin, float tmax, uint32_t id_mask,
                                            uint32_t *sm, int &sp, int &pend, int &nc, Counters &st)
{
    const uint4 *n = reinterpret_cast<const uint4 *>(sc.bvh) + 4u * (size_t)pend;
    const uint4 X = __ldg(n), Y = __ldg(n + 1), Z = __ldg(n + 2), R = __ldg(n + 3);
    if (STATS) st.dbg[13]++;
    const uint32_t fx = cr.nx ^ 0x22u, fy = cr.ny ^ 0x22u, fz = cr.nz ^ 0x22u;
    int next = -1;
    uint32_t rem = 0;
    // (the four tests first without a branch and ONE insertion loop over what they found was tried: wf_cull 0.66 ms against 0.56)
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const uint32_t qx = c == 0 ? X.x : (c == 1 ? X.y : (c == 2 ? X.z : X.w));
        const uint32_t qy = c == 0 ? Y.x : (c == 1 ? Y.y : (c == 2 ? Y.z : Y.w));
        const uint32_t qz = c == 0 ? Z.x : (c == 1 ? Z.y : (c == 2 ? Z.z : Z.w));
        const int ch = (int)(c == 0 ? R.x : (c == 1 ? R.y : (c == 2 ? R.z : R.w)));
        const float t0 = fmaxf(fmaxf(cull_plane(qx, cr.nx, cr.sx, cr.ax), cull_plane(qy, cr.ny, cr.sy, cr.ay)),
                               fmaxf(cull_plane(qz, cr.nz, cr.sz, cr.az), tmin));
        const float t1 = fminf(fminf(cull_plane(qx, fx, cr.sx, cr.ax), cull_plane(qy, fy, cr.sy, cr.ay)),
                               fminf(cull_plane(qz, fz, cr.sz, cr.az), tmax));
        if (!(t0 <= t1)) continue;
        if (ch >= 0) {
            if (next < 0) next = ch; else rem |= 1u << c;
        } else {                                                      // instance: insertion sort by key
            const uint32_t key = cand_key(t0, (uint32_t)~ch, id_mask);
            int k = nc++;
            if (STATS) st.dbg[14]++;
            while (k > 0) {
                const uint32_t prev = sm[(k - 1) * DUST_THREADS];
                if (prev <= key) break;
                sm[k * DUST_THREADS] = prev; --k;
            }
            sm[k * DUST_THREADS] = key;
        }
    }
    if (rem) sm[--sp * DUST_THREADS] = ((uint32_t)pend << 8) | rem;
    if (next < 0 && sp < DUST_SCRATCH_WORDS) {                          // pop: the lowest child still to visit of the top entry
        const uint32_t e = sm[sp * DUST_THREADS];
        const uint32_t c = (uint32_t)__ffs((int)(e & 0xFFu)) - 1u;
        const uint32_t left = e & (e - 1u);                              // the low byte is non-zero: clears that child's bit
        if (left & 0xFFu) sm[sp * DUST_THREADS] = left; else sp++;
        next = __ldg(reinterpret_cast<const int *>(reinterpret_cast<const uint4 *>(sc.bvh) + 4u * (size_t)(e >> 8) + 3u) + c);
    }
    pend = next;
}
#endif

// ------------------------------------------------------------------------------- the instance walk
// State of one ray inside one instance.  Coordinates are in LEAF units (voxel >> 2): no consumer reads
// the low two bits (floor / clamp commute with the monotone shift).

struct Walk {
    f3 o, d, inv;                    // object-space ray, exact 1/d
    int jx, jy, jz;                  // current cell
    int L4;                          // != 0: the cell of that edge (leaf units) containing (jx,jy,jz) is empty / done: leave it first
    float t_exit, t_box;             // inclusive end of the walk (box exit or padded tmax); exit time of the occupied bounds
    const DNode16 *n16;
#ifdef DUST_DENSE_MIDS
    const uint4 *dense;              // DInstance::dense_root
#else
    const uint2 *blocks;
#endif
#ifdef DUST_TMA_TOP
    const uint4 *topw;               // 64 words of the root 16^3 node: global, or the CTA's TMA-staged shared copy (generic loads)
    const uint2 *topb;               // its block mask (4-level trees)
#endif
    const uint4 *mids, *leaves;
    uint32_t blk_lo, blk_hi;         // block mask of the ROOT node (of the 16^3 node above the mids only for 3-level trees)
    uint32_t mid_lo, mid_hi, mid_base;
    uint32_t inst, leaf_i;
    bool have_mid;                   // the 4^3 node of the current 16-cell is in registers
    bool four;                       // this instance has 4 levels
};

enum : int { WS_WALK = 0, WS_LEAF = 1, WS_OUT = 2 };

// Enter instance `inst`: world -> object ray (hit.rint:140-141), reciprocals, clip against the occupied
// bounds.  Returns false when the ray does not reach the instance within [tmin, cur].
template <int NL, bool STATS>
__device__ __forceinline__ bool walk_enter(const DSceneView &sc, Walk &w, uint32_t inst, f3 ro, f3 rd, float tmin, float cur, Counters &st)
{
    w.inst = inst;
    if (STATS) st.dbg[12]++;
    const DInstance &in = sc.inst[inst];
    object_ray(in, ro, rd, w.o, w.d);
    w.inv = mk3(1.0f / w.d.x, 1.0f / w.d.y, 1.0f / w.d.z);
    const float4 blo = __ldg(reinterpret_cast<const float4 *>(in.bounds_lo)), bhi = __ldg(reinterpret_cast<const float4 *>(in.bounds_hi));
    // plane times are (plane - o) * inv: difference first, plane*inv - o*inv would cancel (|o/d| >> t)
    const float t0x = (blo.x - w.o.x) * w.inv.x, t0y = (blo.y - w.o.y) * w.inv.y, t0z = (blo.z - w.o.z) * w.inv.z;
    const float t1x = (bhi.x - w.o.x) * w.inv.x, t1y = (bhi.y - w.o.y) * w.inv.y, t1z = (bhi.z - w.o.z) * w.inv.z;
    const float t_enter = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), tmin));
    // walk a hair past tmax (acceptance stays exact): DESIGN.md "Traversal semantics"
    const float tmax_pad = cur + (fabsf(cur) * 3.814697265625e-06f + 9.5367431640625e-07f);
    w.t_box = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
    w.t_exit = fminf(w.t_box, tmax_pad);
    if (!(t_enter <= w.t_exit) || bhi.x <= blo.x) return false;
    // entry cell = floor(fma(d, t_enter, o)) clamped into the bounds: clamping the float to [lo, hi - 1] before the floor gives the
    // same integer (lo, hi are integers below 2^24; floor is monotone), and a NaN goes to lo either way
    const int ix = __float2int_rd(fminf(fmaxf(__fmaf_rn(w.d.x, t_enter, w.o.x), blo.x), bhi.x - 1.0f));
    const int iy = __float2int_rd(fminf(fmaxf(__fmaf_rn(w.d.y, t_enter, w.o.y), blo.y), bhi.y - 1.0f));
    const int iz = __float2int_rd(fminf(fmaxf(__fmaf_rn(w.d.z, t_enter, w.o.z), blo.z), bhi.z - 1.0f));
    w.jx = ix >> 2; w.jy = iy >> 2; w.jz = iz >> 2;
    const ulonglong2 p01 = __ldg(reinterpret_cast<const ulonglong2 *>(&in.n16));
    w.n16 = reinterpret_cast<const DNode16 *>(p01.x);
    w.mids = reinterpret_cast<const uint4 *>(p01.y);
    w.leaves = in.leaves;
    if (NL == 0) w.four = __float_as_uint(blo.w) == 4u;
#ifdef DUST_DENSE_MIDS
    // the root level is walked through the dense table (DInstance::dense_root) and the root's block mask comes with the instance
    {
        const ulonglong2 dr = __ldg(reinterpret_cast<const ulonglong2 *>(&in.dense_root));
        w.dense = reinterpret_cast<const uint4 *>(dr.x);
        w.blk_lo = (uint32_t)dr.y; w.blk_hi = (uint32_t)(dr.y >> 32);     // (only 4-level trees ask it: the table answers for 3-level ones)
    }
#else
    w.blocks = in.blocks;
#ifdef DUST_TMA_TOP
    w.topw = w.n16->word; w.topb = w.blocks;
#endif
    if (NL == 3 || (NL == 0 && !w.four)) { const uint2 b = __ldg(w.blocks); w.blk_lo = b.x; w.blk_hi = b.y; }
#endif
    w.have_mid = false;
    w.L4 = 0;
    return true;
}

// One step of the walk: leave the cell marked by L4 (if any), then descend to the cell that now
// contains (jx,jy,jz) with the InternalNode::get index arithmetic (crates/vdb/src/node/internal.rs:77-95)
// on the re-laid-out nodes.  Every 16^3 node is first asked per 4x4x4 BLOCK of children ("is anything
// in this block": one bit of a 64-bit block mask) — an empty block is left as one cell of 4x the edge.
template <int NL, bool STATS>
__device__ __forceinline__ int walk_step(Walk &w, Counters &st)
{
    if (w.L4) {
        const int L4 = w.L4;
        if (STATS) st.dbg[L4 == 1 ? 7 : (L4 == 4 ? 8 : 9)]++;
        // leave the cell of edge L4 that contains (jx,jy,jz): DESIGN.md "Traversal semantics", "t_next"
        const int cx = w.jx & ~(L4 - 1), cy = w.jy & ~(L4 - 1), cz = w.jz & ~(L4 - 1);
        const bool px_ = w.d.x > 0.0f, py_ = w.d.y > 0.0f, pz_ = w.d.z > 0.0f;
        const float tx = ((float)((px_ ? cx + L4 : cx) << 2) - w.o.x) * w.inv.x;
        const float ty = ((float)((py_ ? cy + L4 : cy) << 2) - w.o.y) * w.inv.y;
        const float tz = ((float)((pz_ ? cz + L4 : cz) << 2) - w.o.z) * w.inv.z;
        const float t_next = fminf(fminf(tx, ty), tz);
        // inclusive end; a stepping axis that leaves the occupied bounds has t_next >= t_box (same float
        // expressions as the bounds planes, see the CPU checker): ONE comparison, no per-axis range checks
        if (!(t_next <= w.t_exit) || !(t_next < w.t_box)) { w.L4 = 0; return WS_OUT; }
        int nx, ny, nz;
        // axes whose plane time ties the minimum step one cell; the others are re-derived from the ray at
        // t_next and clamped into the cell's own range — a no-op at leaf granularity when L4 == 1
        {   // select form — both values for every axis, no divergent regions: +1.6 % (config 2) / +3.4 % (config 5) over
            // `if (t_a <= t_next) step else if (L4 > 1) re-derive` per axis (profiles/r01_history.md)
            const int hx_ = cx + L4 - 1, hy_ = cy + L4 - 1, hz_ = cz + L4 - 1;
            const int rx = min(max(__float2int_rd(__fmaf_rn(w.d.x, t_next, w.o.x)) >> 2, cx), hx_);
            const int ry = min(max(__float2int_rd(__fmaf_rn(w.d.y, t_next, w.o.y)) >> 2, cy), hy_);
            const int rz = min(max(__float2int_rd(__fmaf_rn(w.d.z, t_next, w.o.z)) >> 2, cz), hz_);
            // (for L4 == 1 the clamp range is the single value c = j: the re-derived coordinate is the old one)
            nx = tx <= t_next ? (px_ ? cx + L4 : cx - 1) : rx;
            ny = ty <= t_next ? (py_ ? cy + L4 : cy - 1) : ry;
            nz = tz <= t_next ? (pz_ ? cz + L4 : cz - 1) : rz;
        }
        // the cached 4^3 node survives a step that stays inside its 16-cell
        // (with the dense root table a 4^3 node word is one L1 hit away; NOT keeping it in registers measured the same: 2 884 against 2 880 Mrays/s)
        w.have_mid = w.have_mid && (((uint32_t)((nx ^ w.jx) | (ny ^ w.jy) | (nz ^ w.jz)) >> 2) == 0u);
        w.jx = nx; w.jy = ny; w.jz = nz;
        w.L4 = 0;
    }
    const int jx = w.jx, jy = w.jy, jz = w.jz;
    const bool four = NL == 4 || (NL == 0 && w.four);
    if (!w.have_mid) {
        const DNode16 *node = w.n16;
        uint32_t blo = w.blk_lo, bhi = w.blk_hi;
        if (four) {
            // (keeping the 16^3 node of the current 256-cell in registers, like the 4^3 node below, was tried:
            //  these words are L1 hits and the extra state cost more than the three loads — config 5 lost 9 %)
#ifdef DUST_DENSE_MIDS
            if (STATS) st.mask_words++;
            if (!bit64(w.blk_lo, w.blk_hi, ((uint32_t)((jx >> 8) & 3) << 4) | ((uint32_t)((jy >> 8) & 3) << 2) | (uint32_t)((jz >> 8) & 3))) { w.L4 = 256; return WS_WALK; }
            const uint32_t tidx = ((uint32_t)(jx >> 6) << 8) | ((uint32_t)(jy >> 6) << 4) | (uint32_t)(jz >> 6);
            // the root cell's entry of the dense table: the child 16^3 node's block mask and index, zero = no child.  Counted like
            // the compact look-up it replaces (root word, child link).
            const uint4 e = __ldg(&w.dense[tidx]);
            if (STATS) { st.mask_words++; st.mask_words_strict++; }
            if ((e.x | e.y) == 0u) { w.L4 = 64; return WS_WALK; }
            if (STATS) st.child_ptrs++;
            node = w.n16 + e.z;
            blo = e.x; bhi = e.y;
#else
#ifdef DUST_TMA_TOP
            const uint2 tb = *w.topb;
#else
            const uint2 tb = __ldg(w.blocks);
#endif
            if (STATS) st.mask_words++;
            if (!bit64(tb.x, tb.y, ((uint32_t)((jx >> 8) & 3) << 4) | ((uint32_t)((jy >> 8) & 3) << 2) | (uint32_t)((jz >> 8) & 3))) { w.L4 = 256; return WS_WALK; }
            const uint32_t tidx = ((uint32_t)(jx >> 6) << 8) | ((uint32_t)(jy >> 6) << 4) | (uint32_t)(jz >> 6);
#ifdef DUST_TMA_TOP
            const uint4 q = w.topw[tidx >> 6];
#else
            const uint4 q = __ldg(&node->word[tidx >> 6]);
#endif
            if (STATS) { st.mask_words++; st.mask_words_strict++; }
            const uint32_t b = tidx & 63u;
            if (!bit64(q.x, q.y, b)) { w.L4 = 64; return WS_WALK; }
            if (STATS) st.child_ptrs++;
            const uint32_t ni = q.z + rank64(q.x, q.y, b);
            node = w.n16 + ni;
            const uint2 cb = __ldg(w.blocks + ni);
            blo = cb.x; bhi = cb.y;
#endif
        }
        uint4 mid;
#ifdef DUST_DENSE_MIDS
        if (!four) {
            // the cell's entry of the dense root table: the 4^3 node's word; zero = no child (a 4^3 node that exists has a leaf), and
            // then w says whether the whole 4x4x4 block of cells is empty (what the root's block mask would have said first).
            // Counted like the compact look-ups it replaces (the algorithmic words are a property of the walk, not of the table).
            const uint32_t ridx = ((uint32_t)((jx >> 2) & 15) << 8) | ((uint32_t)((jy >> 2) & 15) << 4) | (uint32_t)((jz >> 2) & 15);
            mid = __ldg(&w.dense[ridx]);
            if (STATS) st.mask_words++;
            if ((mid.x | mid.y) == 0u) {
                if (STATS && !mid.w) { st.mask_words++; st.mask_words_strict++; }
                w.L4 = mid.w ? 16 : 4;
                return WS_WALK;
            }
            if (STATS) { st.mask_words++; st.mask_words_strict++; }
        } else
#endif
        {
            if (STATS) st.mask_words++;
            if (!bit64(blo, bhi, ((uint32_t)((jx >> 4) & 3) << 4) | ((uint32_t)((jy >> 4) & 3) << 2) | (uint32_t)((jz >> 4) & 3))) { w.L4 = 16; return WS_WALK; }
            const uint32_t ridx = ((uint32_t)((jx >> 2) & 15) << 8) | ((uint32_t)((jy >> 2) & 15) << 4) | (uint32_t)((jz >> 2) & 15);
            if (STATS) { st.mask_words++; st.mask_words_strict++; }
#ifdef DUST_TMA_TOP
            const uint4 q = four ? __ldg(&node->word[ridx >> 6]) : w.topw[ridx >> 6];
#else
            const uint4 q = __ldg(&node->word[ridx >> 6]);
#endif
            const uint32_t b = ridx & 63u;
            if (!bit64(q.x, q.y, b)) { w.L4 = 4; return WS_WALK; }
            mid = __ldg(&w.mids[q.z + rank64(q.x, q.y, b)]);
        }
        if (STATS) st.mask_words_strict++;                              // a mid node entered
        w.mid_lo = mid.x; w.mid_hi = mid.y; w.mid_base = mid.z;
        w.have_mid = true;
    } else if (STATS) {
        if (four) { st.mask_words += 2; st.child_ptrs++; }
        st.mask_words += 2;
    }
    if (STATS) { st.child_ptrs++; st.mask_words++; }
    const uint32_t midx = ((uint32_t)(jx & 3) << 4) | ((uint32_t)(jy & 3) << 2) | (uint32_t)(jz & 3);
    if (!bit64(w.mid_lo, w.mid_hi, midx)) { w.L4 = 1; return WS_WALK; }
    if (STATS) { st.child_ptrs++; st.mask_words++; }
    w.leaf_i = w.mid_base + rank64(w.mid_lo, w.mid_hi, midx);
    // (a prefetch.global.L1 of the leaf word here, a vote before the LEAF operation reads it: 2 988 against 3 003 Mrays/s — not kept)
    return WS_LEAF;
}

// Leaf step on leaf w.leaf_i: one 128-bit load (occupancy + material_ptr), then hit.rint's DDA.  Returns
// true with (t, voxel) when a voxel is reached; the caller applies reportIntersectionEXT's [tmin, tmax].
// (A resumable DDA — at most K iterations per voted LEAF operation — was tried: more trips, slower.)
// leaf_mode (uniform per launch): 0 = hit.rint (precise DDA), 1 = rough.rint (solid block), 2 = ambient_occlusion.rint
enum : int { LEAF_PRECISE = 0, LEAF_ROUGH = 1, LEAF_AO = 2 };
template <bool STATS>
__device__ __forceinline__ bool walk_leaf(const Walk &w, float tmin, float &th, uint32_t &vox, Counters &st, const int leaf_mode = LEAF_PRECISE)
{
    const uint4 leaf = __ldg(&w.leaves[w.leaf_i]);
    if (STATS) { st.dbg[10]++; st.mask_words_strict++; }
    const f3 lo = mk3(w.o.x - (float)(w.jx << 2), w.o.y - (float)(w.jy << 2), w.o.z - (float)(w.jz << 2));   // hit.rint:140
    if (leaf_mode == LEAF_ROUGH) return leaf_rough(lo, w.d, leaf.x, leaf.y, th, vox);     // block-granular far rays (rough.rint)
    if (leaf_mode == LEAF_AO) return leaf_dda_inv<STATS, true>(lo, w.d, w.inv, leaf.x, leaf.y, tmin, th, vox, st.dbg[11]);
    return leaf_dda_inv<STATS>(lo, w.d, w.inv, leaf.x, leaf.y, tmin, th, vox, st.dbg[11]);
}

__device__ __forceinline__ uint32_t walk_pos(const Walk &w) { return (uint32_t)w.jx | ((uint32_t)w.jy << 10) | ((uint32_t)w.jz << 20); }

// ----------------------------------------------------------------------------------- one ray, one warp
// Closest hit = lexicographic minimum of (t, instance) over all instances; `any`: the first accepted
// hit ends the ray (gl_RayFlagsTerminateOnFirstHitEXT).  The BVH only culls (boxes are padded, see
// dust_scene_create), so the result equals the CPU checker's linear instance loop.
// All 32 lanes of the warp must call this together; lanes without a ray pass live = false.
// NL = 3 / 4: every instance has that many levels; NL = 0: mixed scene, decided per instance.
// `sm` = this lane's column of the block's scratch: word k lives at sm[k * DUST_THREADS].
template <int NL, bool STATS>
__device__ __forceinline__ void trace_scene(const DSceneView &sc, const f3 ro, const f3 rd, const float tmin, const float tmax,
                                            const bool any, const bool live, uint32_t *sm, Hit &best, Counters &st, const int leaf_mode = LEAF_PRECISE,
                                            const uint32_t *tile_list = nullptr)
{
    const unsigned full = 0xffffffffu;
    best.hit = false; best.inst = 0xFFFFFFFFu; best.t = tmax;
    float cur = tmax;                                                 // current tmax: shrinks to the best hit
    const uint32_t id_mask = (1u << sc.id_bits) - 1u;
    if (STATS && (threadIdx.x & 31u) == 0) st.dbg[15]++;
    const f3 winv = cull_inverse(rd);
    const CullRay cr = cull_ray(sc, ro, winv);
    int sp = DUST_SCRATCH_WORDS, pend = -1, nc = 0;
    if (live) {
        // warp-uniform: a tile list (camera rays of one tile) replaces the BVH walk unless the tile overflowed its list
        if (tile_list && __ldg(&tile_list[0]) != 0xFFFFFFFFu) cull_from_list(sc, tile_list, ro, winv, tmin, cur, id_mask, sm, nc);
        else cull_begin(sc, ro, winv, tmin, cur, id_mask, sm, sp, pend, nc);
    }

    Walk w;
    w.four = NL == 4; w.have_mid = false; w.L4 = 0; w.blk_lo = w.blk_hi = 0;
    for (;;) {
        // ---- phase 1: cull
        while (pend >= 0 && cull_room(sp, nc)) cull_expand<STATS>(sc, cr, tmin, cur, id_mask, sm, sp, pend, nc, st);

        // ---- phase 2: walk.  NEXT: enter the next candidate; WALK: one cell; LEAF: the 4x4x4 DDA.  Each trip
        // the warp votes and executes the ONE operation most lanes wait for (ties to the cheaper one).
        int mode = nc > 0 ? M_NEXT : M_DONE, j = 0;
        for (;;) {
            int op = M_WALK, cnt = 0, n_next = 0, n_walk = 0, n_leaf = 0;
#ifndef DUST_NO_FAST_VOTE_TS
            // fast path: half of the warp or more waits for WALK — the plurality whatever the others want (one ballot + popc)
            n_walk = __popc(__ballot_sync(full, mode == M_WALK));
            cnt = n_walk;
            if (n_walk < DUST_FAST_VOTE_LANES_TS)
#endif
            {
                // one REDUX of three packed 8-bit counters instead of three ballots + popcounts
                const uint32_t packed = __reduce_add_sync(full, mode < M_DONE ? (1u << (8 * mode)) : 0u);
                if (packed == 0u) break;
                n_next = (int)(packed & 255u); n_walk = (int)((packed >> 8) & 255u); n_leaf = (int)(packed >> 16);
                cnt = n_walk;
                if (n_next > cnt) { op = M_NEXT; cnt = n_next; }
                if (n_leaf > cnt) { op = M_LEAF; cnt = n_leaf; }
            }
            if (STATS && (threadIdx.x & 31u) == 0) {
                st.dbg[op == M_WALK ? 0 : (op == M_LEAF ? 1 : 2)]++;
                st.dbg[op == M_WALK ? 3 : (op == M_LEAF ? 4 : 5)] += cnt;
            }
            if (op == M_WALK) {
#pragma unroll 1
                for (int rep = 0; rep < DUST_WALK_REPS_TS; rep++)             // several cells per vote: WALK -> WALK is the common transition
                    if (mode == M_WALK) {
                        const int r = walk_step<NL, STATS>(w, st);
                        mode = r == WS_LEAF ? M_LEAF : (r == WS_OUT ? M_NEXT : M_WALK);
                    }
            } else if (op == M_LEAF) {
                if (mode == M_LEAF) {
                    float th; uint32_t vox;
                    if (walk_leaf<STATS>(w, tmin, th, vox, st, leaf_mode) && tmin <= th && th <= cur) {       // reportIntersectionEXT
                        if (any) { best.hit = true; best.inst = w.inst; mode = M_DONE; pend = -1; sp = DUST_SCRATCH_WORDS; }
                        else {
                            if (!best.hit || th < cur || w.inst < best.inst) {                     // (t, instance) lexicographic
                                best.hit = true; best.t = th; cur = th; best.inst = w.inst; best.leaf = w.leaf_i;
                                best.pos = walk_pos(w); best.voxel = vox;
                            }
                            mode = M_NEXT;                                                          // first accepted hit is this instance's closest
                        }
                    } else { mode = M_WALK; w.L4 = 1; }
                }
            } else {
                if (mode == M_NEXT) {
                    mode = M_DONE;
                    if (j < nc) {
                        const uint32_t key = sm[j * DUST_THREADS];
                        ++j;
                        if (!(__uint_as_float(key & ~id_mask) > cur))                               // sorted: nothing further can beat the best hit
                            mode = walk_enter<NL, STATS>(sc, w, key & id_mask, ro, rd, tmin, cur, st) ? M_WALK : M_NEXT;
                    }
                }
            }
        }
        if (!__any_sync(full, pend >= 0)) break;                      // candidate list overflowed: resume the cull
        nc = 0;
    }
}

}  // namespace dust
