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
};

#define DUST_DIR_EPS 9.094947017729282e-13f   // 2^-40

#ifndef DUST_THREADS
#define DUST_THREADS 128                       // block size of every tracing kernel
#endif
#ifndef DUST_SCRATCH_WORDS
#define DUST_SCRATCH_WORDS 24                  // shared words per lane: candidates grow up, the BVH stack grows down
#endif
#define DUST_MAX_BVH_DEPTH (DUST_SCRATCH_WORDS - 4)

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

// Closest hit = lexicographic minimum of (t, instance) over all instances; `any`: the first accepted
// hit ends the ray (gl_RayFlagsTerminateOnFirstHitEXT).  The BVH only culls (boxes are padded, see
// dust_scene_create), so the result equals the oracle's linear instance loop.
// All 32 lanes of the warp must call this together; lanes without a ray pass live = false.
// NL = 3 / 4: every instance has that many levels; NL = 0: mixed scene, decided per instance.
// `sm` = this lane's column of the block's scratch: word k lives at sm[k * DUST_THREADS].
template <int NL, bool STATS>
__device__ __forceinline__ void trace_scene(const DSceneView &sc, const f3 ro, const f3 rd, const float tmin, const float tmax,
                                            const bool any, const bool live, uint32_t *sm, Hit &best, Counters &st)
{
    const unsigned full = 0xffffffffu;
    best.hit = false; best.inst = 0xFFFFFFFFu; best.t = tmax;
    float cur = tmax;                                                 // current tmax: shrinks to the best hit
    const uint32_t id_mask = (1u << sc.id_bits) - 1u;
    if (STATS && (threadIdx.x & 31u) == 0) st.dbg[15]++;

    // ---- ray set-up for the cull: a zero component gets a huge FINITE inverse so (lo - o) * inv is never 0 * inf
    const f3 winv = mk3(fabsf(rd.x) < 1e-30f ? copysignf(1e30f, rd.x) : 1.0f / rd.x,
                        fabsf(rd.y) < 1e-30f ? copysignf(1e30f, rd.y) : 1.0f / rd.y,
                        fabsf(rd.z) < 1e-30f ? copysignf(1e30f, rd.z) : 1.0f / rd.z);
    int sp = DUST_SCRATCH_WORDS;                                      // stack = sm[sp .. WORDS)
    int pend = -1;                                                    // inner BVH node waiting to be expanded
    int nc = 0;                                                       // candidates = sm[0 .. nc)
    if (live && sc.n_inst) {
        const float te = box_entry(sc.root_lo[0], sc.root_lo[1], sc.root_lo[2], sc.root_hi[0], sc.root_hi[1], sc.root_hi[2], ro, winv, tmin, cur);
        if (te != CUDART_INF_F) {
            if (sc.root_ref >= 0) pend = sc.root_ref;
            else { sm[0] = (__float_as_uint(fmaxf(te, 0.0f)) & ~id_mask) | (uint32_t)~sc.root_ref; nc = 1; }
        }
    }

    // per-instance walk state
    int mode, j;
    uint32_t inst = 0, leaf_i = 0;
    f3 o = mk3(0.0f, 0.0f, 0.0f), d = mk3(1.0f, 1.0f, 1.0f), inv = mk3(1.0f, 1.0f, 1.0f);
    int jx = 0, jy = 0, jz = 0;                                       // current cell, leaf units (voxel >> 2)
    int L4 = 0;                                                       // != 0: leave the cell of that edge (leaf units) first
    float t_exit = 0.0f;
    uint32_t bnd0 = 0, bnd1 = 0, bnd2 = 0;
    const DNode16 *n16 = nullptr;
    const uint4 *mids = nullptr, *leaves = nullptr;
    uint32_t mid_lo = 0, mid_hi = 0, mid_base = 0;
    bool have_mid = false;                                            // the 4^3 node of the current 16-cell is in registers
    bool four = NL == 4;

    for (;;) {
        // ------------------------------------------------------------------ phase 1: cull
        // Collect the instances whose padded box the ray touches, sorted by (entry t, instance).  The
        // result does not depend on the visiting order (lexicographic min of (t, instance)), only the pruning does.
        while (pend >= 0 && nc + 3 <= sp) {
            const float4 *n = reinterpret_cast<const float4 *>(sc.bvh + pend);
            const float4 a = __ldg(n), b = __ldg(n + 1), c = __ldg(n + 2), e = __ldg(n + 3);
            if (STATS) st.dbg[13]++;
            float t0 = box_entry(a.x, a.y, a.z, b.x, b.y, b.z, ro, winv, tmin, cur);
            float t1 = box_entry(c.x, c.y, c.z, e.x, e.y, e.z, ro, winv, tmin, cur);
            int c0 = __float_as_int(a.w), c1 = __float_as_int(b.w);
            if (t1 < t0) { const float tf = t0; t0 = t1; t1 = tf; const int ci = c0; c0 = c1; c1 = ci; }
            int next = -1;
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const float te = s ? t1 : t0;
                const int ch = s ? c1 : c0;
                if (te == CUDART_INF_F) continue;
                if (ch >= 0) {
                    if (next < 0) next = ch; else sm[--sp * DUST_THREADS] = (uint32_t)ch;
                } else {                                              // instance: insertion sort by key
                    const uint32_t key = (__float_as_uint(fmaxf(te, 0.0f)) & ~id_mask) | (uint32_t)~ch;
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

        // ------------------------------------------------------------------ phase 2: walk
        mode = nc > 0 ? M_NEXT : M_DONE;
        j = 0;
        for (;;) {
            const unsigned m_next = __ballot_sync(full, mode == M_NEXT);
            const unsigned m_walk = __ballot_sync(full, mode == M_WALK);
            const unsigned m_leaf = __ballot_sync(full, mode == M_LEAF);
            if ((m_next | m_walk | m_leaf) == 0) break;
            // the operation most lanes wait for; ties go to the cheaper one (it feeds the dearer one)
            const int n_next = __popc(m_next), n_walk = __popc(m_walk), n_leaf = __popc(m_leaf);
            int op = M_WALK, cnt = n_walk;
            if (n_next > cnt) { op = M_NEXT; cnt = n_next; }
            if (n_leaf > cnt) { op = M_LEAF; }
            if (STATS && (threadIdx.x & 31u) == 0) {
                st.dbg[op == M_WALK ? 0 : (op == M_LEAF ? 1 : 2)]++;
                st.dbg[op == M_WALK ? 3 : (op == M_LEAF ? 4 : 5)] += op == M_WALK ? n_walk : (op == M_LEAF ? n_leaf : n_next);
                st.dbg[6] += 32 - n_walk - n_leaf - n_next;
            }

            if (op == M_WALK) {
                if (mode == M_WALK) {
                    if (L4) {
                        if (STATS) st.dbg[L4 == 1 ? 7 : (L4 == 4 ? 8 : 9)]++;
                        // leave the cell of edge L4 that contains (jx,jy,jz): oracle trace_instance, "t_next"
                        const int cx = jx & ~(L4 - 1), cy = jy & ~(L4 - 1), cz = jz & ~(L4 - 1);
                        const bool px_ = d.x > 0.0f, py_ = d.y > 0.0f, pz_ = d.z > 0.0f;
                        const float tx = ((float)((px_ ? cx + L4 : cx) << 2) - o.x) * inv.x;
                        const float ty = ((float)((py_ ? cy + L4 : cy) << 2) - o.y) * inv.y;
                        const float tz = ((float)((pz_ ? cz + L4 : cz) << 2) - o.z) * inv.z;
                        const float t_next = fminf(fminf(tx, ty), tz);
                        bool out = !(t_next <= t_exit);               // inclusive, DESIGN.md "Traversal semantics"
                        int nx = jx, ny = jy, nz = jz;
                        // axes whose plane time ties the minimum step one cell; the others are re-derived from the
                        // ray at t_next and clamped into the cell's own range — a no-op at leaf granularity when
                        // L4 == 1 (the clamp range is the single leaf cell)
                        if (tx <= t_next) { nx = px_ ? cx + L4 : cx - 1; out |= (uint32_t)(nx - (int)(bnd0 & 0xFFFFu)) >= (bnd0 >> 16); }
                        else if (L4 > 1) { const int ip = __float2int_rd(__fmaf_rn(d.x, t_next, o.x)) >> 2; nx = min(max(ip, cx), cx + L4 - 1); }
                        if (ty <= t_next) { ny = py_ ? cy + L4 : cy - 1; out |= (uint32_t)(ny - (int)(bnd1 & 0xFFFFu)) >= (bnd1 >> 16); }
                        else if (L4 > 1) { const int ip = __float2int_rd(__fmaf_rn(d.y, t_next, o.y)) >> 2; ny = min(max(ip, cy), cy + L4 - 1); }
                        if (tz <= t_next) { nz = pz_ ? cz + L4 : cz - 1; out |= (uint32_t)(nz - (int)(bnd2 & 0xFFFFu)) >= (bnd2 >> 16); }
                        else if (L4 > 1) { const int ip = __float2int_rd(__fmaf_rn(d.z, t_next, o.z)) >> 2; nz = min(max(ip, cz), cz + L4 - 1); }
                        // the cached 4^3 node survives a step that stays inside its 16-cell
                        have_mid = have_mid && (((uint32_t)((nx ^ jx) | (ny ^ jy) | (nz ^ jz)) >> 2) == 0u);
                        jx = nx; jy = ny; jz = nz;
                        L4 = 0;
                        if (out) mode = M_NEXT;
                    }
                    if (mode == M_WALK) {
                        // descend to the cell containing (jx,jy,jz): InternalNode::get index arithmetic
                        // (crates/vdb/src/node/internal.rs:77-95) on the re-laid-out nodes
                        if (!have_mid) {
                            const DNode16 *node = n16;
                            bool alive = true;
                            if (NL == 4 || (NL == 0 && four)) {
                                const uint32_t tidx = ((uint32_t)(jx >> 6) << 8) | ((uint32_t)(jy >> 6) << 4) | (uint32_t)(jz >> 6);
                                const uint4 w = __ldg(&node->word[tidx >> 6]);
                                if (STATS) st.mask_words++;
                                const uint32_t b = tidx & 63u;
                                if (!bit64(w.x, w.y, b)) { L4 = 64; alive = false; }
                                else {
                                    if (STATS) st.child_ptrs++;
                                    node = n16 + (w.z + rank64(w.x, w.y, b));
                                }
                            }
                            if (alive) {
                                const uint32_t ridx = ((uint32_t)((jx >> 2) & 15) << 8) | ((uint32_t)((jy >> 2) & 15) << 4) | (uint32_t)((jz >> 2) & 15);
                                if (STATS) st.mask_words++;
                                const uint4 w = __ldg(&node->word[ridx >> 6]);
                                const uint32_t b = ridx & 63u;
                                L4 = 4;
                                if (bit64(w.x, w.y, b)) {
                                    const uint4 mid = __ldg(&mids[w.z + rank64(w.x, w.y, b)]);
                                    mid_lo = mid.x; mid_hi = mid.y; mid_base = mid.z;
                                    have_mid = true;
                                }
                            }
                        } else if (STATS) {
                            if (NL == 4 || (NL == 0 && four)) { st.mask_words++; st.child_ptrs++; }
                            st.mask_words++;
                        }
                        if (have_mid) {
                            if (STATS) { st.child_ptrs++; st.mask_words++; }
                            const uint32_t midx = ((uint32_t)(jx & 3) << 4) | ((uint32_t)(jy & 3) << 2) | (uint32_t)(jz & 3);
                            L4 = 1;
                            if (bit64(mid_lo, mid_hi, midx)) {
                                if (STATS) { st.child_ptrs++; st.mask_words++; }
                                leaf_i = mid_base + rank64(mid_lo, mid_hi, midx);
                                L4 = 0;
                                mode = M_LEAF;
                            }
                        }
                    }
                }
            } else if (op == M_LEAF) {
                if (mode == M_LEAF) {
                    // leaf step: one 128-bit load (occupancy + material_ptr), then hit.rint's DDA
                    const uint4 leaf = __ldg(&leaves[leaf_i]);
                    if (STATS) st.dbg[10]++;
                    const f3 lo = mk3(o.x - (float)(jx << 2), o.y - (float)(jy << 2), o.z - (float)(jz << 2));   // hit.rint:140
                    float th; uint32_t vox;
                    if (leaf_dda_inv<STATS>(lo, d, inv, leaf.x, leaf.y, tmin, th, vox, st.dbg[11]) && tmin <= th && th <= cur) {   // reportIntersectionEXT
                        if (any) { best.hit = true; best.inst = inst; mode = M_DONE; pend = -1; sp = DUST_SCRATCH_WORDS; }
                        else {
                            if (!best.hit || th < cur || inst < best.inst) {          // (t, instance) lexicographic
                                best.hit = true; best.t = th; cur = th; best.inst = inst; best.leaf = leaf_i;
                                best.pos = (uint32_t)jx | ((uint32_t)jy << 10) | ((uint32_t)jz << 20); best.voxel = vox;
                            }
                            mode = M_NEXT;                                             // first accepted hit is this instance's closest
                        }
                    } else { mode = M_WALK; L4 = 1; }
                }
            } else {
                if (mode == M_NEXT) {
                    mode = M_DONE;
                    if (j < nc) {
                        const uint32_t key = sm[j * DUST_THREADS];
                        ++j;
                        if (!(__uint_as_float(key & ~id_mask) > cur)) {               // sorted: nothing further can beat the best hit
                            // enter instance: world -> object ray, reciprocals, clip against the occupied bounds
                            mode = M_NEXT;
                            inst = key & id_mask;
                            if (STATS) st.dbg[12]++;
                            const DInstance &in = sc.inst[inst];
                            object_ray(in, ro, rd, o, d);
                            inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
                            const uint4 bw = __ldg(reinterpret_cast<const uint4 *>(in.bounds));
                            bnd0 = bw.x; bnd1 = bw.y; bnd2 = bw.z;
                            const int bx0 = (int)(bnd0 & 0xFFFFu) << 2, by0 = (int)(bnd1 & 0xFFFFu) << 2, bz0 = (int)(bnd2 & 0xFFFFu) << 2;
                            const int bx1 = bx0 + ((int)(bnd0 >> 16) << 2), by1 = by0 + ((int)(bnd1 >> 16) << 2), bz1 = bz0 + ((int)(bnd2 >> 16) << 2);
                            // plane times are (plane - o) * inv: difference first, plane*inv - o*inv would cancel (|o/d| >> t)
                            const float t0x = ((float)bx0 - o.x) * inv.x, t0y = ((float)by0 - o.y) * inv.y, t0z = ((float)bz0 - o.z) * inv.z;
                            const float t1x = ((float)bx1 - o.x) * inv.x, t1y = ((float)by1 - o.y) * inv.y, t1z = ((float)bz1 - o.z) * inv.z;
                            const float t_enter = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), tmin));
                            // walk a hair past tmax (acceptance stays exact): DESIGN.md "Traversal semantics"
                            const float tmax_pad = cur + (fabsf(cur) * 3.814697265625e-06f + 9.5367431640625e-07f);
                            t_exit = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), tmax_pad));
                            if (t_enter <= t_exit && bx1 > bx0) {
                                const int ix = __float2int_rd(__fmaf_rn(d.x, t_enter, o.x));
                                const int iy = __float2int_rd(__fmaf_rn(d.y, t_enter, o.y));
                                const int iz = __float2int_rd(__fmaf_rn(d.z, t_enter, o.z));
                                jx = min(max(ix, bx0), bx1 - 1) >> 2; jy = min(max(iy, by0), by1 - 1) >> 2; jz = min(max(iz, bz0), bz1 - 1) >> 2;
                                const ulonglong2 p01 = __ldg(reinterpret_cast<const ulonglong2 *>(&in.n16));
                                n16 = reinterpret_cast<const DNode16 *>(p01.x);
                                mids = reinterpret_cast<const uint4 *>(p01.y);
                                leaves = in.leaves;
                                if (NL == 0) four = bw.w == 4u;
                                have_mid = false;
                                L4 = 0;
                                mode = M_WALK;
                            }
                        }
                    }
                }
            }
        }
        if (!__any_sync(full, pend >= 0)) break;                      // candidate list overflowed: resume the cull
        nc = 0;
    }
}

}  // namespace dust
