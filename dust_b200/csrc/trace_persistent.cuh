// Persistent-threads frame kernel (the production path of dust_trace).
//
// frame_kernel (trace_kernels.cuh) gives every pixel a thread that runs its whole path through
// nested loops (BVH loop > cell-walk loop > leaf-DDA loop).  On B200 that kernel issues at 57 % of
// peak but with 3.96 of 32 lanes active per instruction (profiles/r01_*): every nested loop waits
// for its slowest lane.  A first flat version (one step of whatever each lane needs per trip)
// reached 6.4 lanes: the expensive operations (leaf DDA, instance set-up, sky evaluation) were each
// wanted by only ~2 lanes per trip.  This kernel therefore schedules OPERATIONS, not lanes:
//
//   every lane always has exactly one pending operation
//     OP_PATH   consume the finished ray (shade, accumulate), emit the next ray of the path
//               (sun-shadow, bounce, next sample) or fetch a new pixel
//     OP_BVH    one node of the instance BVH (pop / test two children / reach an instance)
//     OP_SETUP  enter an instance: world->object ray, reciprocals, clip against the tree box
//     OP_WALK   one cell of the tree walk: descend the masks, leave an empty cell
//     OP_LEAF   the 4x4x4 leaf step (hit.rint) and leaving the leaf
//   each trip of the single loop the warp votes (5 ballots) and executes only the operation that the
//   most lanes are waiting for; the others keep waiting and their group grows until it wins.
//
// That is warp-ballot compaction in time instead of in space: no ray state is moved between lanes,
// yet every operation body runs with a well-filled warp.  Finished pixels are replaced from a global
// counter with one warp-aggregated atomicAdd (ballot -> popc -> atomic by the leader -> rank by lane
// mask), so a warp never idles behind its slowest ray.  Arithmetic is the same, expression for
// expression, as in trace_kernels.cuh: results are bit-identical to frame_kernel and to the CPU
// reference, independent of scheduling.
#pragma once
#include "trace_kernels.cuh"

namespace dust {

enum : int { OP_PATH = 0, OP_BVH = 1, OP_SETUP = 2, OP_WALK = 3, OP_LEAF = 4, OP_DEAD = 5 };
enum : int { RAY_NONE = -1, RAY_PRIMARY = 0, RAY_SHADOW = 1, RAY_BOUNCE = 2 };

#ifndef DUST_PT_THREADS
#define DUST_PT_THREADS 128
#endif
#ifndef DUST_PT_MIN_BLOCKS
#define DUST_PT_MIN_BLOCKS 4
#endif
#define DUST_PT_STACK 24

struct Lane {
    int op;
    // ---- pixel / path
    uint32_t px, py, out_index, pixel_index;
    uint32_t smp, s0, s1, v;
    bool have_pixel, want_fetch;
    int kind;
    f3 acc, L, T;
    f3 sx, sn;                 // current path vertex: offset hit point, world normal
    // ---- ray (world space)
    f3 ro, rd, winv;
    float tmin, cur;           // cur = current tmax (shrinks to the best hit)
    // ---- best hit of the ray
    bool hit;
    float b_t;
    uint32_t b_inst, b_leaf, b_pack0, b_pack1;
    // ---- BVH walk
    int pending, sp;
    // ---- instance walk
    uint32_t inst, leaf_i;
    f3 o, d, inv;
    int ix, iy, iz;
    float t_exit;
    uint32_t mid_key, mid_lo, mid_hi, mid_base;
};

// object-space ray of an instance (hit.rint:140-141) with the zero-direction guard
__device__ __forceinline__ void object_ray(const DInstance &in, f3 ow, f3 dw, f3 &o, f3 &d)
{
    const float *M = in.w2o;
    o.x = ((M[0] * ow.x + M[1] * ow.y) + M[2] * ow.z) + M[3];
    o.y = ((M[4] * ow.x + M[5] * ow.y) + M[6] * ow.z) + M[7];
    o.z = ((M[8] * ow.x + M[9] * ow.y) + M[10] * ow.z) + M[11];
    d.x = (M[0] * dw.x + M[1] * dw.y) + M[2] * dw.z;
    d.y = (M[4] * dw.x + M[5] * dw.y) + M[6] * dw.z;
    d.z = (M[8] * dw.x + M[9] * dw.y) + M[10] * dw.z;
    if (fabsf(d.x) < DUST_DIR_EPS) d.x = copysignf(DUST_DIR_EPS, d.x);
    if (fabsf(d.y) < DUST_DIR_EPS) d.y = copysignf(DUST_DIR_EPS, d.y);
    if (fabsf(d.z) < DUST_DIR_EPS) d.z = copysignf(DUST_DIR_EPS, d.z);
}

// Start the ray (ro, rd, tmin, cur): reset the best hit, test the BVH root.
__device__ __forceinline__ void launch_ray(Lane &l, const DSceneView &sc)
{
    l.hit = false;
    l.b_inst = 0xFFFFFFFFu;
    l.sp = 0;
    l.winv = mk3(fabsf(l.rd.x) < 1e-30f ? copysignf(1e30f, l.rd.x) : 1.0f / l.rd.x,
                 fabsf(l.rd.y) < 1e-30f ? copysignf(1e30f, l.rd.y) : 1.0f / l.rd.y,
                 fabsf(l.rd.z) < 1e-30f ? copysignf(1e30f, l.rd.z) : 1.0f / l.rd.z);
    if (sc.n_inst == 0 || box_entry(sc.bvh[0], l.ro, l.winv, l.tmin, l.cur) == CUDART_INF_F) {
        l.op = OP_PATH;          // nothing to traverse: consumed as a miss
        return;
    }
    l.pending = 0;
    l.op = OP_BVH;
}

// OP_BVH: one node.  pending >= 0: expand that node; pending < 0: pop the next candidate.
__device__ __forceinline__ void op_bvh(Lane &l, const DSceneView &sc, int *stack, int stride)
{
    if (l.pending < 0) {
        if (l.sp == 0) { l.op = OP_PATH; return; }                        // ray finished
        const int cand = stack[--l.sp * stride];
        if (box_entry(sc.bvh[cand], l.ro, l.winv, l.tmin, l.cur) == CUDART_INF_F) return;
        l.pending = cand;
        return;
    }
    const DBvhNode n = sc.bvh[l.pending];
    l.pending = -1;
    if (n.left < 0) { l.inst = (uint32_t)~n.left; l.op = OP_SETUP; return; }
    float tl = box_entry(sc.bvh[n.left], l.ro, l.winv, l.tmin, l.cur);
    float tr = box_entry(sc.bvh[n.right], l.ro, l.winv, l.tmin, l.cur);
    int near_ = n.left, far_ = n.right;
    if (tr < tl) { near_ = n.right; far_ = n.left; const float s = tl; tl = tr; tr = s; }
    if (tl != CUDART_INF_F) {
        if (tr != CUDART_INF_F && l.sp < DUST_PT_STACK) stack[l.sp++ * stride] = far_;
        l.pending = near_;
    }
}

// OP_SETUP: enter instance l.inst (same arithmetic as trace_instance's prologue)
template <int NL>
__device__ __forceinline__ void op_setup(Lane &l, const DSceneView &sc)
{
    const DInstance &in = sc.inst[l.inst];
    l.op = OP_BVH;
    object_ray(in, l.ro, l.rd, l.o, l.d);
    l.inv = mk3(1.0f / l.d.x, 1.0f / l.d.y, 1.0f / l.d.z);
    const int bx0 = in.bmin[0], by0 = in.bmin[1], bz0 = in.bmin[2], bx1 = in.bmax[0], by1 = in.bmax[1], bz1 = in.bmax[2];
    const float t0x = ((float)bx0 - l.o.x) * l.inv.x, t0y = ((float)by0 - l.o.y) * l.inv.y, t0z = ((float)bz0 - l.o.z) * l.inv.z;
    const float t1x = ((float)bx1 - l.o.x) * l.inv.x, t1y = ((float)by1 - l.o.y) * l.inv.y, t1z = ((float)bz1 - l.o.z) * l.inv.z;
    const float t_enter = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), l.tmin));
    const float tmax_pad = l.cur + (fabsf(l.cur) * 3.814697265625e-06f + 9.5367431640625e-07f);
    l.t_exit = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), tmax_pad));
    if (!(t_enter <= l.t_exit) || bx1 <= bx0) return;
    const int ix = (int)floorf(__fmaf_rn(l.d.x, t_enter, l.o.x));
    const int iy = (int)floorf(__fmaf_rn(l.d.y, t_enter, l.o.y));
    const int iz = (int)floorf(__fmaf_rn(l.d.z, t_enter, l.o.z));
    l.ix = min(max(ix, bx0), bx1 - 1); l.iy = min(max(iy, by0), by1 - 1); l.iz = min(max(iz, bz0), bz1 - 1);
    l.mid_key = 0;
    l.op = OP_WALK;
}

// leave the cell of edge L that contains (ix,iy,iz); op becomes OP_WALK, or OP_BVH when the walk ends
template <int NL>
__device__ __forceinline__ void leave_cell(Lane &l, const DInstance &in, int L)
{
    const int cx = l.ix & ~(L - 1), cy = l.iy & ~(L - 1), cz = l.iz & ~(L - 1);
    const bool px_ = l.d.x > 0.0f, py_ = l.d.y > 0.0f, pz_ = l.d.z > 0.0f;
    const float tx = ((float)(px_ ? cx + L : cx) - l.o.x) * l.inv.x;
    const float ty = ((float)(py_ ? cy + L : cy) - l.o.y) * l.inv.y;
    const float tz = ((float)(pz_ ? cz + L : cz) - l.o.z) * l.inv.z;
    const float t_next = fminf(fminf(tx, ty), tz);
    bool out = !(t_next <= l.t_exit);
    int nx, ny, nz;
    if (tx <= t_next) { nx = px_ ? cx + L : cx - 1; out |= (nx < in.bmin[0] || nx >= in.bmax[0]); }
    else { const int ip = (int)floorf(__fmaf_rn(l.d.x, t_next, l.o.x)); nx = min(max(ip, cx), cx + L - 1); }
    if (ty <= t_next) { ny = py_ ? cy + L : cy - 1; out |= (ny < in.bmin[1] || ny >= in.bmax[1]); }
    else { const int ip = (int)floorf(__fmaf_rn(l.d.y, t_next, l.o.y)); ny = min(max(ip, cy), cy + L - 1); }
    if (tz <= t_next) { nz = pz_ ? cz + L : cz - 1; out |= (nz < in.bmin[2] || nz >= in.bmax[2]); }
    else { const int ip = (int)floorf(__fmaf_rn(l.d.z, t_next, l.o.z)); nz = min(max(ip, cz), cz + L - 1); }
    l.ix = nx; l.iy = ny; l.iz = nz;
    l.op = out ? OP_BVH : OP_WALK;
}

// OP_WALK: descend to the cell containing (ix,iy,iz); an occupied leaf defers to OP_LEAF
template <int NL, bool STATS>
__device__ __forceinline__ void op_walk(Lane &l, const DSceneView &sc, Counters &st)
{
    const DInstance &in = sc.inst[l.inst];
    const int ix = l.ix, iy = l.iy, iz = l.iz;
    int L;
    const DNode16 *n16 = in.n16;
    bool alive = true;
    uint32_t key = 1;
    if (NL == 4) {
        const uint32_t tidx = ((uint32_t)(ix >> 8) << 8) | ((uint32_t)(iy >> 8) << 4) | (uint32_t)(iz >> 8);
        const uint64_t m = __ldg(&n16->mask[tidx >> 6]);
        if (STATS) st.mask_words++;
        const uint32_t b = tidx & 63u;
        if (!((m >> b) & 1ull)) { L = 256; alive = false; }
        else {
            if (STATS) st.child_ptrs++;
            n16 = in.n16 + (__ldg(&n16->child[tidx >> 6]) + __popcll(m & ((1ull << b) - 1ull)));
            key += tidx << 12;
        }
    }
    if (alive) {
        const uint32_t ridx = ((uint32_t)((ix >> 4) & 15) << 8) | ((uint32_t)((iy >> 4) & 15) << 4) | (uint32_t)((iz >> 4) & 15);
        key += ridx;
        if (STATS) st.mask_words++;
        bool have_mid = l.mid_key == key;            // the 4^3 node of this 16-cell is still in registers
        L = 16;
        if (!have_mid) {
            const uint64_t m = __ldg(&n16->mask[ridx >> 6]);
            const uint32_t b = ridx & 63u;
            if ((m >> b) & 1ull) {
                const uint32_t mid_i = __ldg(&n16->child[ridx >> 6]) + __popcll(m & ((1ull << b) - 1ull));
                const uint4 mid = __ldg(&in.mids[mid_i]);
                l.mid_key = key; l.mid_lo = mid.x; l.mid_hi = mid.y; l.mid_base = mid.z;
                have_mid = true;
            }
        }
        if (have_mid) {
            if (STATS) { st.child_ptrs++; st.mask_words++; }
            const uint64_t mm = u64_of(l.mid_lo, l.mid_hi);
            const uint32_t midx = ((uint32_t)((ix >> 2) & 3) << 4) | ((uint32_t)((iy >> 2) & 3) << 2) | (uint32_t)((iz >> 2) & 3);
            L = 4;
            if ((mm >> midx) & 1ull) {
                if (STATS) { st.child_ptrs++; st.mask_words++; }
                l.leaf_i = l.mid_base + __popcll(mm & ((1ull << midx) - 1ull));
                l.op = OP_LEAF;
                return;
            }
        }
    }
    leave_cell<NL>(l, in, L);
}

// OP_LEAF: the leaf step (hit.rint) on leaf l.leaf_i, then leave the leaf
template <int NL>
__device__ __forceinline__ void op_leaf(Lane &l, const DSceneView &sc)
{
    const DInstance &in = sc.inst[l.inst];
    const uint4 leaf = __ldg(&in.leaves[l.leaf_i]);                      // one 128-bit load: occupancy + material_ptr
    const uint64_t grid = u64_of(leaf.x, leaf.y);
    const int cx = l.ix & ~3, cy = l.iy & ~3, cz = l.iz & ~3;
    const f3 lo = mk3(l.o.x - (float)cx, l.o.y - (float)cy, l.o.z - (float)cz);   // hit.rint:140
    float th; uint32_t vox;
    if (leaf_dda_inv(lo, l.d, l.inv, grid, l.tmin, th, vox) && l.tmin <= th && th <= l.cur) {
        if (l.kind == RAY_SHADOW) { l.hit = true; l.op = OP_PATH; return; }     // any hit ends a shadow ray
        if (!l.hit || th < l.b_t || (th == l.b_t && l.inst < l.b_inst)) {
            l.hit = true; l.b_t = th; l.cur = th; l.b_inst = l.inst; l.b_leaf = l.leaf_i;
            l.b_pack0 = (uint32_t)cx | ((uint32_t)cy << 12) | (vox << 24);
            l.b_pack1 = (uint32_t)cz;
        }
        l.op = OP_BVH;                                                  // first accepted hit is this instance's closest
        return;
    }
    leave_cell<NL>(l, in, 4);
}

// hit.rchit:16-76 for the lane's best hit of world ray (ro, rd): fills the path vertex (sx, sn) and
// returns the linear albedo
template <bool STATS>
__device__ __forceinline__ f3 surface_at_best(Lane &l, const DSceneView &sc, Counters &st)
{
    const DInstance &in = sc.inst[l.b_inst];
    f3 o, d;
    object_ray(in, l.ro, l.rd, o, d);
    const uint32_t vox = l.b_pack0 >> 24;
    const int lx = (int)(l.b_pack0 & 4095u), ly = (int)((l.b_pack0 >> 12) & 4095u), lz = (int)l.b_pack1;
    const float t = l.b_t;
    const float vx = (t * d.x + o.x) - (((float)lx + (float)((vox >> 4) & 3u)) + 0.5f);
    const float vy = (t * d.y + o.y) - (((float)ly + (float)((vox >> 2) & 3u)) + 0.5f);
    const float vz = (t * d.z + o.z) - (((float)lz + (float)(vox & 3u)) + 0.5f);
    const float ax = fabsf(vx), ay = fabsf(vy), az = fabsf(vz);
    const float mx = fmaxf(ax, fmaxf(ay, az));
    const f3 n_obj = mk3(signf(vx) * (ax >= mx ? 1.0f : 0.0f), signf(vy) * (ay >= mx ? 1.0f : 0.0f), signf(vz) * (az >= mx ? 1.0f : 0.0f));
    const float *W = in.o2w;
    l.sn = mk3((W[0] * n_obj.x + W[1] * n_obj.y) + W[2] * n_obj.z,
               (W[3] * n_obj.x + W[4] * n_obj.y) + W[5] * n_obj.z,
               (W[6] * n_obj.x + W[7] * n_obj.y) + W[8] * n_obj.z);
    const uint4 leaf = __ldg(&in.leaves[l.b_leaf]);
    const uint64_t mask = u64_of(leaf.x, leaf.y);
    const uint32_t off = __popcll(mask & ((1ull << vox) - 1ull));                 // hit.rchit:60-61
    const uint8_t pi = __ldg(&in.materials[leaf.z + off]);                        // :71
    const float4 c = __ldg(&in.palette_lin[pi]);                                  // :72-74 + SRGBToLinear
    l.sx = mk3((t * l.rd.x + l.ro.x) + l.sn.x * 0.01f, (t * l.rd.y + l.ro.y) + l.sn.y * 0.01f, (t * l.rd.z + l.ro.z) + l.sn.z * 0.01f);
    if (STATS) st.hits_shaded++;
    return mk3(c.x, c.y, c.z);
}

// OP_PATH body: runs until the lane has a ray in flight (op != OP_PATH) or needs a pixel (want_fetch)
template <bool STATS>
__device__ __forceinline__ void path_step(Lane &l, const FrameParams &p, Counters &st)
{
    const DustTraceConfig &cfg = p.cfg;
    const f3 sun_dir = mk3(p.sun_dir[0], p.sun_dir[1], p.sun_dir[2]);
    const f3 sky_dir = mk3(p.sky.direction[0], p.sky.direction[1], p.sky.direction[2]);
    enum { GO_DISPATCH, GO_VERTEX, GO_AFTER_NEE, GO_FINISH } go = GO_DISPATCH;
#pragma unroll 1
    for (;;) {
        if (go == GO_DISPATCH) {
            if (l.kind == RAY_NONE) {
                if (!l.have_pixel || l.smp >= cfg.spp) {
                    if (l.have_pixel) {
                        const float n = (float)cfg.spp;
                        const float4 outv = make_float4(l.acc.x / n, l.acc.y / n, l.acc.z / n, 1.0f);
                        if (p.radiance32) p.radiance32[l.out_index] = outv;
                        if (p.radiance16) {
                            __half2 a = __floats2half2_rn(outv.x, outv.y), b = __floats2half2_rn(outv.z, outv.w);
                            uint2 pk;
                            pk.x = *reinterpret_cast<uint32_t *>(&a); pk.y = *reinterpret_cast<uint32_t *>(&b);
                            reinterpret_cast<uint2 *>(p.radiance16)[l.out_index] = pk;
                        }
                        l.have_pixel = false;
                    }
                    l.want_fetch = true;
                    return;
                }
                l.s1 = pcg(l.s0 + l.smp);
                float jx = 0.5f, jy = 0.5f;
                if (l.smp > 0) { jx = u01(pcg(l.s1 + 0u)); jy = u01(pcg(l.s1 + 1u)); }
                camera_ray(p.cam, cfg.width, cfg.height, l.px, l.py, jx, jy, l.ro, l.rd);
                l.tmin = p.cam.near; l.cur = p.cam.far;
                l.kind = RAY_PRIMARY;
                l.L = mk3(0.0f, 0.0f, 0.0f);
                launch_ray(l, p.scene);
                if (l.op != OP_PATH) return;
            } else if (l.kind == RAY_PRIMARY) {
                if (STATS) st.primary_rays++;
                const bool write_hits = l.smp == 0 && (cfg.flags & DUST_TRACE_WRITE_HITS) && p.hit_instance;
                if (!l.hit) {
                    const f3 dn = normalize3(l.rd);                                  // miss.rmiss:8-14
                    const f3 sk = sky_radiance(p.sky, dn);
                    const f3 su = sun_radiance(p.sky, p.sun_ar2, dn);
                    l.L = acescg2srgb(mk3((sk.x + su.x) / 3.14f, (sk.y + su.y) / 3.14f, (sk.z + su.z) / 3.14f));
                    if (write_hits) {
                        p.hit_instance[l.pixel_index] = 0xFFFFFFFFu; p.hit_leaf[l.pixel_index] = 0xFFFFFFFFu;
                        p.hit_voxel[l.pixel_index] = 0xFF; p.hit_t[l.pixel_index] = CUDART_INF_F;
                    }
                    go = GO_FINISH;
                } else {
                    if (write_hits) {
                        p.hit_instance[l.pixel_index] = l.b_inst; p.hit_leaf[l.pixel_index] = l.b_leaf;
                        p.hit_voxel[l.pixel_index] = (uint8_t)(l.b_pack0 >> 24); p.hit_t[l.pixel_index] = l.b_t;
                    }
                    l.T = surface_at_best<STATS>(l, p.scene, st);                    // tone_map.comp:195-203
                    l.v = 0;
                    go = GO_VERTEX;
                }
            } else if (l.kind == RAY_SHADOW) {
                if (!l.hit) {                                                        // nee.rmiss:11-22
                    const f3 su = sun_radiance(p.sky, p.sun_ar2, normalize3(sun_dir));
                    const float ndl = dot3(l.sn, sun_dir);
                    const f3 ts = acescg2srgb(mk3((su.x * p.sun_cone) * ndl, (su.y * p.sun_cone) * ndl, (su.z * p.sun_cone) * ndl));
                    l.L.x += l.T.x * ts.x; l.L.y += l.T.y * ts.y; l.L.z += l.T.z * ts.z;
                }
                go = GO_AFTER_NEE;
            } else {                                                                 // RAY_BOUNCE
                if (STATS) st.bounce_rays++;
                if (!l.hit) {                                                        // final_gather.rmiss:13-17
                    const f3 ts = acescg2srgb(sky_radiance(p.sky, normalize3(l.rd)));
                    l.L.x += l.T.x * ts.x; l.L.y += l.T.y * ts.y; l.L.z += l.T.z * ts.z;
                    go = GO_FINISH;
                } else {
                    const f3 a = surface_at_best<STATS>(l, p.scene, st);             // final_gather.rchit:76-80
                    l.T.x = l.T.x * a.x; l.T.y = l.T.y * a.y; l.T.z = l.T.z * a.z;
                    l.v += 1;
                    go = GO_VERTEX;
                }
            }
        } else if (go == GO_VERTEX) {
            go = GO_AFTER_NEE;
            if (dot3(sky_dir, l.sn) > 0.0f) {                                        // ambient_occlusion.rgen:33-50
                if (STATS) st.shadow_rays++;
                l.ro = l.sx; l.rd = sun_dir; l.tmin = 0.1f; l.cur = 10000.0f;
                l.kind = RAY_SHADOW;
                launch_ray(l, p.scene);
                if (l.op != OP_PATH) return;
                go = GO_DISPATCH;
            }
        } else if (go == GO_AFTER_NEE) {
            if (l.v == cfg.max_bounces) { go = GO_FINISH; continue; }
            const float u1 = u01(pcg(l.s1 + 2u + 2u * l.v)), u2 = u01(pcg(l.s1 + 3u + 2u * l.v));   // final_gather.rgen:32-34
            const float r = sqrtf(u1), cz = sqrtf(1.0f - u1);
            float c, sn;
            cossin_2pi(u2, c, sn);
            l.rd = normalize3(rotate_by_normal(l.sn, mk3(r * c, r * sn, cz)));
            l.ro = l.sx; l.tmin = 0.1f; l.cur = p.cam.far;
            l.kind = RAY_BOUNCE;
            launch_ray(l, p.scene);
            if (l.op != OP_PATH) return;
            go = GO_DISPATCH;
        } else {                                                                     // GO_FINISH: sample done
            const f3 Lc = srgb2acescg(l.L);
            l.acc.x += Lc.x; l.acc.y += Lc.y; l.acc.z += Lc.z;
            l.smp += 1;
            l.kind = RAY_NONE;
            go = GO_DISPATCH;
        }
    }
}

// OP_PATH for the lanes in `group` (all of them execute this together): path logic + pixel fetch
template <bool STATS>
__device__ __forceinline__ void op_path(Lane &l, const FrameParams &p, Counters &st, unsigned group, unsigned lane,
                                        unsigned int *work_counter, uint32_t n_items, uint32_t key)
{
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
        if (!l.want_fetch && l.op == OP_PATH) path_step<STATS>(l, p, st);
        const unsigned want = __ballot_sync(group, l.want_fetch);
        if (!want) break;
        // warp-aggregated work fetch: one atomic for all requesting lanes
        const int leader = __ffs(want) - 1;
        unsigned base = 0;
        if ((int)lane == leader) base = atomicAdd(work_counter, (unsigned)__popc(want));
        base = __shfl_sync(group, base, leader);
        if (l.want_fetch) {
            const unsigned item = base + (unsigned)__popc(want & ((1u << lane) - 1u));
            if (item >= n_items) { l.op = OP_DEAD; l.want_fetch = false; }
            else {
                // item -> (local tile, 8x4 patch, lane in patch): consecutive items share a patch
                const uint32_t lt = item >> 10, b = (item >> 5) & 31u, q = item & 31u;
                const uint32_t t = lt * p.cfg.shard_count + p.cfg.shard_index;
                const uint32_t tile_x = t % p.tiles_x, tile_y = t / p.tiles_x;
                const uint32_t lx = (b & 3u) * 8u + (q & 7u), ly = (b >> 2) * 4u + (q >> 3);
                const uint32_t px = tile_x * 32u + lx, py = tile_y * 32u + ly;
                if (tile_y < p.tiles_y && px < p.cfg.width && py < p.cfg.height) {
                    l.px = px; l.py = py;
                    l.pixel_index = py * p.cfg.width + px;
                    l.out_index = p.cfg.shard_count > 1 ? lt * 1024u + ly * 32u + lx : l.pixel_index;
                    l.s0 = pcg(l.pixel_index ^ key);
                    l.smp = 0; l.acc = mk3(0.0f, 0.0f, 0.0f);
                    l.have_pixel = true; l.kind = RAY_NONE;
                    l.want_fetch = false;
                }
                // else: padding of a ragged tile — keep want_fetch and take another item next time
            }
        }
    }
}

template <int NL, bool STATS>
__global__ void __launch_bounds__(DUST_PT_THREADS, DUST_PT_MIN_BLOCKS)
frame_kernel_persistent(const __grid_constant__ FrameParams p, unsigned int *work_counter, const uint32_t n_items)
{
    __shared__ int s_stack[DUST_PT_STACK * DUST_PT_THREADS];
    int *stack = s_stack + threadIdx.x;                  // lane-interleaved: level k lives at stack[k * blockDim]
    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    Counters st = { 0, 0, 0, 0, 0, 0 };
    Lane l;
    l.op = OP_PATH;
    l.kind = RAY_NONE;
    l.have_pixel = false; l.want_fetch = false;
    l.smp = 0;
    l.hit = false;
    l.pending = -1; l.sp = 0;
    l.acc = mk3(0.0f, 0.0f, 0.0f);
    const uint32_t key = pcg(p.cfg.seed + 0x9E3779B9u * p.cfg.frame_index);

#pragma unroll 1
    for (;;) {
        // vote: which operation do most lanes wait for?
        const unsigned m_path = __ballot_sync(full, l.op == OP_PATH);
        const unsigned m_bvh = __ballot_sync(full, l.op == OP_BVH);
        const unsigned m_setup = __ballot_sync(full, l.op == OP_SETUP);
        const unsigned m_walk = __ballot_sync(full, l.op == OP_WALK);
        const unsigned m_leaf = __ballot_sync(full, l.op == OP_LEAF);
        if ((m_path | m_bvh | m_setup | m_walk | m_leaf) == 0) break;      // every lane is OP_DEAD
        // argmax of the group sizes; ties go to the operation later in the pipeline
        int best = OP_LEAF, cnt = __popc(m_leaf);
        int c = __popc(m_walk);  if (c > cnt) { cnt = c; best = OP_WALK; }
        c = __popc(m_setup);     if (c > cnt) { cnt = c; best = OP_SETUP; }
        c = __popc(m_bvh);       if (c > cnt) { cnt = c; best = OP_BVH; }
        c = __popc(m_path);      if (c > cnt) { cnt = c; best = OP_PATH; }
        if (l.op == best) {
            if (best == OP_WALK) op_walk<NL, STATS>(l, p.scene, st);
            else if (best == OP_LEAF) op_leaf<NL>(l, p.scene);
            else if (best == OP_BVH) op_bvh(l, p.scene, stack, DUST_PT_THREADS);
            else if (best == OP_SETUP) op_setup<NL>(l, p.scene);
            else op_path<STATS>(l, p, st, m_path, lane, work_counter, n_items, key);
        }
    }
    if (STATS) flush_counters(p.counters, st);
}

}  // namespace dust
