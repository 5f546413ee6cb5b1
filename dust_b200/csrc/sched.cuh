// sm_100a device code, part 4: persistent lanes, fully operation-scheduled (the production kernels).
//
// kernels.cuh's frame_kernel votes among three operations INSIDE one ray; a lane whose ray has ended
// then idles until the slowest ray of its warp ends (16.8 of 32 lanes on the castle frame,
// profiles/r01_history.md).  Here a lane never waits for another lane's ray.  Every lane always has
// exactly one pending operation:
//     PATH   consume the finished ray (shade, accumulate), emit the path's next ray (sun shadow,
//            bounce, next sample) or claim a new pixel / ray from the global work counter
//     CULL   expand one node of the instance BVH (both children), collect touched instances
//     NEXT   enter the next candidate instance: world->object ray, reciprocals, clip
//     WALK   leave a cell / descend the masks
//     LEAF   hit.rint's 4x4x4 DDA on one leaf
// and each trip of the single loop the warp counts the lanes per operation with ONE redux.sync and
// executes the operation most lanes wait for.  This is warp-ballot compaction in time instead of in
// space: no ray state moves between lanes, all of it stays in registers, yet every operation body
// runs with a well-filled warp.  Arithmetic is expression for expression that of trace_core.cuh:
// results are independent of the scheduling (closest hit = lexicographic minimum of (t, instance)).
#pragma once
#include "kernels.cuh"

namespace dust {

enum : int { P_PATH = 0, P_CULL = 1, P_NEXT = 2, P_WALK = 3, P_LEAF = 4, P_DEAD = 5 };

struct RayState {
    f3 ro, rd;                 // world-space ray
    float tmin, cur;           // cur = current tmax: shrinks to the best hit (== its t once hit is set)
    bool any, hit;             // any: first accepted hit ends the ray
    uint32_t b_inst, b_leaf, b_pos, b_vox;
};

// Client::path_op(group, lane, ray, st) -> 0: lane is finished (no more work), 1: `ray` holds a new ray, 2: call again
template <int NL, bool STATS, class Client>
__device__ __forceinline__ void run_lanes(const DSceneView &sc, Client &cl, uint32_t *sm, Counters &st)
{
    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t id_mask = (1u << sc.id_bits) - 1u;

    RayState r;
    r.ro = mk3(0.0f, 0.0f, 0.0f); r.rd = mk3(1.0f, 1.0f, 1.0f); r.tmin = 0.0f; r.cur = 0.0f; r.any = false; r.hit = false;
    r.b_inst = 0xFFFFFFFFu; r.b_leaf = 0; r.b_pos = 0; r.b_vox = 0;
    int mode = P_PATH;
    int sp = DUST_SCRATCH_WORDS, pend = -1, nc = 0, j = 0;
    uint32_t inst = 0, leaf_i = 0;
    f3 o = mk3(0.0f, 0.0f, 0.0f), d = mk3(1.0f, 1.0f, 1.0f);
    f3 inv = mk3(1.0f, 1.0f, 1.0f);       // object-space 1/d while walking; world-space (cull) inverse while culling
    int jx = 0, jy = 0, jz = 0, L4 = 0;
    float t_exit = 0.0f;
    uint32_t bnd0 = 0, bnd1 = 0, bnd2 = 0;
    const DNode16 *n16 = nullptr;
    const uint4 *mids = nullptr, *leaves = nullptr;
    uint32_t mid_lo = 0, mid_hi = 0, mid_base = 0;
    bool have_mid = false, four = NL == 4;

    // box tests only: a zero component gets a huge FINITE inverse so (lo - o) * inv is never 0 * inf
    auto cull_inverse = [&]() {
        inv = mk3(fabsf(r.rd.x) < 1e-30f ? copysignf(1e30f, r.rd.x) : 1.0f / r.rd.x,
                  fabsf(r.rd.y) < 1e-30f ? copysignf(1e30f, r.rd.y) : 1.0f / r.rd.y,
                  fabsf(r.rd.z) < 1e-30f ? copysignf(1e30f, r.rd.z) : 1.0f / r.rd.z);
    };

#pragma unroll 1
    for (;;) {
        // one redux.sync counts the lanes of all five operations (6 bits each)
        const uint32_t packed = __reduce_add_sync(full, mode < P_DEAD ? (1u << (6 * mode)) : 0u);
        if (packed == 0u) break;
        const uint32_t n_path = packed & 63u, n_cull = (packed >> 6) & 63u, n_next = (packed >> 12) & 63u,
                       n_walk = (packed >> 18) & 63u, n_leaf = packed >> 24;
        // the operation most lanes wait for; ties go to the cheaper one (it feeds the dearer ones)
        int op = P_WALK;
        uint32_t cnt = n_walk;
        if (n_next > cnt) { op = P_NEXT; cnt = n_next; }
        if (n_cull > cnt) { op = P_CULL; cnt = n_cull; }
        if (n_leaf > cnt) { op = P_LEAF; cnt = n_leaf; }
        if (n_path > cnt) { op = P_PATH; cnt = n_path; }
        if (STATS && lane == 0) {
            st.dbg[op == P_WALK ? 0 : (op == P_LEAF ? 1 : (op == P_NEXT ? 2 : (op == P_CULL ? 15 : 14)))]++;
            st.dbg[op == P_WALK ? 3 : (op == P_LEAF ? 4 : (op == P_NEXT ? 5 : (op == P_CULL ? 13 : 12)))] += cnt;
            st.dbg[6] += 32u - (n_path + n_cull + n_next + n_walk + n_leaf);
        }

        if (op == P_WALK) {
            if (mode == P_WALK) {
                if (L4) {
                    if (STATS) st.dbg[L4 == 1 ? 7 : (L4 == 4 ? 8 : 9)]++;
                    // leave the cell of edge L4 that contains (jx,jy,jz)
                    const int cx = jx & ~(L4 - 1), cy = jy & ~(L4 - 1), cz = jz & ~(L4 - 1);
                    const bool px_ = d.x > 0.0f, py_ = d.y > 0.0f, pz_ = d.z > 0.0f;
                    const float tx = ((float)((px_ ? cx + L4 : cx) << 2) - o.x) * inv.x;
                    const float ty = ((float)((py_ ? cy + L4 : cy) << 2) - o.y) * inv.y;
                    const float tz = ((float)((pz_ ? cz + L4 : cz) << 2) - o.z) * inv.z;
                    const float t_next = fminf(fminf(tx, ty), tz);
                    bool out = !(t_next <= t_exit);                   // inclusive, DESIGN.md "Traversal semantics"
                    int nx = jx, ny = jy, nz = jz;
                    if (tx <= t_next) { nx = px_ ? cx + L4 : cx - 1; out |= (uint32_t)(nx - (int)(bnd0 & 0xFFFFu)) >= (bnd0 >> 16); }
                    else if (L4 > 1) { const int ip = __float2int_rd(__fmaf_rn(d.x, t_next, o.x)) >> 2; nx = min(max(ip, cx), cx + L4 - 1); }
                    if (ty <= t_next) { ny = py_ ? cy + L4 : cy - 1; out |= (uint32_t)(ny - (int)(bnd1 & 0xFFFFu)) >= (bnd1 >> 16); }
                    else if (L4 > 1) { const int ip = __float2int_rd(__fmaf_rn(d.y, t_next, o.y)) >> 2; ny = min(max(ip, cy), cy + L4 - 1); }
                    if (tz <= t_next) { nz = pz_ ? cz + L4 : cz - 1; out |= (uint32_t)(nz - (int)(bnd2 & 0xFFFFu)) >= (bnd2 >> 16); }
                    else if (L4 > 1) { const int ip = __float2int_rd(__fmaf_rn(d.z, t_next, o.z)) >> 2; nz = min(max(ip, cz), cz + L4 - 1); }
                    have_mid = have_mid && (((uint32_t)((nx ^ jx) | (ny ^ jy) | (nz ^ jz)) >> 2) == 0u);
                    jx = nx; jy = ny; jz = nz;
                    L4 = 0;
                    if (out) mode = P_NEXT;
                }
                if (mode == P_WALK) {
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
                            mode = P_LEAF;
                        }
                    }
                }
            }
        } else if (op == P_LEAF) {
            if (mode == P_LEAF) {
                // leaf step: one 128-bit load (occupancy + material_ptr), then hit.rint's DDA
                const uint4 leaf = __ldg(&leaves[leaf_i]);
                if (STATS) st.dbg[10]++;
                const f3 lo = mk3(o.x - (float)(jx << 2), o.y - (float)(jy << 2), o.z - (float)(jz << 2));   // hit.rint:140
                float th; uint32_t vox;
                if (leaf_dda_inv<STATS>(lo, d, inv, leaf.x, leaf.y, r.tmin, th, vox, st.dbg[11]) && r.tmin <= th && th <= r.cur) {   // reportIntersectionEXT
                    if (r.any) { r.hit = true; r.b_inst = inst; mode = P_PATH; }
                    else {
                        if (!r.hit || th < r.cur || inst < r.b_inst) {                // (t, instance) lexicographic
                            r.hit = true; r.cur = th; r.b_inst = inst; r.b_leaf = leaf_i;
                            r.b_pos = (uint32_t)jx | ((uint32_t)jy << 10) | ((uint32_t)jz << 20); r.b_vox = vox;
                        }
                        mode = P_NEXT;                                                 // first accepted hit is this instance's closest
                    }
                } else { mode = P_WALK; L4 = 1; }
            }
        } else if (op == P_NEXT) {
            if (mode == P_NEXT) {
                bool entered = false;
                if (j < nc) {
                    const uint32_t key = sm[j * DUST_THREADS];
                    ++j;
                    if (__uint_as_float(key & ~id_mask) > r.cur) j = nc;             // sorted: nothing further can beat the best hit
                    else {
                        // enter instance: world -> object ray, reciprocals, clip against the occupied bounds
                        inst = key & id_mask;
                        const DInstance &in = sc.inst[inst];
                        object_ray(in, r.ro, r.rd, o, d);
                        inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
                        const uint4 bw = __ldg(reinterpret_cast<const uint4 *>(in.bounds));
                        bnd0 = bw.x; bnd1 = bw.y; bnd2 = bw.z;
                        const int bx0 = (int)(bnd0 & 0xFFFFu) << 2, by0 = (int)(bnd1 & 0xFFFFu) << 2, bz0 = (int)(bnd2 & 0xFFFFu) << 2;
                        const int bx1 = bx0 + ((int)(bnd0 >> 16) << 2), by1 = by0 + ((int)(bnd1 >> 16) << 2), bz1 = bz0 + ((int)(bnd2 >> 16) << 2);
                        // plane times are (plane - o) * inv: difference first, plane*inv - o*inv would cancel (|o/d| >> t)
                        const float t0x = ((float)bx0 - o.x) * inv.x, t0y = ((float)by0 - o.y) * inv.y, t0z = ((float)bz0 - o.z) * inv.z;
                        const float t1x = ((float)bx1 - o.x) * inv.x, t1y = ((float)by1 - o.y) * inv.y, t1z = ((float)bz1 - o.z) * inv.z;
                        const float t_enter = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), r.tmin));
                        // walk a hair past tmax (acceptance stays exact): DESIGN.md "Traversal semantics"
                        const float tmax_pad = r.cur + (fabsf(r.cur) * 3.814697265625e-06f + 9.5367431640625e-07f);
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
                            mode = P_WALK;
                            entered = true;
                        }
                    }
                }
                if (!entered && j >= nc) {                                            // candidates exhausted
                    if (pend >= 0) { nc = 0; j = 0; cull_inverse(); mode = P_CULL; }  // the list had overflowed: resume the cull
                    else mode = P_PATH;                                               // ray finished
                }
            }
        } else if (op == P_CULL) {
            if (mode == P_CULL) {
                // expand BVH node `pend`: both children; instances go to the sorted candidate list (key =
                // entry t truncated downwards | instance), inner nodes to `pend` / the stack
                const float4 *n = reinterpret_cast<const float4 *>(sc.bvh + pend);
                const float4 a = __ldg(n), b = __ldg(n + 1), c = __ldg(n + 2), e = __ldg(n + 3);
                float t0 = box_entry(a.x, a.y, a.z, b.x, b.y, b.z, r.ro, inv, r.tmin, r.cur);
                float t1 = box_entry(c.x, c.y, c.z, e.x, e.y, e.z, r.ro, inv, r.tmin, r.cur);
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
                    } else {
                        const uint32_t key = (__float_as_uint(fmaxf(te, 0.0f)) & ~id_mask) | (uint32_t)~ch;
                        int k = nc++;
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
                if (pend < 0 && nc == 0) mode = P_PATH;                               // nothing touched: a miss
                else if (pend < 0 || nc + 3 > sp) { j = 0; mode = P_NEXT; }           // list complete (or full): walk it
            }
        } else {
            const unsigned group = __ballot_sync(full, mode == P_PATH);
            if (mode == P_PATH) {
                const int rc = cl.path_op(group, lane, r, st);
                if (rc == 0) mode = P_DEAD;
                else if (rc == 1) {
                    // launch the ray: reset the best hit, test the scene box
                    r.hit = false; r.b_inst = 0xFFFFFFFFu;
                    sp = DUST_SCRATCH_WORDS; pend = -1; nc = 0; j = 0;
                    cull_inverse();
                    if (sc.n_inst) {
                        const float te = box_entry(sc.root_lo[0], sc.root_lo[1], sc.root_lo[2], sc.root_hi[0], sc.root_hi[1], sc.root_hi[2],
                                                   r.ro, inv, r.tmin, r.cur);
                        if (te != CUDART_INF_F) {
                            if (sc.root_ref >= 0) { pend = sc.root_ref; mode = P_CULL; }
                            else { sm[0] = (__float_as_uint(fmaxf(te, 0.0f)) & ~id_mask) | (uint32_t)~sc.root_ref; nc = 1; mode = P_NEXT; }
                        }
                    }
                    // else / box missed: stays P_PATH and is consumed as a miss on the next PATH trip
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------- frame client

// One lane = one pixel at a time, its samples in order (so the per-pixel mean is accumulated exactly
// like the sequential loop).  Pixels are claimed from a global counter with one warp-aggregated
// atomicAdd for all requesting lanes; consecutive items follow a Morton curve inside 32x32 tiles.
template <bool STATS>
struct FrameClient {
    const FrameParams &p;
    unsigned long long *work;
    uint32_t n_items;
    bool have_pixel;
    int kind;
    uint32_t item, pxy, smp, s1, v;
    f3 acc, L, T, sn;

    __device__ __forceinline__ FrameClient(const FrameParams &p_, unsigned long long *work_, uint32_t n_items_)
        : p(p_), work(work_), n_items(n_items_), have_pixel(false), kind(RAY_NONE), item(0), pxy(0), smp(0), s1(0), v(0)
    {
        acc = L = sn = mk3(0.0f, 0.0f, 0.0f); T = mk3(1.0f, 1.0f, 1.0f);
    }

    __device__ __forceinline__ int path_op(const unsigned group, const unsigned lane, RayState &r, Counters &st)
    {
        const DustTraceConfig &cfg = p.cfg;
        const uint32_t px = pxy & 0xFFFFu, py = pxy >> 16;
        const uint32_t pixel_index = py * cfg.width + px;
        // ---- 1. consume the finished ray
        bool to_vertex = false, to_bounce = false, need_sky = false;
        if (kind == RAY_SHADOW) {
            if (!r.hit) {                                                              // nee.rmiss:11-22
                const float ndl = dot3(sn, mk3(p.sun_dir[0], p.sun_dir[1], p.sun_dir[2]));
                const f3 ts = acescg2srgb(mk3(p.sun_nee[0] * ndl, p.sun_nee[1] * ndl, p.sun_nee[2] * ndl));
                L.x += T.x * ts.x; L.y += T.y * ts.y; L.z += T.z * ts.z;
            }
            to_bounce = true;
        } else if (kind != RAY_NONE) {
            if (STATS) { if (kind == RAY_PRIMARY) st.primary_rays++; else st.bounce_rays++; }
            if (kind == RAY_PRIMARY && smp == 0 && (cfg.flags & DUST_TRACE_WRITE_HITS) && p.hit_instance) {
                p.hit_instance[pixel_index] = r.hit ? r.b_inst : 0xFFFFFFFFu; p.hit_leaf[pixel_index] = r.hit ? r.b_leaf : 0xFFFFFFFFu;
                p.hit_voxel[pixel_index] = r.hit ? (uint8_t)r.b_vox : (uint8_t)0xFF; p.hit_t[pixel_index] = r.hit ? r.cur : CUDART_INF_F;
            }
            if (r.hit) {
                Hit h;
                h.hit = true; h.t = r.cur; h.inst = r.b_inst; h.leaf = r.b_leaf; h.pos = r.b_pos; h.voxel = r.b_vox;
                const f3 a = surface_at<STATS>(p.scene, h, r.ro, r.rd, sn, st);
                if (kind == RAY_PRIMARY) { T = a; v = 0; }                             // tone_map.comp:195-203
                else { T.x = T.x * a.x; T.y = T.y * a.y; T.z = T.z * a.z; v += 1; }   // final_gather.rchit:76-80
                to_vertex = true;
            } else need_sky = true;
        }
        if (need_sky) {
            const f3 dn = normalize3(r.rd);
            f3 sk, su;
            sky_eval(p.sky, p.sun_ar2, dn.x, dn.y, dn.z, kind == RAY_PRIMARY, sk.x, sk.y, sk.z, su.x, su.y, su.z);
            if (kind == RAY_PRIMARY) {                                                 // miss.rmiss:8-14
                L = acescg2srgb(mk3((sk.x + su.x) / 3.14f, (sk.y + su.y) / 3.14f, (sk.z + su.z) / 3.14f));
            } else {                                                                   // final_gather.rmiss:13-17
                const f3 ts = acescg2srgb(sk);
                L.x += T.x * ts.x; L.y += T.y * ts.y; L.z += T.z * ts.z;
            }
            kind = RAY_NONE;
        }
        if (to_vertex) {
            if (dot3(mk3(p.sky.direction[0], p.sky.direction[1], p.sky.direction[2]), sn) > 0.0f) {   // ambient_occlusion.rgen:33-50
                if (STATS) st.shadow_rays++;
                r.rd = mk3(p.sun_dir[0], p.sun_dir[1], p.sun_dir[2]); r.tmin = 0.1f; r.cur = 10000.0f; r.any = true;
                kind = RAY_SHADOW;
            } else to_bounce = true;
        }
        if (to_bounce) {
            if (v == cfg.max_bounces) kind = RAY_NONE;
            else {
                const float u1 = u01(pcg(s1 + 2u + 2u * v)), u2 = u01(pcg(s1 + 3u + 2u * v));   // final_gather.rgen:32-34
                const float rr = sqrtf(u1), cz = sqrtf(1.0f - u1);
                float c, s;
                cossin_2pi(u2, c, s);
                r.rd = normalize3(rotate_by_normal(sn, mk3(rr * c, rr * s, cz)));
                r.tmin = 0.1f; r.cur = p.cam.far; r.any = false;
                kind = RAY_BOUNCE;
            }
        }
        // lanes whose path continues leave here; the others finish the sample together
        const unsigned rest = __ballot_sync(group, kind == RAY_NONE);
        if (kind != RAY_NONE) return 1;

        // ---- 2. the sample is finished: accumulate; the pixel may be finished: write it
        if (have_pixel) {
            const f3 Lc = srgb2acescg(L);
            acc.x += Lc.x; acc.y += Lc.y; acc.z += Lc.z;
            smp += 1;
            if (smp == cfg.spp) {
                const uint32_t lt = item >> 10, q = item & 1023u;
                const uint32_t out_index = cfg.shard_count > 1 ? lt * 1024u + even_bits(q >> 1) * 32u + even_bits(q) : pixel_index;
                const float n = (float)cfg.spp;
                const float4 outv = make_float4(acc.x / n, acc.y / n, acc.z / n, 1.0f);
                if (p.radiance32) p.radiance32[out_index] = outv;
                if (p.radiance16) {
                    __half2 a = __floats2half2_rn(outv.x, outv.y), b = __floats2half2_rn(outv.z, outv.w);
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t *>(&a); pk.y = *reinterpret_cast<uint32_t *>(&b);
                    reinterpret_cast<uint2 *>(p.radiance16)[out_index] = pk;
                }
                have_pixel = false;
            }
        }
        // ---- 3. claim a new pixel: one atomic for all requesting lanes of the group
        const unsigned want = __ballot_sync(rest, !have_pixel);
        if (want) {
            const int leader = __ffs(want) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(work, (unsigned long long)__popc(want));
            base = __shfl_sync(rest, base, leader);
            if (!have_pixel) {
                const unsigned long long it = base + (unsigned)__popc(want & ((1u << lane) - 1u));
                if (it >= n_items) return 0;
                item = (uint32_t)it;
                const uint32_t lt = item >> 10, q = item & 1023u;
                const uint32_t t = lt * cfg.shard_count + cfg.shard_index;
                const uint32_t tile_x = t % p.tiles_x, tile_y = t / p.tiles_x;
                const uint32_t nx = tile_x * 32u + even_bits(q), ny = tile_y * 32u + even_bits(q >> 1);
                if (!(tile_y < p.tiles_y && nx < cfg.width && ny < cfg.height)) return 2;   // padding of a ragged tile: claim another
                pxy = nx | (ny << 16);
                smp = 0; acc = mk3(0.0f, 0.0f, 0.0f);
                have_pixel = true;
            }
        }
        // ---- 4. primary ray of sample `smp` (primary.rgen:9-21)
        {
            const uint32_t qx = pxy & 0xFFFFu, qy = pxy >> 16;
            const uint32_t s0 = pcg((qy * cfg.width + qx) ^ pcg(cfg.seed + 0x9E3779B9u * cfg.frame_index));
            s1 = pcg(s0 + smp);
            float jx = 0.5f, jy = 0.5f;
            if (smp > 0) { jx = u01(pcg(s1 + 0u)); jy = u01(pcg(s1 + 1u)); }
            camera_ray(p.cam, cfg.width, cfg.height, qx, qy, jx, jy, r.ro, r.rd);
            r.tmin = p.cam.near; r.cur = p.cam.far; r.any = false;
            L = mk3(0.0f, 0.0f, 0.0f);
            kind = RAY_PRIMARY;
        }
        return 1;
    }
};

#ifndef DUST_PT_MIN_BLOCKS
#define DUST_PT_MIN_BLOCKS 4
#endif

template <int NL, bool STATS>
__global__ void __launch_bounds__(DUST_THREADS, DUST_PT_MIN_BLOCKS)
frame_kernel_pt(const __grid_constant__ FrameParams p, unsigned long long *work, const uint32_t n_items)
{
    __shared__ uint32_t s_scratch[DUST_SCRATCH_WORDS * DUST_THREADS];
    Counters st = {};
    FrameClient<STATS> cl(p, work, n_items);
    run_lanes<NL, STATS>(p.scene, cl, s_scratch + threadIdx.x, st);
    if (STATS) flush_counters(p.counters, st);
}

// --------------------------------------------------------------------------------- ray-batch client

template <bool STATS>
struct RayClient {
    const RayParams &p;
    unsigned long long *work;
    bool have_ray;
    unsigned long long index;

    __device__ __forceinline__ RayClient(const RayParams &p_, unsigned long long *work_) : p(p_), work(work_), have_ray(false), index(0) {}

    __device__ __forceinline__ int path_op(const unsigned group, const unsigned lane, RayState &r, Counters &st)
    {
        if (have_ray) {
            if (STATS) st.primary_rays++;
            uint4 o;
            if (r.hit) { o.x = __float_as_uint(r.cur); o.y = r.b_inst; o.z = r.b_leaf; o.w = r.b_vox; }
            else { o.x = __float_as_uint(CUDART_INF_F); o.y = o.z = o.w = 0xFFFFFFFFu; }
            reinterpret_cast<uint4 *>(p.hits)[index] = o;
        }
        // every lane of the group wants a ray: one atomic, consecutive rays to consecutive lanes (coalesced loads)
        const int leader = __ffs(group) - 1;
        unsigned long long base = 0;
        if ((int)lane == leader) base = atomicAdd(work, (unsigned long long)__popc(group));
        base = __shfl_sync(group, base, leader);
        index = base + (unsigned)__popc(group & ((1u << lane) - 1u));
        if (index >= p.n) { have_ray = false; return 0; }
        const float4 a = __ldg(reinterpret_cast<const float4 *>(p.rays) + 2 * index);      // 32-byte ray = two 128-bit loads
        const float4 b = __ldg(reinterpret_cast<const float4 *>(p.rays) + 2 * index + 1);
        r.ro = mk3(a.x, a.y, a.z); r.tmin = a.w; r.rd = mk3(b.x, b.y, b.z); r.cur = b.w; r.any = false;
        have_ray = true;
        return 1;
    }
};

template <int NL, bool STATS>
__global__ void __launch_bounds__(DUST_THREADS, DUST_PT_MIN_BLOCKS)
rays_kernel_pt(const __grid_constant__ RayParams p, unsigned long long *work)
{
    __shared__ uint32_t s_scratch[DUST_SCRATCH_WORDS * DUST_THREADS];
    Counters st = {};
    RayClient<STATS> cl(p, work);
    run_lanes<NL, STATS>(p.scene, cl, s_scratch + threadIdx.x, st);
    if (STATS) flush_counters(p.counters, st);
}

}  // namespace dust
