// sm_100a device code, part 4: the wavefront pipeline of dust_trace (frames with GI bounces).
//
// frame_kernel (kernels.cuh) keeps a path in one thread from camera ray to pixel.  That is ideal for
// coherent rays (primary-only frames: 27 of 32 lanes busy per traversal step), but diffuse bounce rays
// of neighbouring pixels meet unrelated fates: on the castle frame the warp then runs with 13 busy
// lanes and 17 lanes waiting for the slowest ray of the warp (profiles/r01_history.md).  The wavefront
// pipeline cuts a path at its vertices and keeps rays, not paths, in flight:
//
//   (depth-0 queue)  implicit: slot i is the camera ray of path i, recomputed by the kernels that need it
//   per depth v = 0 .. max_bounces:
//     trace          closest hit of every ray of the queue                           -> hit records
//     wf_shade       apply the previous vertex's sun visibility, consume the hit (sky on a miss,
//                    surface on a hit), emit the sun-shadow ray -> shadow queue and the cosine-weighted
//                    bounce ray -> next closest queue (compacted, one atomic per CTA)
//     trace          any hit of every shadow ray                                     -> visibility byte per path
//   wf_resolve       last visibility, sRGB -> ACEScg, mean over samples in sample order -> frame
//
// trace = wf_trace (one thread per ray) for the coherent depth-0 queues; for the incoherent queues
// behind a bounce it is wf_cull + wf_trace_pt + wf_trace_slow: candidate instances are collected with all lanes
// busy (the 8 nearest per ray), persistent lanes walk them and CLAIM THE NEXT RAY of the queue the moment their own
// is finished, and the few rays whose truncated list was not conclusive are redone one CTA per ray.  Sun shadow rays take
// their candidates from the sun grid (kernels.cuh wf_sun_grid) instead of the BVH, camera rays from their tile's list.
// Every ray is traced by the same walk code as in the megakernel and every float of a path is
// accumulated in the same order, so results are bit-identical to frame_kernel; queue order (atomics)
// varies from run to run but no result depends on it.  (Sorting the bounce rays by direction bin was
// tried and removed: with refilling lanes it only added a pass and scattered the path-state accesses.)
// Path state and queues live in HBM: ~150 B per path per batch.
#pragma once
#include "kernels.cuh"

namespace dust {

struct WfQueue {
    float4 *a;          // origin.xyz, tmin   of ray i at a[i * stride]
    float4 *b;          // direction.xyz, tmax           b[i * stride]
    uint32_t *path;     // path index; bit 31: any-hit (shadow) ray.  NULL: closest-hit rays, path = i (ray batches)
    uint32_t stride;    // 1: separate arrays; 2: interleaved DustRay records (a = rays, b = rays + 1);
                        // 0: no arrays at all — slot i is the camera ray of path i (wf_primary_ray), the depth-0 queue of a frame
    // Compact records of the frame's own queues (everything a consumer can know from the launch is not stored per ray):
    //   kind 0  as described by `stride`
    //   kind 1  any-hit (sun shadow) rays: a[i] = (origin, path word); direction and tmax are the uniform `ub`, tmin is `utmin` — 16 bytes per ray
    //   kind 2  closest-hit (bounce) rays: a[i] = (origin, path word), b[i] = (direction, tmax); tmin is `utmin`        — 32 bytes per ray
    // (36 bytes each before: 0.6 GB less DRAM traffic per configs[2] frame, profiles/r02_history.md)
    uint32_t kind;
    float utmin;
    float4 ub;
};

// ray i of a queue: origin + tmin, direction + tmax, path word (bit 31: any-hit ray).  KIND >= 0: the queue's kind is known at compile time.
template <int KIND = -1>
__device__ __forceinline__ void wf_fetch(const WfQueue &q, uint32_t i, float4 &a, float4 &b, uint32_t &pw)
{
    const uint32_t kind = KIND >= 0 ? (uint32_t)KIND : q.kind;
    if (kind == 0u) {
        a = __ldg(&q.a[(size_t)i * q.stride]); b = __ldg(&q.b[(size_t)i * q.stride]);
        pw = q.path ? __ldg(&q.path[i]) : i;
        return;
    }
    const float4 r = __ldg(&q.a[i]);
    pw = __float_as_uint(r.w);
    a = make_float4(r.x, r.y, r.z, q.utmin);
    if (kind == 1u) b = q.ub; else b = __ldg(&q.b[i]);
}

struct WfParams {
    FrameParams f;
    uint32_t n_slots;                 // pixel slots of this shard (tiles * 1024, Morton order inside a tile)
    uint32_t n_paths;                 // n_slots * batch_spp
    uint32_t batch_spp, batch_first;  // samples per pixel in this batch, index of its first sample
    uint32_t depth;                   // vertex index of the closest queue being traced / shaded
    uint32_t last_batch;
    float4 *ps0, *ps1;                // path state: (T, s1) (L, n.l of the pending sun term or 0)
    WfQueue qc, qs, qn;               // closest queue of this depth, shadow queue, closest queue of the next depth
    uint4 *hits;                      // per closest-queue slot: t, instance | voxel << 20, leaf, leaf origin
    DustHit *hits_out;                // != NULL (dust_trace_rays): write the caller's DustHit records instead
    uint8_t *vis;                     // per path: 1 = its pending shadow ray was occluded
    uint4 *cand;                      // per queue slot: WF_CANDS / 4 x uint4 = up to WF_CANDS sorted candidate keys (wf_cull)
    uint32_t *overflow;               // queue slots whose candidate list did not fit: traced by wf_trace with the full cull
    uint32_t *counts;                 // per depth: [4v+1] shadow rays, [4v+2] bounce rays (= next closest queue); [3] = n_paths
    float4 *acc;                      // per pixel slot, carried across batches
};

#ifndef WF_CANDS
#define WF_CANDS 8                    // pre-culled candidates per ray (multiple of 4); rays touching more instances take the slow path
#endif
#define WF_CVECS (WF_CANDS / 4)

__device__ __forceinline__ uint32_t wf_in_count(const WfParams &p)
{
    return p.depth == 0 ? p.n_paths : p.counts[4 * (p.depth - 1) + 2];
}

// pixel of slot (Morton inside 32x32 tiles; tile t = lt * shard_count + shard_index)
__device__ __forceinline__ bool wf_pixel(const FrameParams &f, uint32_t slot, uint32_t &px, uint32_t &py, uint32_t &out_index)
{
    const uint32_t lt = slot >> 10, q = slot & 1023u;
    const uint32_t lx = even_bits(q), ly = even_bits(q >> 1);
    const uint32_t t = lt * f.cfg.shard_count + f.cfg.shard_index;
    const uint32_t tile_x = t % f.tiles_x, tile_y = t / f.tiles_x;
    px = tile_x * 32u + lx; py = tile_y * 32u + ly;
    out_index = f.cfg.shard_count > 1 ? lt * 1024u + ly * 32u + lx : py * f.cfg.width + px;
    return tile_y < f.tiles_y && px < f.cfg.width && py < f.cfg.height;
}

// Camera ray of path i (primary.rgen:9-21) and the path's PRNG stream; paths of pixels outside the frame get a ray
// with tmax < tmin, which traces nothing.  Computed where it is needed (wf_trace, wf_shade) instead of being
// written to and re-read from a queue: 100 B per path of HBM traffic less than a generated depth-0 queue.
__device__ __forceinline__ bool wf_primary_ray(const WfParams &p, uint32_t i, float4 &a, float4 &b, uint32_t &s1, uint32_t &smp,
                                               uint32_t &px, uint32_t &py, uint32_t &oi)
{
    const DustTraceConfig &cfg = p.f.cfg;
    const uint32_t slot = i / p.batch_spp;
    smp = p.batch_first + (i - slot * p.batch_spp);
    a = make_float4(0.0f, 0.0f, 0.0f, 1.0f); b = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    s1 = 0;
    if (!wf_pixel(p.f, slot, px, py, oi)) return false;
    const uint32_t s0 = pcg((py * cfg.width + px) ^ pcg(cfg.seed + 0x9E3779B9u * cfg.frame_index));
    s1 = pcg(s0 + smp);
    float jx = 0.5f, jy = 0.5f;
    if (smp > 0) { jx = u01(pcg(s1 + 0u)); jy = u01(pcg(s1 + 1u)); }
    f3 ro, rd;
    camera_ray(p.f.cam, cfg.width, cfg.height, px, py, jx, jy, ro, rd);
    a = make_float4(ro.x, ro.y, ro.z, p.f.cam.near);
    b = make_float4(rd.x, rd.y, rd.z, p.f.cam.far);
    return true;
}

// materialises the depth-0 queue (only when that queue goes to the persistent tracer: DUST_WF_PT_DEPTH=0, A/B runs)
__global__ void __launch_bounds__(256) wf_gen_primary(const __grid_constant__ WfParams p)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= p.n_paths) return;
    float4 a, b;
    uint32_t s1, smp, px, py, oi;
    wf_primary_ray(p, i, a, b, s1, smp, px, py, oi);
    p.qc.b[i] = b;
    if (p.qc.kind == 2u) p.qc.a[i] = make_float4(a.x, a.y, a.z, __uint_as_float(i));     // compact closest queue (utmin = camera near)
    else { p.qc.a[i] = a; p.qc.path[i] = i; }
    if (i == 0) p.counts[3] = p.n_paths;                               // size of the depth-0 closest queue, for kernels that read it from memory
}

__device__ __forceinline__ void wf_store_hit(const WfParams &p, uint32_t i, bool hit, float t, uint32_t inst, uint32_t leaf, uint32_t pos, uint32_t voxel)
{
    if (p.hits_out) {                                                  // DustHit: t, instance, leaf, voxel (one 128-bit store)
        reinterpret_cast<uint4 *>(p.hits_out)[i] = hit ? make_uint4(__float_as_uint(t), inst, leaf, voxel)
                                                       : make_uint4(__float_as_uint(CUDART_INF_F), 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    } else {
        p.hits[i] = hit ? make_uint4(__float_as_uint(t), inst | (voxel << 20), leaf, pos)
                        : make_uint4(__float_as_uint(CUDART_INF_F), 0xFFFFFFFFu, 0xFFFFFFFFu, 0u);
    }
}

// One thread per ray of a queue; the same trace_scene as the megakernel.
// `list` != NULL: trace only the queue slots list[0 .. *count) (the rays wf_cull could not pre-cull).
// ROUGH (ray batches only, DUST_RAYS_ROUGH): block-granular leaves; a compile-time flag — as a run-time one the cold branch cost the
// hot path registers (-2.5 % on config 5).
template <int NL, bool STATS, int ROUGH = 0>          // ROUGH: the leaf mode (LEAF_PRECISE / LEAF_ROUGH / LEAF_AO)
__global__ void __launch_bounds__(DUST_THREADS, DUST_MIN_BLOCKS) wf_trace(const __grid_constant__ WfParams p, const WfQueue q, const uint32_t *count,
                                                                          const uint32_t *list)
{
    __shared__ uint32_t s_scratch[DUST_SCRATCH_WORDS * DUST_THREADS];
    const uint32_t n = count ? *count : p.n_paths;
    if (blockIdx.x * DUST_THREADS >= n) return;
    const uint32_t k = blockIdx.x * DUST_THREADS + threadIdx.x;
    const bool live = k < n;
    const uint32_t i = live && list ? __ldg(&list[k]) : k;
    float4 a = make_float4(0.0f, 0.0f, 0.0f, 1.0f), b = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    uint32_t pw = 0;
    if (live) {
        if (q.stride == 0u) { uint32_t s1, smp, px, py, oi; wf_primary_ray(p, i, a, b, s1, smp, px, py, oi); pw = i; }
        else wf_fetch(q, i, a, b, pw);
    }
    Counters st = {};
    Hit h;
    // implicit depth-0 queue: the 32 rays of a warp are camera rays of ONE tile (1024 * batch_spp consecutive slots per tile)
    const uint32_t *tl = nullptr;
    if (q.stride == 0u && p.f.tile_lists) {
        const uint32_t i0 = blockIdx.x * DUST_THREADS + (threadIdx.x & ~31u);
        tl = p.f.tile_lists + (size_t)min((i0 / p.batch_spp) >> 10, p.f.n_local_tiles - 1u) * DUST_TILE_LIST_WORDS;
    }
    // sun shadow rays (queue kind 1: every ray runs along sun_dir): the list of the sun-grid cell the origin projects into (per lane)
    else if (q.kind == 1u && p.f.sun_lists && live) tl = sun_list_of(p.f, mk3(a.x, a.y, a.z));
    trace_scene<NL, STATS>(p.f.scene, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), a.w, b.w, (pw >> 31) != 0u, live, s_scratch + threadIdx.x, h, st, ROUGH, tl);
    if (live) {
        if (STATS && q.stride == 2u) st.primary_rays++;
        if (pw >> 31) p.vis[pw & 0x7FFFFFFFu] = h.hit ? 1 : 0;
        else wf_store_hit(p, i, h.hit, h.t, h.inst, h.leaf, h.pos, h.voxel);
    }
    if (STATS) flush_counters(p.f.counters, st);
}


// ---------------------------------------------------------------- incoherent rays: pre-cull + persistent lanes
//
// wf_trace gives every ray a thread; a warp lasts as long as its longest ray.  For bounce rays that
// leaves 13 of 32 lanes idle (profiles/r01_history.md) — also after sorting them by direction bin:
// neighbouring rays start a few voxels apart and meet different fates.  For those queues the work is split:
//   wf_cull      one thread per ray, all lanes busy: walk the instance BVH, store the sorted candidate
//                keys of the ray (<= WF_CANDS; rays that touch more go to an overflow list for wf_trace)
//   wf_trace_pt  persistent lanes over the queue: NEXT / WALK / LEAF voted as in trace_scene, and a lane
//                whose ray is finished writes its result and claims the next ray of the queue (one
//                warp-aggregated atomic) inside the NEXT operation — no lane waits for another lane's ray.

// (Persistent lanes for this pass too — a lane whose BVH walk is finished writes its record and claims the next ray while the others go on
// expanding nodes — were built and measured: 1.02 ms with one refill per trip and one atomic per refill, 0.74 ms with 128-ray chunks per
// warp, 0.56 ms = no gain once the refill waits for 16 lanes.  A walk is 5 node visits; writing a record, fetching a ray, three IEEE
// reciprocals and the root box test are a fifth of it, and run with a handful of lanes they cost what the idle lanes cost here.)
template <bool STATS>
__global__ void __launch_bounds__(DUST_THREADS) wf_cull(const __grid_constant__ WfParams p, const WfQueue q, const uint32_t *count, uint32_t *overflow_count)
{
    __shared__ uint32_t s_scratch[DUST_SCRATCH_WORDS * DUST_THREADS];
    const unsigned full = 0xffffffffu;
    const uint32_t n = count ? *count : p.n_paths;
    if (blockIdx.x * DUST_THREADS >= n) return;
    const uint32_t i = blockIdx.x * DUST_THREADS + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const DSceneView &sc = p.f.scene;
    uint32_t *sm = s_scratch + threadIdx.x;
    const uint32_t id_mask = (1u << sc.id_bits) - 1u;
    Counters st = {};
    bool over = false;
    if (i < n) {
        float4 a, b;
        uint32_t pw_unused;
        wf_fetch(q, i, a, b, pw_unused);
        const f3 ro = mk3(a.x, a.y, a.z), rd = mk3(b.x, b.y, b.z);
        const float tmin = a.w, tmax = b.w;
        const f3 winv = cull_inverse(rd);
        int sp, pend, nc;
        bool trunc = false;
        const uint32_t *sl = q.kind == 1u && p.f.sun_lists ? sun_list_of(p.f, ro) : nullptr;
        if (sl && __ldg(&sl[0]) != 0xFFFFFFFFu) {
            // sun shadow rays: the candidates come from the sun grid (kernels.cuh sun_list_of), no BVH walk
            sp = DUST_SCRATCH_WORDS; pend = -1; nc = 0;
            cull_from_list(sc, sl, ro, winv, tmin, tmax, id_mask, sm, nc);
            if (nc > WF_CANDS) { nc = WF_CANDS; trunc = true; }         // sorted: the farthest ones go
        } else
        cull_begin(sc, ro, winv, tmin, tmax, id_mask, sm, sp, pend, nc);
        const CullRay cr = cull_ray(sc, ro, winv);
        while (pend >= 0 && cull_room(sp, nc)) {
            cull_expand<STATS>(sc, cr, tmin, tmax, id_mask, sm, sp, pend, nc, st);
            // The record holds WF_CANDS keys: drop the farthest ones as the walk goes (the list is sorted, and the last kept key only
            // moves nearer afterwards, so every dropped candidate starts at or behind it) — the scratch then always has room for
            // the BVH stack and the walk completes.
            if (nc > WF_CANDS) { nc = WF_CANDS; trunc = true; }
        }
        over = pend >= 0;                                             // BVH walk not finished (scratch full): slow path
        uint32_t c[WF_CANDS];
#pragma unroll
        for (int k = 0; k < WF_CANDS; k++) c[k] = (!over && k < nc) ? sm[k * DUST_THREADS] : 0xFFFFFFFFu;
        if (over) c[0] = 0xFFFFFFFEu;                                  // marker: not pre-culled
        // A truncated list is flagged in its last key (keys are non-negative floats: bit 31 is free).  The tracer finishes such a
        // ray itself when its best hit lies before that key's entry distance and only otherwise hands it to the slow path.
        else if (trunc) c[WF_CANDS - 1] |= 0x80000000u;
#pragma unroll
        for (int v = 0; v < WF_CVECS; v++) p.cand[WF_CVECS * (size_t)i + v] = make_uint4(c[4 * v], c[4 * v + 1], c[4 * v + 2], c[4 * v + 3]);
    }
    const unsigned m = __ballot_sync(full, over);
    if (m) {
        uint32_t base = 0;
        if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(overflow_count, (uint32_t)__popc(m));
        base = __shfl_sync(full, base, __ffs(m) - 1);
        if (over) p.overflow[base + (uint32_t)__popc(m & ((1u << lane) - 1u))] = i;
    }
    if (STATS) flush_counters(p.f.counters, st);
}

// ---- the slow path: ONE CTA PER RAY ------------------------------------------------------------------------------
// Rays whose candidate list was not conclusive (they graze dozens of instance boxes and hit nothing early; a few hundred per frame) used to be
// re-traced one thread per ray: a handful of warps, each walking 30+ instances in sequence, held the whole pass for 0.19 ms.
// Here the threads of a CTA take the INSTANCES of one ray in parallel: thread k enters instances k, k + 128, ... (walk_enter
// is the exact box clip, the BVH is only an accelerator), walks each to its closest hit, and the CTA reduces the
// lexicographic minimum of (t, instance) — the same result as any other visiting order.  Persistent grid: CTAs stride
// over the list, whose length lives on the device.
template <int NL, bool STATS, bool ROUGH = false>
__global__ void __launch_bounds__(DUST_THREADS) wf_trace_slow(const __grid_constant__ WfParams p, const WfQueue q, const uint32_t *count, const uint32_t *list)
{
    // ONE CTA PER RAY since v13 (a warp per ray before: 110 instances = 4 in sequence per lane, and the pass is nothing but that chain's latency)
    __shared__ float s_t[DUST_THREADS / 32];
    __shared__ uint32_t s_i[DUST_THREADS / 32];
    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const DSceneView &sc = p.f.scene;
    const uint32_t n = *count;
    Counters st = {};
    for (uint32_t k = blockIdx.x; k < n; k += gridDim.x) {
        const uint32_t i = __ldg(&list[k]);
        float4 a, b;
        uint32_t pw;
        wf_fetch(q, i, a, b, pw);
        const f3 ro = mk3(a.x, a.y, a.z), rd = mk3(b.x, b.y, b.z);
        const float tmin = a.w, tmax = b.w;
        const bool any = (pw >> 31) != 0u;
        bool hit = false;
        float bt = tmax;
        uint32_t b_inst = 0xFFFFFFFFu, b_leaf = 0, b_pos = 0, b_vox = 0;
        Walk w;
        w.four = NL == 4; w.have_mid = false; w.L4 = 0; w.blk_lo = w.blk_hi = 0;
        for (uint32_t inst = threadIdx.x; inst < sc.n_inst; inst += DUST_THREADS) {
            if (!walk_enter<NL, STATS>(sc, w, inst, ro, rd, tmin, bt, st)) continue;       // bt: this thread's own best so far (tmax shrinks per thread)
            for (;;) {
                const int r = walk_step<NL, STATS>(w, st);
                if (r == WS_OUT) break;
                if (r == WS_LEAF) {
                    float th; uint32_t vox;
                    if (walk_leaf<STATS>(w, tmin, th, vox, st, ROUGH) && tmin <= th && th <= bt) {  // reportIntersectionEXT; ascending inst: ties keep the earlier one
                        if (!hit || th < bt) { hit = true; bt = th; b_inst = inst; b_leaf = w.leaf_i; b_pos = walk_pos(w); b_vox = vox; }
                        break;                                                              // first accepted hit is this instance's closest
                    }
                    w.L4 = 1;
                }
            }
            if (any && hit) break;
        }
        // lexicographic minimum of (t, instance) over the threads: shuffles inside a warp, shared memory across the warps
        float rt = hit ? bt : CUDART_INF_F;
        uint32_t ri = hit ? b_inst : 0xFFFFFFFFu;
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) {
            const float ot = __shfl_xor_sync(full, rt, ofs);
            const uint32_t oi = __shfl_xor_sync(full, ri, ofs);
            if (ot < rt || (ot == rt && oi < ri)) { rt = ot; ri = oi; }
        }
        if (lane == 0) { s_t[wid] = rt; s_i[wid] = ri; }
        __syncthreads();
#pragma unroll
        for (int v = 0; v < DUST_THREADS / 32; v++) {
            const float ot = s_t[v];
            const uint32_t oi = s_i[v];
            if (ot < rt || (ot == rt && oi < ri)) { rt = ot; ri = oi; }
        }
        __syncthreads();                                                                    // s_t / s_i are rewritten by the CTA's next ray
        const bool any_hit = ri != 0xFFFFFFFFu;
        if (any) { if (threadIdx.x == 0) p.vis[pw & 0x7FFFFFFFu] = any_hit ? 1 : 0; }
        else if (!any_hit) { if (threadIdx.x == 0) wf_store_hit(p, i, false, tmax, 0xFFFFFFFFu, 0u, 0u, 0u); }
        else if (hit && b_inst == ri) wf_store_hit(p, i, true, bt, b_inst, b_leaf, b_pos, b_vox);        // exactly one thread owns the winner
        if (STATS && q.stride == 2u && threadIdx.x == 0) st.primary_rays++;
    }
    if (STATS) flush_counters(p.f.counters, st);
}

#ifndef DUST_PT_MIN_BLOCKS2
#define DUST_PT_MIN_BLOCKS2 8                 // 64 registers: 8 CTAs / SM beat 6 with no spills that matter (A/B in profiles/r01_history.md)
#endif
#ifndef DUST_PT_MIN_BLOCKS4
#define DUST_PT_MIN_BLOCKS4 10                // 4-level / mixed scenes: 48 registers; the 523 MB tree of configs[4] misses L2, warps in flight pay (7 / 8 / 10 CTAs: 2 120 / 2 125 / 2 199 Mrays/s)
#endif

#ifndef DUST_WALK_REPS3
#define DUST_WALK_REPS3 1                   // cells per voted WALK of the persistent tracer in 3-level scenes
#endif
enum : int { W_NEXT = 0, W_WALK = 1, W_LEAF = 2, W_DEAD = 3 };

// ---- TMA staging of the ray queue (BASELINE.json north_star: "stages ... through TMA into shared memory") --------------------------
// EXPERIMENT, compiled only with -DDUST_TMA_QUEUE (off in the shipped library): parity-green (the whole GPU suite passes with it),
// but 2.2 % slower on configs[2] and 6 % slower on configs[4] than per-lane __ldg + one atomic per trip (A/B in
// profiles/r02_history.md).  The tracer is bound by instruction issue on the ALU pipe: a claim's latency is already hidden by the
// other seven warps of the scheduler, while the mbarrier waits, the proxy fence and the slot arithmetic are instructions it
// has to issue.  A round-1 experiment that staged the root node words the same way (-DDUST_TMA_TOP) lost 5.6 %.
// The queues are the one thing on this path that STREAMS: every ray record is read once, in queue order.  Each warp of the
// persistent tracer therefore owns two half-buffers of WF_STAGE rays in shared memory.  One lane claims WF_STAGE consecutive
// queue slots with ONE atomic and issues bulk copies (cp.async.bulk global -> shared, completion on an mbarrier: SASS UBLKCP) of
// their origin / direction / path / candidate records; lanes whose ray is finished take the next staged ray with a ballot prefix —
// no per-lane global loads, no per-trip atomic — and the half that ran empty is refilled while the other one is consumed.
// (The node words of the walk are NOT staged: they are data-dependent single 16-byte reads that hit L1; measured in round 1.)
#ifndef WF_STAGE
#define WF_STAGE 16                   // rays per half-buffer (multiple of 4: the path words of a copy are a multiple of 16 bytes)
#endif
#define WF_STAGE_BYTES (WF_STAGE * (16 + 16 + 4 * WF_CANDS + 4))     // a | b | cand | path of one half-buffer

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0u;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Claim the next WF_STAGE queue slots and start their copies into half-buffer `half` of this warp (called by ONE lane).
// stage layout: [a: WF_STAGE x 16 | b: WF_STAGE x 16 | cand: WF_STAGE x 4 WF_CANDS | path: WF_STAGE x 4]; interleaved DustRay records
// (stride 2) land in the a|b region as they are.  `meta` = {first queue slot, rays} of the half, read by the consumers after the wait.
template <bool CAND>
__device__ __forceinline__ void wf_stage_refill(const WfParams &p, const WfQueue &q, uint32_t n, unsigned long long *work,
                                                uint8_t *stage, uint32_t bar, uint2 *meta)
{
    const unsigned long long base = atomicAdd(work, (unsigned long long)WF_STAGE);
    const uint32_t c = base >= n ? 0u : min((uint32_t)WF_STAGE, n - (uint32_t)base);
    *meta = make_uint2((uint32_t)base, c);
    if (c == 0u) { mbar_arrive(bar); return; }                          // queue exhausted: the phase completes with nothing in flight
    const uint32_t b32 = (uint32_t)base, dst = smem_u32(stage);
    const uint32_t c4 = (c + 3u) & ~3u;                                 // path words: whole 16-byte units (the arrays are padded to 256 bytes)
    uint32_t bytes = c * 32u;
    if (CAND) bytes += c * (4u * WF_CANDS);
    if (q.path) bytes += c4 * 4u;
    mbar_expect_tx(bar, bytes);
    if (q.stride == 2u) bulk_g2s(dst, q.a + 2u * (size_t)b32, c * 32u, bar);
    else {
        bulk_g2s(dst, q.a + b32, c * 16u, bar);
        bulk_g2s(dst + WF_STAGE * 16u, q.b + b32, c * 16u, bar);
    }
    if (CAND) bulk_g2s(dst + WF_STAGE * 32u, p.cand + WF_CVECS * (size_t)b32, c * (4u * WF_CANDS), bar);
    if (q.path) bulk_g2s(dst + WF_STAGE * (32u + 4u * WF_CANDS), q.path + b32, c4 * 4u, bar);
}
#ifndef DUST_FAST_VOTE_LANES
#define DUST_FAST_VOTE_LANES (NL == 3 ? 12u : 16u)   // 256^3 models: 8 / 12 / 16 lanes = 2 946 / 2 996 / 2 971 Mrays/s on configs[2]; the 4096^3 domain loses 1.5 % at 12
#endif

// CAND: the queue comes with pre-culled candidate records (p.cand, multi-instance scenes); without them the only candidate is the
// scene box and the truncated-list logic compiles away.
template <int NL, bool STATS, bool CAND, bool ROUGH = false, int KIND = 0>     // KIND: WfQueue::kind of the queue (compile-time: the claim is hot)
__global__ void __launch_bounds__(DUST_THREADS, NL == 3 ? DUST_PT_MIN_BLOCKS2 : DUST_PT_MIN_BLOCKS4) wf_trace_pt(const __grid_constant__ WfParams p, const WfQueue q, const uint32_t *count, unsigned long long *work,
                                                                                                                  uint32_t *overflow_count)
{
    __shared__ uint32_t s_cand[WF_CANDS * DUST_THREADS];
    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const DSceneView &sc = p.f.scene;
    uint32_t *sm = s_cand + threadIdx.x;
    const uint32_t id_mask = (1u << sc.id_bits) - 1u;
    const uint32_t n = count ? *count : p.n_paths;
    Counters st = {};

    // the ray in flight and its best hit
    bool have_ray = false, hit = false;
    uint32_t ri = 0, pw = 0;
    f3 ro = mk3(0.0f, 0.0f, 0.0f), rd = mk3(1.0f, 1.0f, 1.0f);
    float tmin = 0.0f, cur = 0.0f;
    uint32_t b_inst = 0xFFFFFFFFu, b_leaf = 0, b_pos = 0, b_vox = 0;
    int mode = W_NEXT, j = 0;                                          // j == WF_CANDS or key 0xFFFFFFFF: list exhausted
    bool exhausted = true;
    float e_trunc = CUDART_INF_F;                                      // truncated candidate list: entry distance of its last key (+inf: complete list)
    Walk w;
    w.four = NL == 4; w.have_mid = false; w.L4 = 0; w.blk_lo = w.blk_hi = 0;

#ifdef DUST_PAD_SMEM
    __shared__ uint8_t s_pad[DUST_PAD_SMEM];                           // A/B only: what does the L1 / shared carve-out cost?
    if (n == 0xFFFFFFFFu) s_pad[threadIdx.x] = 1;
    if (n == 0xFFFFFFFEu && s_pad[threadIdx.x ^ 1]) return;
#endif
#ifdef DUST_TMA_QUEUE
    // per warp: two half-buffers of staged rays, their mbarriers and {first slot, count} words; `taken` = staged slots this warp
    // has handed out so far (slot g lives in half (g / WF_STAGE) & 1, generation g / (2 WF_STAGE))
    __shared__ __align__(128) uint8_t s_stage[DUST_THREADS / 32][2][WF_STAGE_BYTES];
    __shared__ __align__(8) unsigned long long s_bar[DUST_THREADS / 32][2];
    __shared__ uint2 s_meta[DUST_THREADS / 32][2];
    const unsigned wid = threadIdx.x >> 5;
    uint32_t taken = 0;
    if (lane == 0) {
        mbar_init(smem_u32(&s_bar[wid][0]), 1u); mbar_init(smem_u32(&s_bar[wid][1]), 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        wf_stage_refill<CAND>(p, q, n, work, s_stage[wid][0], smem_u32(&s_bar[wid][0]), &s_meta[wid][0]);
        wf_stage_refill<CAND>(p, q, n, work, s_stage[wid][1], smem_u32(&s_bar[wid][1]), &s_meta[wid][1]);
    }
    __syncwarp();
#endif

#ifdef DUST_TMA_TOP
    // Experiment (profiles/r01_history.md): single-instance scenes stage the root 16^3 node (64 x 16 B) and its
    // block mask in shared memory with one TMA bulk copy per CTA; the walk then reads them with generic loads.
    __shared__ __align__(128) uint4 s_top[64];
    __shared__ __align__(16) uint2 s_topb[2];
    __shared__ __align__(8) unsigned long long s_bar;
    const bool staged = sc.n_inst == 1;
    if (staged) {
        const DInstance &in0 = sc.inst[0];
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(1024u + 16u));
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"((uint32_t)__cvta_generic_to_shared(s_top)), "l"(in0.n16), "r"(1024u), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"((uint32_t)__cvta_generic_to_shared(s_topb)), "l"(in0.blocks), "r"(16u), "r"(bar) : "memory");
        }
        __syncthreads();
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar) : "memory");
    }
#endif

#pragma unroll 1
    for (;;) {
        int op = W_WALK;
        uint32_t n_next = 0, n_walk = 0, n_leaf = 0, cnt = 0;
#ifndef DUST_NO_FAST_VOTE
        // fast path: half of the warp or more waits for WALK — it is the plurality whatever the others want (one ballot + popc
        // instead of the REDUX and its decoding: the vote was 12 % of this kernel's instructions, profiles/r01_v8_*)
        n_walk = (uint32_t)__popc(__ballot_sync(full, mode == W_WALK));
        if (n_walk < DUST_FAST_VOTE_LANES)
#endif
        {
            const uint32_t packed = __reduce_add_sync(full, mode < W_DEAD ? (1u << (8 * mode)) : 0u);
            if (packed == 0u) break;
            n_next = packed & 255u; n_walk = (packed >> 8) & 255u; n_leaf = packed >> 16;
            cnt = n_walk;
            if (n_next > cnt) { op = W_NEXT; cnt = n_next; }
            if (n_leaf > cnt) { op = W_LEAF; cnt = n_leaf; }
        }
        if (STATS && lane == 0) {
            if (op == W_WALK) cnt = n_walk;
            st.dbg[op == W_WALK ? 0 : (op == W_LEAF ? 1 : 2)]++;
            st.dbg[op == W_WALK ? 3 : (op == W_LEAF ? 4 : 5)] += cnt;
        }

        if (op == W_WALK) {
#pragma unroll 1
            // cells per vote: two for the long empty runs of the 4096^3 domain (one: -5 % on config 5), one for 256^3 models, where
            // a second cell serves too few lanes (+0.9 % on config 2)
            for (int rep = 0; rep < (NL == 3 ? DUST_WALK_REPS3 : DUST_WALK_REPS); rep++)
                if (mode == W_WALK) {
                    const int r = walk_step<NL, STATS>(w, st);
                    mode = r == WS_LEAF ? W_LEAF : (r == WS_OUT ? W_NEXT : W_WALK);
                }
        } else if (op == W_LEAF) {
            if (mode == W_LEAF) {
                float th; uint32_t vox;
                if (walk_leaf<STATS>(w, tmin, th, vox, st, ROUGH) && tmin <= th && th <= cur) {   // reportIntersectionEXT
                    if (pw >> 31) { hit = true; exhausted = true; }                    // any hit ends a shadow ray
                    else if (!hit || th < cur || w.inst < b_inst) {                    // (t, instance) lexicographic
                        hit = true; cur = th; b_inst = w.inst; b_leaf = w.leaf_i; b_pos = walk_pos(w); b_vox = vox;
                    }
                    mode = W_NEXT;                                                     // first accepted hit is this instance's closest
                } else { mode = W_WALK; w.L4 = 1; }
            }
        } else {
            bool need = false;
            if (mode == W_NEXT) {
                bool entered = false;
                if (!exhausted) {
                    const uint32_t key = sm[j * DUST_THREADS];
                    ++j;
                    if (j == WF_CANDS) exhausted = true;
                    if (key == 0xFFFFFFFFu || __uint_as_float(key & ~id_mask) > cur) exhausted = true;   // end of list / sorted: nothing further can beat the best hit
                    else if (walk_enter<NL, STATS>(sc, w, key & id_mask, ro, rd, tmin, cur, st)) {
                        mode = W_WALK; entered = true;
#ifdef DUST_TMA_TOP
                        if (staged) { w.topw = s_top; w.topb = s_topb; }
#endif
                    }
                }
                need = !entered && exhausted;
            }
            // lanes whose ray is finished: write the result, claim the next ray of the queue (one atomic per warp)
            const unsigned want = __ballot_sync(full, need);
            if (want) {
                if (need && have_ray) {
                    // a truncated list is conclusive only if the dropped candidates (all starting at or behind e_trunc) cannot change the result
                    const bool conclusive = !CAND || e_trunc == CUDART_INF_F || (hit && ((pw >> 31) != 0u || cur < e_trunc));
                    if (!conclusive) p.overflow[atomicAdd(overflow_count, 1u)] = ri;             // rare (hundreds per frame): wf_trace_slow redoes the ray
                    else {
                        if (STATS && q.stride == 2u) st.primary_rays++;
                        if (pw >> 31) p.vis[pw & 0x7FFFFFFFu] = hit ? 1 : 0;
                        else wf_store_hit(p, ri, hit, cur, b_inst, b_leaf, b_pos, b_vox);
                    }
                }
#ifdef DUST_TMA_QUEUE
                // hand staged rays out: the lanes that need one take consecutive staged slots (ballot prefix) — from the current
                // half-buffer and, past its end, from the other one (lanes beyond both are served by the next NEXT operation);
                // a half that ran empty is refilled at once (one atomic + bulk copies) while the other half is consumed
                if (need) have_ray = false;
                const uint32_t pos = taken % WF_STAGE, half0 = (taken / WF_STAGE) & 1u;
                const uint32_t rank = (uint32_t)__popc(want & ((1u << lane) - 1u));
                const uint32_t take = min((uint32_t)__popc(want), 2u * WF_STAGE - pos);
                const bool second = rank + pos >= (uint32_t)WF_STAGE;                 // this lane's slot lies in the other half
                const uint32_t half = half0 ^ (second ? 1u : 0u), sl = second ? rank + pos - WF_STAGE : rank + pos;
                // generations: slot g = taken + rank lives in generation g / (2 WF_STAGE); both copies were issued long ago
                while (!mbar_try_wait(smem_u32(&s_bar[wid][half0]), (taken / (2u * WF_STAGE)) & 1u)) { }
                if (pos + take > (uint32_t)WF_STAGE)
                    while (!mbar_try_wait(smem_u32(&s_bar[wid][half0 ^ 1u]), ((taken + WF_STAGE) / (2u * WF_STAGE)) & 1u)) { }
                if (need && rank < take) {
                    const uint2 meta = s_meta[wid][half];
                    if (sl >= meta.y) mode = W_DEAD;                                 // past the end of the queue
                    else {
                        const uint8_t *st_ = s_stage[wid][half];
                        ri = meta.x + sl;
                        float4 a, b;
                        if (q.stride == 2u) { a = *reinterpret_cast<const float4 *>(st_ + sl * 32u); b = *reinterpret_cast<const float4 *>(st_ + sl * 32u + 16u); }
                        else { a = *reinterpret_cast<const float4 *>(st_ + sl * 16u); b = *reinterpret_cast<const float4 *>(st_ + WF_STAGE * 16u + sl * 16u); }
                        uint32_t first_key;
                        if (CAND) {
                            e_trunc = CUDART_INF_F;
#pragma unroll
                            for (int v = 0; v < WF_CVECS; v++) {
                                uint4 c = *reinterpret_cast<const uint4 *>(st_ + WF_STAGE * 32u + sl * (4u * WF_CANDS) + 16u * v);
                                if (v == 0) first_key = c.x;
                                if (v == WF_CVECS - 1 && (c.w >> 31) && c.w < 0xFFFFFFFEu) { c.w &= 0x7FFFFFFFu; e_trunc = __uint_as_float(c.w & ~id_mask); }
                                sm[(4 * v + 0) * DUST_THREADS] = c.x; sm[(4 * v + 1) * DUST_THREADS] = c.y;
                                sm[(4 * v + 2) * DUST_THREADS] = c.z; sm[(4 * v + 3) * DUST_THREADS] = c.w;
                            }
                        } else {
                            // single-instance scene (no wf_cull pass): the candidate list is the scene box test
                            const float te = sc.n_inst ? box_entry(sc.root_lo[0], sc.root_lo[1], sc.root_lo[2], sc.root_hi[0], sc.root_hi[1], sc.root_hi[2],
                                                                   mk3(a.x, a.y, a.z), cull_inverse(mk3(b.x, b.y, b.z)), a.w, b.w) : CUDART_INF_F;
                            first_key = te != CUDART_INF_F ? cand_key(te, 0u, id_mask) : 0xFFFFFFFFu;
                            sm[0] = first_key; sm[DUST_THREADS] = 0xFFFFFFFFu;
                        }
                        if (first_key != 0xFFFFFFFEu) {                                // else: wf_cull could not pre-cull it, wf_trace_slow does it
                            pw = q.path ? *reinterpret_cast<const uint32_t *>(st_ + WF_STAGE * (32u + 4u * WF_CANDS) + sl * 4u) : ri;
                            ro = mk3(a.x, a.y, a.z); tmin = a.w; rd = mk3(b.x, b.y, b.z); cur = b.w;
                            hit = false; b_inst = 0xFFFFFFFFu;
                            j = 0; exhausted = false;
                            have_ray = true;
                        }
                    }
                }
                taken += take;
                if (pos + take >= (uint32_t)WF_STAGE) {                               // half0 (and maybe the other half too) is used up: refill
                    __syncwarp();
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the reads above before the async-proxy writes below
                    if (lane == 0) {
                        wf_stage_refill<CAND>(p, q, n, work, s_stage[wid][half0], smem_u32(&s_bar[wid][half0]), &s_meta[wid][half0]);
                        if (pos + take == 2u * WF_STAGE)
                            wf_stage_refill<CAND>(p, q, n, work, s_stage[wid][half0 ^ 1u], smem_u32(&s_bar[wid][half0 ^ 1u]), &s_meta[wid][half0 ^ 1u]);
                    }
                    __syncwarp();
                }
#else
                const int leader = __ffs(want) - 1;
                unsigned long long base = 0;
                if ((int)lane == leader) base = atomicAdd(work, (unsigned long long)__popc(want));
                base = __shfl_sync(full, base, leader);
                if (need) {
                    const unsigned long long idx = base + (unsigned)__popc(want & ((1u << lane) - 1u));
                    have_ray = false;
                    if (idx >= n) mode = W_DEAD;
                    else {
                        ri = (uint32_t)idx;
                        float4 a, b;
                        uint32_t pw_new;
                        wf_fetch<KIND>(q, ri, a, b, pw_new);
                        uint4 c0, cv[WF_CVECS];
#pragma unroll
                        for (int v = 0; v < WF_CVECS; v++) cv[v] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
                        if (CAND) {
#pragma unroll
                            for (int v = 0; v < WF_CVECS; v++) cv[v] = __ldg(&p.cand[WF_CVECS * (size_t)ri + v]);
                            c0 = cv[0];
                            uint32_t &last = cv[WF_CVECS - 1].w;
                            e_trunc = CUDART_INF_F;
                            if ((last >> 31) && last < 0xFFFFFFFEu) { last &= 0x7FFFFFFFu; e_trunc = __uint_as_float(last & ~id_mask); }
                        } else {
                            // single-instance scene (no wf_cull pass): the candidate list is the scene box test
                            const float te = sc.n_inst ? box_entry(sc.root_lo[0], sc.root_lo[1], sc.root_lo[2], sc.root_hi[0], sc.root_hi[1], sc.root_hi[2],
                                                                   mk3(a.x, a.y, a.z), cull_inverse(mk3(b.x, b.y, b.z)), a.w, b.w) : CUDART_INF_F;
                            c0 = make_uint4(te != CUDART_INF_F ? cand_key(te, 0u, id_mask) : 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
                            cv[0] = c0;
                        }
                        if (c0.x != 0xFFFFFFFEu) {                                     // else: overflow ray, wf_trace does it
                            pw = pw_new;
                            ro = mk3(a.x, a.y, a.z); tmin = a.w; rd = mk3(b.x, b.y, b.z); cur = b.w;
#pragma unroll
                            for (int v = 0; v < WF_CVECS; v++) {
                                sm[(4 * v + 0) * DUST_THREADS] = cv[v].x; sm[(4 * v + 1) * DUST_THREADS] = cv[v].y;
                                sm[(4 * v + 2) * DUST_THREADS] = cv[v].z; sm[(4 * v + 3) * DUST_THREADS] = cv[v].w;
                            }
                            hit = false; b_inst = 0xFFFFFFFFu;
                            j = 0; exhausted = false;
                            have_ray = true;
                        }
                    }
                }
#endif
            }
        }
    }
    if (STATS) flush_counters(p.f.counters, st);
}

// One thread per ray of the closest queue of depth p.depth: apply the pending sun term, consume the
// hit, emit the vertex's shadow and bounce rays.  The per-path arithmetic is frame_kernel's, in order.
#ifndef DUST_SHADE_MIN_BLOCKS
#define DUST_SHADE_MIN_BLOCKS 10              // 48 registers: the kernel waits for its dependent loads (hit -> instance -> leaf -> material -> palette), so
#endif                                        // warps in flight beat registers: 6 / 8 / 10 / 12 CTAs per SM = 2 775 / 2 772 / 2 804 / 2 802 Mrays/s on configs[2]
// All threads of a CTA of DUST_THREADS call this together (the emitted rays are compacted with warp votes and one
// reservation per CTA); threads without a ray pass live = false.  `hw_in` != NULL: the hit record of slot i is passed in registers instead of read from p.hits
// (the fused depth-0 kernel below).
template <bool STATS>
__device__ __forceinline__ void wf_shade_vertex(const WfParams &p, const uint32_t i, const bool live, const uint4 *hw_in, Counters &st)
{
    const unsigned full = 0xffffffffu;
    const DustTraceConfig &cfg = p.f.cfg;
    const unsigned lane = threadIdx.x & 31u;
    uint32_t path = 0;
    bool emit_shadow = false, emit_bounce = false;
    f3 ro = mk3(0.0f, 0.0f, 0.0f), rd = mk3(0.0f, 0.0f, 0.0f);
    if (live) {
        float4 a, b;
        f3 T = mk3(1.0f, 1.0f, 1.0f), L = mk3(0.0f, 0.0f, 0.0f);
        uint32_t s1 = 0, px = 0, py = 0, oi = 0, smp = 1;
        bool valid = true;
        if (p.depth == 0 && p.qc.stride == 0u) {                        // the depth-0 queue is implicit: path i, camera ray recomputed
            path = i;
            valid = wf_primary_ray(p, i, a, b, s1, smp, px, py, oi);
        } else {
            wf_fetch(p.qc, i, a, b, path);
            if (p.depth == 0) { float4 a2, b2; valid = wf_primary_ray(p, path, a2, b2, s1, smp, px, py, oi); }
            else {
                const float4 s0 = p.ps0[path], s1v = p.ps1[path];
                T = mk3(s0.x, s0.y, s0.z); s1 = __float_as_uint(s0.w);
                L = mk3(s1v.x, s1v.y, s1v.z);
                if (s1v.w > 0.0f && !p.vis[path]) {                      // nee.rmiss:11-22: the previous vertex's sun term, with that vertex's throughput
                    const f3 ts = acescg2srgb(mk3(p.f.sun_nee[0] * s1v.w, p.f.sun_nee[1] * s1v.w, p.f.sun_nee[2] * s1v.w));
                    L.x += T.x * ts.x; L.y += T.y * ts.y; L.z += T.z * ts.z;
                }
            }
        }
        const uint4 hw = hw_in ? *hw_in : __ldg(&p.hits[i]);
        ro = mk3(a.x, a.y, a.z); rd = mk3(b.x, b.y, b.z);
        float pend_ndl = 0.0f;                                           // n.l of this vertex's sun term while its shadow ray is in flight
        if (valid) {
            const bool hit = hw.y != 0xFFFFFFFFu;
            if (STATS) { if (p.depth == 0) st.primary_rays++; else st.bounce_rays++; }
            // uniform tests first, the per-lane `smp == 0` innermost: as one short-circuit chain ptxas split the warp at smp == 0 and
            // ran everything below twice, with 8 and with 24 lanes (profiles/r02_history.md 9)
            const bool hits_on = p.depth == 0 && (cfg.flags & DUST_TRACE_WRITE_HITS) && p.f.hit_instance;
            const bool gb_on = p.depth == 0 && gbuffer_wanted(p.f.gb);
            if (hits_on && smp == 0) {
                const uint32_t pi = py * cfg.width + px;
                p.f.hit_instance[pi] = hit ? (hw.y & 0xFFFFFu) : 0xFFFFFFFFu; p.f.hit_leaf[pi] = hit ? hw.z : 0xFFFFFFFFu;
                p.f.hit_voxel[pi] = hit ? (uint8_t)(hw.y >> 20) : (uint8_t)0xFF; p.f.hit_t[pi] = hit ? __uint_as_float(hw.x) : CUDART_INF_F;
            }
            const bool want_gb = gb_on && smp == 0;
            if (want_gb && !hit) gbuffer_store(p.f.gb, p.f.scene, py * cfg.width + px, false, 0.0f, 0u, 0u, 0u, ro, rd, ro);
            if (hit) {
                Hit h;
                h.hit = true; h.t = __uint_as_float(hw.x); h.inst = hw.y & 0xFFFFFu; h.voxel = hw.y >> 20; h.leaf = hw.z; h.pos = hw.w;
                f3 sn;
                const f3 ro0 = ro;
                uint32_t pidx = 0;
                const f3 al = surface_at<STATS>(p.f.scene, h, ro, rd, sn, st, &pidx);   // ro becomes the offset hit point
                if (want_gb) gbuffer_store(p.f.gb, p.f.scene, py * cfg.width + px, true, h.t, h.inst, h.voxel, pidx, ro0, rd, sn);
                if (p.depth == 0) { if (!(cfg.flags & DUST_TRACE_DEMODULATE)) T = al; }   // tone_map.comp:195-203
                else { T.x = T.x * al.x; T.y = T.y * al.y; T.z = T.z * al.z; }          // final_gather.rchit:76-80
                if (dot3(mk3(p.f.sky.direction[0], p.f.sky.direction[1], p.f.sky.direction[2]), sn) > 0.0f) {   // ambient_occlusion.rgen:33-50
                    if (STATS) st.shadow_rays++;
                    pend_ndl = dot3(sn, mk3(p.f.sun_dir[0], p.f.sun_dir[1], p.f.sun_dir[2]));
                    emit_shadow = true;
                }
                if (p.depth < cfg.max_bounces) {
                    const uint32_t v = p.depth;
                    const float u1 = u01(pcg(s1 + 2u + 2u * v)), u2 = u01(pcg(s1 + 3u + 2u * v));   // final_gather.rgen:32-34
                    const float rr = sqrtf(u1), cz = sqrtf(1.0f - u1);
                    float c, s;
                    cossin_2pi(u2, c, s);
                    rd = normalize3(rotate_by_normal(sn, mk3(rr * c, rr * s, cz)));
                    emit_bounce = true;
                }
            } else {
                const f3 dn = normalize3(rd);
                f3 sk, su;
                sky_eval(p.f.sky, p.f.sun_ar2, dn.x, dn.y, dn.z, p.depth == 0, sk.x, sk.y, sk.z, su.x, su.y, su.z);
                if (p.depth == 0) {                                                     // miss.rmiss:8-14
                    L = acescg2srgb(mk3((sk.x + su.x) / 3.14f, (sk.y + su.y) / 3.14f, (sk.z + su.z) / 3.14f));
                } else {                                                                // final_gather.rmiss:13-17
                    const f3 ts = acescg2srgb(sk);
                    L.x += T.x * ts.x; L.y += T.y * ts.y; L.z += T.z * ts.z;
                }
            }
        }
        p.ps0[path] = make_float4(T.x, T.y, T.z, __uint_as_float(s1));
        p.ps1[path] = make_float4(L.x, L.y, L.z, pend_ndl);
    }
    // ---- compact the emitted rays into the queues: one atomic per CTA and queue.  (One per warp was 518 400 atomics on two
    // words of one L2 sector per depth-0 launch: 56 % of the kernel's stall samples sat on them, profiles/r02_history.md 9.)
    // Every thread of the CTA gets here (callers exit CTA-uniformly), so the two barriers are safe.
    {
        __shared__ uint32_t s_cnt[2][DUST_THREADS / 32], s_base[2];
        const unsigned w = threadIdx.x >> 5;
        const unsigned ms = __ballot_sync(full, emit_shadow), mb = __ballot_sync(full, emit_bounce);
        if (lane == 0) { s_cnt[0][w] = (uint32_t)__popc(ms); s_cnt[1][w] = (uint32_t)__popc(mb); }
        __syncthreads();
        if (lane == 0 && w < 2u) {                                      // warp 0 reserves the shadow records, warp 1 the bounce records
            uint32_t tot = 0;
#pragma unroll
            for (int k = 0; k < DUST_THREADS / 32; k++) tot += s_cnt[w][k];
            s_base[w] = tot ? atomicAdd(&p.counts[4 * p.depth + 1 + w], tot) : 0u;
        }
        __syncthreads();
        uint32_t pre_s = s_base[0], pre_b = s_base[1];
#pragma unroll
        for (int k = 0; k < DUST_THREADS / 32; k++) if ((unsigned)k < w) { pre_s += s_cnt[0][k]; pre_b += s_cnt[1][k]; }
        const unsigned below = (1u << lane) - 1u;
        if (emit_shadow) {
            const uint32_t k = pre_s + (uint32_t)__popc(ms & below);
            if (p.qs.kind == 1u) p.qs.a[k] = make_float4(ro.x, ro.y, ro.z, __uint_as_float(path | 0x80000000u));   // direction, tmin, tmax: uniforms of the queue
            else {
                p.qs.a[k] = make_float4(ro.x, ro.y, ro.z, 0.1f);
                p.qs.b[k] = make_float4(p.f.sun_dir[0], p.f.sun_dir[1], p.f.sun_dir[2], 10000.0f);
                p.qs.path[k] = path | 0x80000000u;
            }
        }
        if (emit_bounce) {
            const uint32_t k = pre_b + (uint32_t)__popc(mb & below);
            p.qn.b[k] = make_float4(rd.x, rd.y, rd.z, p.f.cam.far);
            if (p.qn.kind == 2u) p.qn.a[k] = make_float4(ro.x, ro.y, ro.z, __uint_as_float(path));                 // tmin: uniform of the queue
            else { p.qn.a[k] = make_float4(ro.x, ro.y, ro.z, 0.1f); p.qn.path[k] = path; }
        }
    }
}

template <bool STATS>
__global__ void __launch_bounds__(DUST_THREADS, DUST_SHADE_MIN_BLOCKS) wf_shade(const __grid_constant__ WfParams p)
{
    const uint32_t n = wf_in_count(p);
    if (blockIdx.x * DUST_THREADS >= n) return;
    const uint32_t i = blockIdx.x * DUST_THREADS + threadIdx.x;
    Counters st = {};
    // (Dealing the CTA's bounce rays out again, hits to the low threads and misses behind them, so that a warp runs either the sky
    // path or the surface path, was tried: two barriers and 24 more bytes of spills cost what the shorter paths saved — 2 788 against 2 800 Mrays/s.)
    wf_shade_vertex<STATS>(p, i, i < n, nullptr, st);
    if (STATS) flush_counters(p.f.counters, st);
}

// Experiment (DUST_WF_FUSE0=1; off by default): depth 0 of a frame in ONE kernel — the camera ray of path i is traced and its
// vertex shaded at once, the hit record stays in registers.  Same arithmetic as wf_trace followed by wf_shade, bit-identical
// frames, but 2.8 % slower on config 2: the latency-bound shading then runs at the tracer's occupancy (6 CTAs/SM, 80 registers)
// instead of 8 (profiles/r01_history.md).
template <int NL, bool STATS>
__global__ void __launch_bounds__(DUST_THREADS, DUST_MIN_BLOCKS) wf_trace_shade0(const __grid_constant__ WfParams p)
{
    __shared__ uint32_t s_scratch[DUST_SCRATCH_WORDS * DUST_THREADS];
    const uint32_t n = p.n_paths;
    if (blockIdx.x * DUST_THREADS >= n) return;
    const uint32_t i = blockIdx.x * DUST_THREADS + threadIdx.x;
    const bool live = i < n;
    float4 a = make_float4(0.0f, 0.0f, 0.0f, 1.0f), b = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    if (live) { uint32_t s1, smp, px, py, oi; wf_primary_ray(p, i, a, b, s1, smp, px, py, oi); }
    Counters st = {};
    Hit h;
    trace_scene<NL, STATS>(p.f.scene, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), a.w, b.w, false, live, s_scratch + threadIdx.x, h, st);
    const uint4 hw = h.hit ? make_uint4(__float_as_uint(h.t), h.inst | (h.voxel << 20), h.leaf, h.pos)
                           : make_uint4(__float_as_uint(CUDART_INF_F), 0xFFFFFFFFu, 0xFFFFFFFFu, 0u);
    wf_shade_vertex<STATS>(p, i, live, &hw, st);
    if (STATS) flush_counters(p.f.counters, st);
}

// One thread per pixel slot: last pending sun term, sRGB -> ACEScg, sum in sample order (tone_map.comp:192-203 input)
__global__ void __launch_bounds__(256) wf_resolve(const __grid_constant__ WfParams p)
{
    const uint32_t slot = blockIdx.x * 256u + threadIdx.x;
    if (slot >= p.n_slots) return;
    uint32_t px, py, oi;
    if (!wf_pixel(p.f, slot, px, py, oi)) return;
    f3 acc = mk3(0.0f, 0.0f, 0.0f);
    if (p.batch_first) { const float4 a = p.acc[slot]; acc = mk3(a.x, a.y, a.z); }
    for (uint32_t s = 0; s < p.batch_spp; s++) {
        const uint32_t path = slot * p.batch_spp + s;
        const float4 s1v = p.ps1[path];
        f3 L = mk3(s1v.x, s1v.y, s1v.z);
        if (s1v.w > 0.0f && !p.vis[path]) {
            const float4 s0 = p.ps0[path];
            const f3 ts = acescg2srgb(mk3(p.f.sun_nee[0] * s1v.w, p.f.sun_nee[1] * s1v.w, p.f.sun_nee[2] * s1v.w));
            L.x += s0.x * ts.x; L.y += s0.y * ts.y; L.z += s0.z * ts.z;
        }
        const f3 Lc = srgb2acescg(L);
        acc.x += Lc.x; acc.y += Lc.y; acc.z += Lc.z;
    }
    if (!p.last_batch) { p.acc[slot] = make_float4(acc.x, acc.y, acc.z, 0.0f); return; }
    const float n = (float)p.f.cfg.spp;
    store_radiance(p.f.radiance32, p.f.radiance16, p.f.ycocg, oi, acc.x / n, acc.y / n, acc.z / n);
}

}  // namespace dust
