// C ABI of libdust_b200.so, part 2: the radiance cache and dust_trace_gi (include/dust_b200.h, "the reference's own global
// illumination").  Kernels: gi.cuh + wf_trace (wavefront.cuh).  The two key sorts and the prefix sum of the deterministic insert
// are CUB's device-wide primitives (part of the CUDA toolkit): plumbing around the kernels, 345 600 keys per frame.
// Included at the end of dust_abi.cu (one translation unit: the kernels of the shared headers are defined once).
#include "abi_internal.h"
#include "gi.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>

using namespace dust;

struct DustRadianceCache {
    DCacheEntry *table = nullptr;      // capacity + 2 entries
    float4 *pool = nullptr;
    uint32_t capacity = 0, pool_size = 0;
    float *srgb_lut = nullptr;         // 1024 floats
    unsigned long long *stats = nullptr;
    // per-frame scratch (queues, records, images, sort buffers), grown on demand
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
};


namespace {
inline float srgb_to_linear_host(float c) { return c < 0.04045f ? c / 12.92f : powf(fabsf(c + 0.055f) / 1.055f, 2.4f); }   // color.glsl:1-5
}

extern "C" {

int32_t dust_cache_create(DustCtx *ctx, uint32_t capacity, uint32_t surfel_pool_size, DustRadianceCache **out)
{
    if (!ctx || !out || !capacity || !surfel_pool_size) return DUST_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (capacity > 0xFFFFFFF0u || surfel_pool_size > (1u << 30)) return dust_fail(ctx, DUST_ERR_INVALID_ARGUMENT, "cache capacity / pool size too large");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    DustRadianceCache *c = new (std::nothrow) DustRadianceCache();
    if (!c) return DUST_ERR_OUT_OF_MEMORY;
    c->capacity = capacity; c->pool_size = surfel_pool_size;
    float lut[1024];
    for (int v = 0; v < 1024; v++) lut[v] = srgb_to_linear_host((float)v / 1023.0f);
    cudaError_t e = cudaMalloc(&c->table, ((size_t)capacity + 2) * sizeof(DCacheEntry));
    if (e == cudaSuccess) e = cudaMalloc(&c->pool, (size_t)surfel_pool_size * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc(&c->srgb_lut, sizeof lut);
    if (e == cudaSuccess) e = cudaMalloc(&c->stats, GI_N_STATS * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->srgb_lut, lut, sizeof lut, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->table, 0, ((size_t)capacity + 2) * sizeof(DCacheEntry), ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->pool, 0xFF, (size_t)surfel_pool_size * sizeof(float4), ctx->stream);   // fill_buffer(u32::MAX), standard.rs:347-349
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);        // lut dies at return
    if (e != cudaSuccess) {
        if (c->table) cudaFree(c->table);
        if (c->pool) cudaFree(c->pool);
        if (c->srgb_lut) cudaFree(c->srgb_lut);
        if (c->stats) cudaFree(c->stats);
        delete c;
        return dust_fail(ctx, e == cudaErrorMemoryAllocation ? DUST_ERR_OUT_OF_MEMORY : DUST_ERR_CUDA, cudaGetErrorString(e));
    }
    *out = c;
    return DUST_OK;
}

int32_t dust_cache_destroy(DustCtx *ctx, DustRadianceCache *cache)
{
    if (!ctx || !cache) return DUST_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(cache->table); cudaFree(cache->pool); cudaFree(cache->srgb_lut); cudaFree(cache->stats);
    if (cache->scratch) cudaFree(cache->scratch);
    delete cache;
    return DUST_OK;
}

int32_t dust_cache_clear(DustCtx *ctx, DustRadianceCache *cache)
{
    if (!ctx || !cache) return DUST_ERR_INVALID_ARGUMENT;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaMemsetAsync(cache->table, 0, ((size_t)cache->capacity + 2) * sizeof(DCacheEntry), ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(cache->pool, 0xFF, (size_t)cache->pool_size * sizeof(float4), ctx->stream));
    return DUST_OK;
}

int32_t dust_cache_download(DustCtx *ctx, const DustRadianceCache *cache, DustCacheEntry *entries_out, uint64_t n_entries,
                            DustSurfel *surfels_out, uint64_t n_surfels)
{
    if (!ctx || !cache || (n_entries && !entries_out) || (n_surfels && !surfels_out)) return DUST_ERR_INVALID_ARGUMENT;
    if (n_entries > (uint64_t)cache->capacity + 2 || n_surfels > cache->pool_size) return dust_fail(ctx, DUST_ERR_INVALID_ARGUMENT, "more entries than the cache holds");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if (n_entries) CU_TRY(ctx, cudaMemcpyAsync(entries_out, cache->table, n_entries * sizeof(DCacheEntry), cudaMemcpyDeviceToHost, ctx->stream));
    if (n_surfels) CU_TRY(ctx, cudaMemcpyAsync(surfels_out, cache->pool, n_surfels * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return DUST_OK;
}

int32_t dust_cache_upload(DustCtx *ctx, DustRadianceCache *cache, const DustCacheEntry *entries, uint64_t n_entries,
                          const DustSurfel *surfels, uint64_t n_surfels)
{
    if (!ctx || !cache || (n_entries && !entries) || (n_surfels && !surfels)) return DUST_ERR_INVALID_ARGUMENT;
    if (n_entries > (uint64_t)cache->capacity + 2 || n_surfels > cache->pool_size) return dust_fail(ctx, DUST_ERR_INVALID_ARGUMENT, "more entries than the cache holds");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if (n_entries) CU_TRY(ctx, cudaMemcpyAsync(cache->table, entries, n_entries * sizeof(DCacheEntry), cudaMemcpyHostToDevice, ctx->stream));
    if (n_surfels) CU_TRY(ctx, cudaMemcpyAsync(cache->pool, surfels, n_surfels * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return DUST_OK;
}

int32_t dust_trace_gi(DustCtx *ctx, const DustScene *scene, DustRadianceCache *cache, const DustCamera *camera, const DustSky *sky,
                      const DustTraceConfig *config, const DustFrame *frame, DustGiStats *stats_out)
{
    if (!ctx || !scene || !cache || !camera || !sky || !config || !frame || !frame->radiance) return DUST_ERR_INVALID_ARGUMENT;
    const uint32_t W = config->width, H = config->height, S = cache->pool_size;
    if (!W || !H || (uint64_t)W * H > 0x7FFFFFFFull) return dust_fail(ctx, DUST_ERR_INVALID_ARGUMENT, "bad frame size");
    if (frame->radiance_format != DUST_FMT_YCOCG16F) return dust_fail(ctx, DUST_ERR_INVALID_ARGUMENT, "dust_trace_gi writes the reference's img_illuminance: DUST_FMT_YCOCG16F");
    if (frame->memory != DUST_MEM_HOST && frame->memory != DUST_MEM_DEVICE) return dust_fail(ctx, DUST_ERR_INVALID_ARGUMENT, "bad memory space");
    const bool host = frame->memory == DUST_MEM_HOST;
    if (!host && (reinterpret_cast<uintptr_t>(frame->radiance) & 15u)) return dust_fail(ctx, DUST_ERR_INVALID_ARGUMENT, "device radiance buffer must be 16-byte aligned");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint32_t npx = W * H, nmax = std::max(npx, S);

    // ---- scratch: [illum | 5 G-buffer images | 2 queues (a, b, path) | hits | vis | sched | insert records | sort buffers | cub temp]
    size_t temp1 = 0, temp2 = 0, temp3 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp1, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)S, 0, 32, st);
    cub::DeviceRadixSort::SortKeys(nullptr, temp2, (const unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)S, 0, 64, st);
    cub::DeviceScan::InclusiveSum(nullptr, temp3, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)S, st);
    const size_t temp_bytes = std::max(temp1, std::max(temp2, temp3));
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = dust_align_up(off + bytes, 256); return o; };
    const size_t o_illum = take((size_t)npx * 8), o_alb = take((size_t)npx * 4), o_dep = take((size_t)npx * 4), o_nrm = take((size_t)npx * 4),
                 o_mot = take((size_t)npx * 8), o_vid = take((size_t)npx * 4);
    size_t o_q[2][3];
    for (int q = 0; q < 2; q++) { o_q[q][0] = take((size_t)nmax * 16); o_q[q][1] = take((size_t)nmax * 16); o_q[q][2] = take((size_t)nmax * 4); }
    const size_t o_hits = take((size_t)nmax * 16), o_vis = take(nmax), o_sched = take((size_t)nmax * 16);
    const size_t o_loc = take((size_t)S * 4), o_fp = take((size_t)S * 4), o_val = take((size_t)S * 16);
    const size_t o_idx = take((size_t)S * 4), o_loc_s = take((size_t)S * 4), o_idx_s = take((size_t)S * 4), o_flag = take((size_t)S * 4),
                 o_seg = take((size_t)S * 4), o_key2 = take((size_t)S * 8), o_key2s = take((size_t)S * 8), o_temp = take(temp_bytes);
    if (off > cache->scratch_bytes) {
        if (cache->scratch) { CU_TRY(ctx, cudaStreamSynchronize(st)); CU_TRY(ctx, cudaFree(cache->scratch)); cache->scratch = nullptr; cache->scratch_bytes = 0; }
        CU_TRY(ctx, cudaMalloc(&cache->scratch, off));
        cache->scratch_bytes = off;
    }
    uint8_t *base = static_cast<uint8_t *>(cache->scratch);

    GiParams p;
    memset(&p, 0, sizeof p);
    p.scene = scene->view; p.cam = *camera; p.sky = *sky;
    dust_sun_terms(sky, p.sun_dir, &p.sun_ar2, p.sun_nee);
    p.width = W; p.height = H; p.frame_index = config->frame_index;
    {
        // pcg(seed + 0x9E3779B9 * frame_index): the per-frame key of every noise stream (the path tracer's, too)
        const uint32_t v = config->seed + 0x9E3779B9u * config->frame_index;
        const uint32_t state = v * 747796405u + 2891336453u;
        const uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
        p.key = (word >> 22u) ^ word;
    }
    p.illum = host ? reinterpret_cast<uint2 *>(base + o_illum) : static_cast<uint2 *>(frame->radiance);
    p.gb.albedo = reinterpret_cast<uint32_t *>(base + o_alb); p.gb.depth = reinterpret_cast<float *>(base + o_dep);
    p.gb.normal = reinterpret_cast<uint32_t *>(base + o_nrm); p.gb.motion = reinterpret_cast<uint16_t *>(base + o_mot);
    p.gb.voxel_id = reinterpret_cast<uint32_t *>(base + o_vid);
    p.table = cache->table; p.capacity = cache->capacity; p.pool = cache->pool; p.pool_size = S;
    WfQueue *qs[2] = { &p.q_any, &p.q_closest };
    for (int q = 0; q < 2; q++) {
        qs[q]->a = reinterpret_cast<float4 *>(base + o_q[q][0]); qs[q]->b = reinterpret_cast<float4 *>(base + o_q[q][1]);
        qs[q]->path = reinterpret_cast<uint32_t *>(base + o_q[q][2]); qs[q]->stride = 1;
    }
    p.hits = reinterpret_cast<uint4 *>(base + o_hits); p.vis = base + o_vis; p.sched = reinterpret_cast<float4 *>(base + o_sched);
    p.ins_loc = reinterpret_cast<uint32_t *>(base + o_loc); p.ins_fp = reinterpret_cast<uint32_t *>(base + o_fp);
    p.ins_val = reinterpret_cast<float4 *>(base + o_val);
    p.stats = cache->stats; p.srgb_lut = cache->srgb_lut;
    uint32_t *d_idx = reinterpret_cast<uint32_t *>(base + o_idx), *d_loc_s = reinterpret_cast<uint32_t *>(base + o_loc_s),
             *d_idx_s = reinterpret_cast<uint32_t *>(base + o_idx_s), *d_flag = reinterpret_cast<uint32_t *>(base + o_flag),
             *d_seg = reinterpret_cast<uint32_t *>(base + o_seg);
    unsigned long long *d_key2 = reinterpret_cast<unsigned long long *>(base + o_key2), *d_key2s = reinterpret_cast<unsigned long long *>(base + o_key2s);

    CU_TRY(ctx, cudaMemsetAsync(cache->stats, 0, GI_N_STATS * sizeof(unsigned long long), st));
    // the reference leaves normal / voxel_id of miss pixels untouched: start from the caller's contents (host frames), else from a defined state
    const DustGBuffer *ugb = frame->gbuffer;
    if (ugb && ugb->normal) CU_TRY(ctx, cudaMemcpyAsync(p.gb.normal, ugb->normal, (size_t)npx * 4, host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
    else CU_TRY(ctx, cudaMemsetAsync(p.gb.normal, 0, (size_t)npx * 4, st));
    if (ugb && ugb->voxel_id) CU_TRY(ctx, cudaMemcpyAsync(p.gb.voxel_id, ugb->voxel_id, (size_t)npx * 4, host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
    else CU_TRY(ctx, cudaMemsetAsync(p.gb.voxel_id, 0xFF, (size_t)npx * 4, st));

    const int nl = scene->n_inst == 0 || scene->min_levels == scene->max_levels ? (scene->n_inst ? (int)scene->max_levels : 3) : 0;
    WfParams wp;                                                        // what wf_trace needs of the wavefront parameters
    memset(&wp, 0, sizeof wp);
    wp.f.scene = scene->view; wp.f.counters = ctx->d_counters;
    wp.hits = p.hits; wp.vis = p.vis;
    auto trace = [&](const WfQueue &q, uint32_t n, int leaf) {
        wp.n_paths = n;
        const uint32_t grid = (n + DUST_THREADS - 1) / DUST_THREADS;
#define GI_TRACE(NL)                                                                                               \
        do {                                                                                                       \
            if (leaf == LEAF_AO) wf_trace<NL, false, LEAF_AO><<<grid, DUST_THREADS, 0, st>>>(wp, q, nullptr, nullptr); \
            else wf_trace<NL, false, LEAF_ROUGH><<<grid, DUST_THREADS, 0, st>>>(wp, q, nullptr, nullptr);             \
        } while (0)
        if (nl == 3) GI_TRACE(3); else if (nl == 4) GI_TRACE(4); else GI_TRACE(0);
#undef GI_TRACE
        dust_count_launch(ctx, "wf_trace (gi)");
    };
    const uint32_t gpx = (npx + 255) / 256, gS = (S + 255) / 256;

    // primary
    {
        const uint32_t tiles_x = (W + 31) / 32, tiles_y = (H + 31) / 32, n_tiles = tiles_x * tiles_y;
        p.tiles_x = tiles_x;
        if (scene->n_inst > 1 && scene->d_boxes) {
            // the camera rays' candidate lists, per tile (kernels.cuh wf_tile_cull: only the camera, the frame size and the tile grid are read)
            const size_t words = (size_t)n_tiles * DUST_TILE_LIST_WORDS;
            if (words > ctx->tile_lists_words) {
                if (ctx->d_tile_lists) { CU_TRY(ctx, cudaStreamSynchronize(st)); CU_TRY(ctx, cudaFree(ctx->d_tile_lists)); ctx->d_tile_lists = nullptr; ctx->tile_lists_words = 0; }
                CU_TRY(ctx, cudaMalloc(&ctx->d_tile_lists, words * sizeof(uint32_t)));
                ctx->tile_lists_words = words;
            }
            FrameParams fp;
            memset(&fp, 0, sizeof fp);
            fp.scene = scene->view; fp.cam = *camera;
            fp.cfg.width = W; fp.cfg.height = H; fp.cfg.shard_count = 1; fp.cfg.shard_index = 0;
            fp.tiles_x = tiles_x; fp.tiles_y = tiles_y; fp.n_local_tiles = n_tiles;
            wf_tile_cull<<<(n_tiles + 3) / 4, 128, 0, st>>>(fp, ctx->d_tile_lists);
            dust_count_launch(ctx, "wf_tile_cull");
            p.tile_lists = ctx->d_tile_lists;
        }
        const uint32_t grid = n_tiles * (1024u / DUST_THREADS);
        if (nl == 3) gi_primary<3><<<grid, DUST_THREADS, 0, st>>>(p); else if (nl == 4) gi_primary<4><<<grid, DUST_THREADS, 0, st>>>(p); else gi_primary<0><<<grid, DUST_THREADS, 0, st>>>(p);
        dust_count_launch(ctx, "gi_primary");
    }
    // ambient occlusion: sun shadow ray + near segment, both with ambient_occlusion.rint
    gi_gen_ao<<<gpx, 256, 0, st>>>(p); dust_count_launch(ctx, "gi_gen_ao");
    trace(p.q_any, npx, LEAF_AO);
    trace(p.q_closest, npx, LEAF_AO);
    gi_ao_resolve<<<gpx, 256, 0, st>>>(p); dust_count_launch(ctx, "gi_ao_resolve");
    // final gather: far segment with rough.rint, cache look-up, surfel scheduling
    gi_gen_fg<<<gpx, 256, 0, st>>>(p); dust_count_launch(ctx, "gi_gen_fg");
    trace(p.q_closest, npx, LEAF_ROUGH);
    gi_fg_shade<<<gpx, 256, 0, st>>>(p); dust_count_launch(ctx, "gi_fg_shade");
    gi_pool_update<<<gS, 256, 0, st>>>(p); dust_count_launch(ctx, "gi_pool_update");
    // surfel pass
    gi_gen_surfel<<<gS, 256, 0, st>>>(p); dust_count_launch(ctx, "gi_gen_surfel");
    trace(p.q_any, S, LEAF_ROUGH);
    trace(p.q_closest, S, LEAF_ROUGH);
    gi_surfel_shade<<<gS, 256, 0, st>>>(p); dust_count_launch(ctx, "gi_surfel_shade");
    // the inserts, in ascending surfel index wherever two of them can touch the same entries
    {
        // sort 1: by probe location, surfel index as payload (stable: equal locations stay in ascending surfel index)
        gi_iota<<<gS, 256, 0, st>>>(d_idx, S); dust_count_launch(ctx, "gi_iota");
        size_t tb = temp_bytes;
        CU_TRY(ctx, cub::DeviceRadixSort::SortPairs(base + o_temp, tb, (const uint32_t *)p.ins_loc, d_loc_s, (const uint32_t *)d_idx, d_idx_s, (int)S, 0, 32, st));
        gi_ins_flags<<<gS, 256, 0, st>>>(d_loc_s, S, d_flag); dust_count_launch(ctx, "gi_ins_flags");
        tb = temp_bytes;
        CU_TRY(ctx, cub::DeviceScan::InclusiveSum(base + o_temp, tb, (const uint32_t *)d_flag, d_seg, (int)S, st));
        gi_ins_keys<<<gS, 256, 0, st>>>(d_loc_s, d_idx_s, d_seg, S, d_key2); dust_count_launch(ctx, "gi_ins_keys");
        tb = temp_bytes;
        CU_TRY(ctx, cub::DeviceRadixSort::SortKeys(base + o_temp, tb, (const unsigned long long *)d_key2, d_key2s, (int)S, 0, 64, st));
        gi_ins_apply<<<(S + 127) / 128, 128, 0, st>>>(p, d_key2s, S); dust_count_launch(ctx, "gi_ins_apply");
    }
    CU_TRY(ctx, cudaGetLastError());

    // results
    const size_t gb_bytes[5] = { 4, 4, 4, 8, 4 };
    if (ugb) {
        void *dst[5] = { ugb->albedo, ugb->depth, ugb->normal, ugb->motion, ugb->voxel_id };
        const void *src[5] = { p.gb.albedo, p.gb.depth, p.gb.normal, p.gb.motion, p.gb.voxel_id };
        for (int k = 0; k < 5; k++)
            if (dst[k]) CU_TRY(ctx, cudaMemcpyAsync(dst[k], src[k], (size_t)npx * gb_bytes[k], host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
    }
    if (host) CU_TRY(ctx, cudaMemcpyAsync(frame->radiance, p.illum, (size_t)npx * 8, cudaMemcpyDeviceToHost, st));
    if (stats_out) {
        unsigned long long h[GI_N_STATS];
        CU_TRY(ctx, cudaMemcpyAsync(h, cache->stats, sizeof h, cudaMemcpyDeviceToHost, st));
        CU_TRY(ctx, cudaStreamSynchronize(st));
        stats_out->near_hits = h[GI_NEAR_HITS]; stats_out->far_hits = h[GI_FAR_HITS]; stats_out->far_misses = h[GI_FAR_MISSES];
        stats_out->cache_found = h[GI_CACHE_FOUND]; stats_out->surfels_scheduled = h[GI_SCHEDULED]; stats_out->surfel_rays = h[GI_SURFEL_RAYS];
        stats_out->cache_inserts = h[GI_INSERTS]; stats_out->insert_segments = h[GI_SEGMENTS];
    } else if (host) CU_TRY(ctx, cudaStreamSynchronize(st));
    return DUST_OK;
}

}  // extern "C"
