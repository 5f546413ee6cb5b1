// sm_100a device code, part 7: the reference's own global illumination for one frame — what
// StandardPipeline::render's ambient-occlusion, final-gather and surfel dispatches compute (crates/render/src/pipeline/standard.rs:
// 564-789) over the persistent world-space radiance cache and surfel pool (standard.rs:334-358).  Reference sources restated here:
//   radiance cache      assets/shaders/headers/spatial_hash.glsl:21-93 (LogLuv codec), :105-145 (pcg, xxhash32, fingerprint, location),
//                       :150-220 (insert: 3 probes, capped running mean, LRU eviction; get)
//   face ids            headers/normal.glsl:9-26
//   normal image        headers/nrd.glsl:54-94 (NRD_FrontEnd_UnpackNormalAndRoughness), illuminance image nrd.glsl:99-147
//   ambient occlusion   final_gather/ambient_occlusion.rgen:14-66, .rchit, .rmiss, final_gather/nee.rmiss:11-22
//   final gather        final_gather/final_gather.rgen:14-52, .rchit:35-91, .rmiss:12-24
//   surfel pass         surfel/surfel.rgen:12-67, surfel.rchit:35-102, surfel.rmiss:14-26, surfel/nee.rmiss:15-27
// Rays are traced by wf_trace (wavefront.cuh) with the pass' own leaf step: ambient_occlusion.rint for the two rays of the
// ambient-occlusion pass, rough.rint (block-granular) for the final-gather and surfel rays.  Every kernel is one thread per pixel
// or per pool slot; the passes talk through the RGBA16F illuminance image exactly like the reference (values are rounded to
// half floats between passes).
//
// Determinism.  The reference's invocations race on the cache and the pool.  Here a pass performs all of its look-ups against the
// table as the previous pass left it (a look-up only stamps last_accessed_frame with this frame's index: every writer stores the same
// value), pool writes go to the highest pixel of a slot (gi_pool_update gathers them: no atomics), and the surfel pass' inserts are
// applied by gi_ins_apply in ascending surfel index: sorted by probe location, cut where two consecutive locations lie more
// than two apart (windows of 3 cannot overlap across such a gap), re-sorted by (segment, surfel) — one thread walks one segment.
#pragma once
#include "wavefront.cuh"
#include "post.cuh"

namespace dust {

struct DCacheEntry { uint32_t fingerprint, radiance; uint16_t last_accessed_frame, sample_count; };   // SpatialHashEntry (12 bytes)
static_assert(sizeof(DCacheEntry) == 12, "SpatialHashEntry layout");

enum : int { GI_NEAR_HITS = 0, GI_FAR_HITS, GI_FAR_MISSES, GI_CACHE_FOUND, GI_SCHEDULED, GI_SURFEL_RAYS, GI_INSERTS, GI_SEGMENTS, GI_N_STATS };

struct GiParams {
    DSceneView scene;
    DustCamera cam;
    DustSky sky;
    float sun_dir[3];          // normalize(sunlight_config.direction.xyz)
    float sun_ar2;
    float sun_nee[3];          // sun_radiance(normalize(sun_dir)) * (1 - cos(solar radius)), ACEScg
    uint32_t width, height, frame_index, key;      // key = pcg(seed + 0x9E3779B9 * frame_index)
    uint2 *illum;              // RGBA16F: YCoCg + normHitDist
    DustGBuffer gb;            // device images, all five present
    DCacheEntry *table;
    uint32_t capacity;
    float4 *pool;              // SurfelEntry: xyz = block centre, w = face as bits
    uint32_t pool_size;
    WfQueue q_any, q_closest;  // slot i = pixel / pool slot i (dead rays have tmax < tmin)
    uint4 *hits;
    uint8_t *vis;
    float4 *sched;             // the pool write a pixel / surfel invocation asks for (w = 0xFFFFFFFF: none)
    uint32_t *ins_loc, *ins_fp;       // the insert of surfel invocation i: probe location (0xFFFFFFFF: none), fingerprint
    float4 *ins_val;
    unsigned long long *stats; // GI_N_STATS counters
    const uint32_t *tile_lists;    // per 32x32 tile of the row-major frame: candidate instances of its camera rays (wf_tile_cull), or NULL
    uint32_t tiles_x;
    const float *srgb_lut;     // SRGBToLinear(v / 1023) for the 1024 UNORM10 levels of Block.avg_albedo (host libm powf, like palette_lin)
};

// ------------------------------------------------------------------------------------------ radiance cache

__device__ __forceinline__ uint32_t xxhash32(uint32_t p)                       // spatial_hash.glsl:115-126
{
    const uint32_t PRIME32_2 = 2246822519u, PRIME32_3 = 3266489917u, PRIME32_4 = 668265263u, PRIME32_5 = 374761393u;
    uint32_t h32 = p + PRIME32_5;
    h32 = PRIME32_4 * ((h32 << 17) | (h32 >> (32 - 17)));
    h32 = PRIME32_2 * (h32 ^ (h32 >> 15));
    h32 = PRIME32_3 * (h32 ^ (h32 >> 13));
    return h32 ^ (h32 >> 16);
}
struct CacheKey { int x, y, z; uint32_t direction; };
__device__ __forceinline__ uint32_t cache_fingerprint(const CacheKey &k)      // :128-135
{
    uint32_t hash = xxhash32((uint32_t)k.x);
    hash = xxhash32((uint32_t)k.y + hash);
    hash = xxhash32((uint32_t)k.z + hash);
    hash = xxhash32(k.direction + hash);
    return hash > 1u ? hash : 1u;
}
__device__ __forceinline__ uint32_t cache_location(const CacheKey &k, uint32_t capacity)    // :136-142
{
    uint32_t hash = pcg((uint32_t)k.x);
    hash = pcg((uint32_t)k.y + hash);
    hash = pcg((uint32_t)k.z + hash);
    hash = pcg(k.direction + hash);
    return hash % capacity;
}
__device__ __forceinline__ uint32_t normal_to_face_id(f3 n)                    // normal.glsl:9-18
{
    const float sum = (n.x + n.y) + n.z;
    const float s = sum < 0.0f ? 0.0f : (sum > 1.0f ? 1.0f : sum);
    uint32_t face = (uint32_t)rintf(s);
    face += (uint32_t)rintf(fabsf(n.z)) * 4u + (uint32_t)rintf(fabsf(n.y)) * 2u;
    return face & 0xFFu;
}
__device__ __forceinline__ f3 face_id_to_normal(uint32_t face)                 // normal.glsl:20-26
{
    const float s = (float)(face & 1u) * 2.0f - 1.0f;
    const uint32_t axis = (face >> 1) % 3u;
    return mk3(axis == 0u ? s : 0.0f, axis == 1u ? s : 0.0f, axis == 2u ? s : 0.0f);
}
__device__ __forceinline__ CacheKey cache_key_of(f3 c, f3 normal_world)        // final_gather.rchit:45-47
{
    CacheKey k;
    k.x = (int)(c.x / 4.0f); k.y = (int)(c.y / 4.0f); k.z = (int)(c.z / 4.0f);   // ivec3(): truncation toward zero
    k.direction = normal_to_face_id(normal_world);
    return k;
}

__device__ __forceinline__ uint32_t logluv_encode(f3 rgb)                      // spatial_hash.glsl:27-58
{
    const float m[9] = { 0.66245437f, 0.2722288f, -0.0055746622f, 0.13400422f, 0.6740818f, 0.00406073f, 0.15618773f, 0.05368953f, 1.0103393f };
    const f3 XYZ = mat3_mul(m, rgb);
    if (!(XYZ.y > 0.0f)) return 0u;
    const float logY = 409.6f * (post_log2(XYZ.y) + 20.0f);
    const float cl = logY < 0.0f ? 0.0f : (logY > 16383.0f ? 16383.0f : logY);
    const uint32_t Le = (uint32_t)cl;
    if (Le == 0u) return 0u;
    const float invDenom = 1.0f / ((-2.0f * XYZ.x + 12.0f * XYZ.y) + 3.0f * ((XYZ.x + XYZ.y) + XYZ.z));
    const float u = (4.0f * XYZ.x) * invDenom, v = (9.0f * XYZ.y) * invDenom;
    const float su = 820.0f * u, sv = 820.0f * v;
    const uint32_t ue = (uint32_t)(su < 0.0f ? 0.0f : (su > 511.0f ? 511.0f : su)), ve = (uint32_t)(sv < 0.0f ? 0.0f : (sv > 511.0f ? 511.0f : sv));
    return (Le << 18) | (ue << 9) | ve;
}
__device__ __forceinline__ f3 logluv_decode(uint32_t packed)                   // :62-93
{
    const uint32_t Le = packed >> 18;
    if (Le == 0u) return mk3(0.0f, 0.0f, 0.0f);
    const float logY = ((float)Le + 0.5f) / 409.6f - 20.0f;
    const float Y = exp2_fixed(logY);
    const float u = ((float)((packed >> 9) & 0x1ffu) + 0.5f) / 820.0f, v = ((float)(packed & 0x1ffu) + 0.5f) / 820.0f;
    const float invDenom = 1.0f / ((6.0f * u - 16.0f * v) + 12.0f);
    const float x = (9.0f * u) * invDenom, y = (4.0f * v) * invDenom;
    const float s = Y / y;
    const float m[9] = { 1.6410228f, -0.66366285f, 0.011721907f, -0.32480323f, 1.6153315f, -0.0082844375f, -0.23642465f, 0.016756356f, 0.9883947f };
    const f3 out = mat3_mul(m, mk3(s * x, Y, s * ((1.0f - x) - y)));
    return mk3(out.x > 0.0f ? out.x : 0.0f, out.y > 0.0f ? out.y : 0.0f, out.z > 0.0f ? out.z : 0.0f);
}

// SpatialHashGet, :202-220.  The table is read-only during a pass except for this stamp (the same value from every reader).
__device__ __forceinline__ bool cache_get(DCacheEntry *t, uint32_t capacity, const CacheKey &key, uint32_t frame_index, f3 &value, uint32_t &sample_count)
{
    const uint32_t fingerprint = cache_fingerprint(key), location = cache_location(key, capacity);
    value = mk3(0.0f, 0.0f, 0.0f);
    sample_count = 0u;
    for (uint32_t i = 0; i < 3u; i++) {
        DCacheEntry *e = &t[location + i];
        const uint32_t fp = e->fingerprint;
        if (fp == 0u) return false;
        if (fp == fingerprint) {
            e->last_accessed_frame = (uint16_t)frame_index;
            value = logluv_decode(e->radiance);
            sample_count = e->sample_count;
            return true;
        }
    }
    return false;
}
// SpatialHashInsert, :150-197, as ONE thread applies it (sequential form of the atomicCompSwap)
__device__ __forceinline__ void cache_insert(DCacheEntry *t, uint32_t location, uint32_t fingerprint, f3 value, uint32_t frame_index)
{
    uint32_t i_min = 0, min_frame = 0;
    for (uint32_t i = 0; i < 3u; i++) {
        DCacheEntry *e = &t[location + i];
        const uint32_t current = e->fingerprint;
        if (current == 0u) e->fingerprint = fingerprint;
        const uint32_t frame = e->last_accessed_frame;
        if (i == 0u || frame < min_frame) { i_min = i; min_frame = frame; }
        if (current == fingerprint || current == 0u) {
            f3 cur = mk3(0.0f, 0.0f, 0.0f);
            uint32_t count = 0;
            if (current == fingerprint) { count = e->sample_count; cur = logluv_decode(e->radiance); }
            if (count > 403u) count = 403u;                                      // MAX_SAMPLE_COUNT - 1
            const uint32_t next = count + 1u;
            const float a = 1.0f / (float)next;
            const f3 mixed = mk3(cur.x * (1.0f - a) + value.x * a, cur.y * (1.0f - a) + value.y * a, cur.z * (1.0f - a) + value.z * a);   // mix(x, y, a)
            e->radiance = logluv_encode(mixed);
            e->last_accessed_frame = (uint16_t)frame_index;
            e->sample_count = (uint16_t)next;
            return;
        }
    }
    DCacheEntry *e = &t[location + i_min];                                       // evict the least recently accessed of the three
    e->fingerprint = fingerprint;
    e->radiance = logluv_encode(value);
    e->last_accessed_frame = (uint16_t)frame_index;
    e->sample_count = 1;
}

// ------------------------------------------------------------------------------------------ images

// REBLUR_FrontEnd_PackRadianceAndNormHitDist + the RGBA16F store (nrd.glsl:130-147)
__device__ __forceinline__ uint2 pack_illuminance(f3 r, float norm_hit_dist)
{
    if (norm_hit_dist != 0.0f) norm_hit_dist = fmaxf(norm_hit_dist, 1e-7f);
    const float Y = (r.x * 0.25f + r.y * 0.5f) + r.z * 0.25f;
    const float Co = (r.x * 0.5f + r.y * 0.0f) + r.z * -0.5f;
    const float Cg = (r.x * -0.25f + r.y * 0.5f) + r.z * -0.25f;
    const __half2 lo = __floats2half2_rn(Y, Co), hi = __floats2half2_rn(Cg, norm_hit_dist);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t *>(&lo); pk.y = *reinterpret_cast<const uint32_t *>(&hi);
    return pk;
}
__device__ __forceinline__ float illuminance_w(uint2 px) { return __high2float(*reinterpret_cast<const __half2 *>(&px.y)); }

// texel load of the A2R10G10B10_UNORM normal image + NRD_FrontEnd_UnpackNormalAndRoughness(...).xyz (nrd.glsl:54-94)
__device__ __forceinline__ f3 unpack_normal(uint32_t packed)
{
    float px = (float)((packed >> 20) & 1023u) / 1023.0f, py = (float)((packed >> 10) & 1023u) / 1023.0f;
    px = px * 2.0f - 1.0f; py = py * 2.0f - 1.0f;
    f3 v = mk3(px, py, (1.0f - fabsf(px)) - fabsf(py));
    const float t = fminf(fmaxf(-v.z, 0.0f), 1.0f);
    v.x -= t * ((v.x < 0.0f ? 0.0f : 1.0f) * 2.0f - 1.0f);
    v.y -= t * ((v.y < 0.0f ? 0.0f : 1.0f) * 2.0f - 1.0f);
    return normalize3(v);
}

// the cosine-weighted unit vector the reference fetches from its blue-noise texture (pcg streams instead, DESIGN.md deviation (3))
__device__ __forceinline__ f3 cosine_sample(uint32_t s1)
{
    const float u1 = u01(pcg(s1 + 2u)), u2 = u01(pcg(s1 + 3u));
    const float r = sqrtf(u1), cz = sqrtf(1.0f - u1);
    float c, s;
    cossin_2pi(u2, c, s);
    return mk3(r * c, r * s, cz);
}

// what ambient_occlusion.rgen:15-30 and final_gather.rgen:15-34 both derive from a pixel's G-buffer
__device__ __forceinline__ bool gi_pixel_setup(const GiParams &p, uint32_t pi, f3 &n, f3 &x, f3 &dir, uint32_t &s1)
{
    const float hit_t = p.gb.depth[pi];
    if (hit_t == CUDART_INF_F) return false;
    n = unpack_normal(p.gb.normal[pi]);
    f3 o, d;
    camera_ray(p.cam, p.width, p.height, pi % p.width, pi / p.width, 0.5f, 0.5f, o, d);
    x = mk3((hit_t * d.x + o.x) + n.x * 0.01f, (hit_t * d.y + o.y) + n.y * 0.01f, (hit_t * d.z + o.z) + n.z * 0.01f);
    s1 = pcg(pcg(pi ^ p.key) + 0u);
    dir = normalize3(rotate_by_normal(n, cosine_sample(s1)));
    return true;
}

__device__ __forceinline__ void gi_dead_ray(const WfQueue &q, uint32_t i, uint32_t path)
{
    q.a[i] = make_float4(0.0f, 0.0f, 0.0f, 1.0f); q.b[i] = make_float4(1.0f, 1.0f, 1.0f, 0.0f); q.path[i] = path;   // tmax < tmin: traces nothing
}

__device__ __forceinline__ f3 sky_only(const GiParams &p, f3 dn)
{
    f3 sk, su;
    sky_eval(p.sky, p.sun_ar2, dn.x, dn.y, dn.z, false, sk.x, sk.y, sk.z, su.x, su.y, su.z);
    return sk;
}

// final_gather.rchit:38-47 / surfel.rchit:36-47: the block a block-granular hit landed on
struct BlockHit { f3 center; CacheKey key; uint32_t avg_albedo; };
__device__ __forceinline__ BlockHit gi_block_of_hit(const DSceneView &sc, uint4 hw, f3 ro, f3 rd)
{
    const uint32_t inst = hw.y & 0xFFFFFu;
    const DInstance &in = sc.inst[inst];
    f3 o, d;
    object_ray(in, ro, rd, o, d);
    const float t = __uint_as_float(hw.x);
    const f3 center_obj = mk3((float)((int)(hw.w & 1023u) << 2) + 2.0f, (float)((int)((hw.w >> 10) & 1023u) << 2) + 2.0f, (float)((int)(hw.w >> 20) << 2) + 2.0f);
    const f3 normal_obj = mk3((t * d.x + o.x) - center_obj.x, (t * d.y + o.y) - center_obj.y, (t * d.z + o.z) - center_obj.z);
    const float *W = in.o2w;
    const f3 nw = mk3((W[0] * normal_obj.x + W[1] * normal_obj.y) + W[2] * normal_obj.z, (W[3] * normal_obj.x + W[4] * normal_obj.y) + W[5] * normal_obj.z,
                      (W[6] * normal_obj.x + W[7] * normal_obj.y) + W[8] * normal_obj.z);
    BlockHit b;
    b.center = mk3(((W[0] * center_obj.x + W[1] * center_obj.y) + W[2] * center_obj.z) + in.o2w_t[0],
                   ((W[3] * center_obj.x + W[4] * center_obj.y) + W[5] * center_obj.z) + in.o2w_t[1],
                   ((W[6] * center_obj.x + W[7] * center_obj.y) + W[8] * center_obj.z) + in.o2w_t[2]);
    const float ax = fabsf(nw.x), ay = fabsf(nw.y), az = fabsf(nw.z);
    const float mx = fmaxf(ax, fmaxf(ay, az));                                   // CubedNormalize
    b.key = cache_key_of(b.center, mk3(signf(nw.x) * (ax >= mx ? 1.0f : 0.0f), signf(nw.y) * (ay >= mx ? 1.0f : 0.0f), signf(nw.z) * (az >= mx ? 1.0f : 0.0f)));
    b.avg_albedo = __ldg(&in.leaves[hw.z]).w;
    return b;
}

// final_gather.rchit:68-80 / surfel.rchit:60-71
__device__ __forceinline__ f3 gi_modulate(const float *__restrict__ srgb_lut, uint32_t packed_albedo, f3 radiance)
{
    const f3 s = acescg2srgb(radiance);
    return srgb2acescg(mk3(s.x * __ldg(&srgb_lut[(packed_albedo >> 22) & 1023u]), s.y * __ldg(&srgb_lut[(packed_albedo >> 12) & 1023u]),
                           s.z * __ldg(&srgb_lut[(packed_albedo >> 2) & 1023u])));
}

// ------------------------------------------------------------------------------------------ passes

// primary pass: closest hit of the pixel-centre camera ray, G-buffer (hit.rchit:57-94 / miss.rmiss:8-16), illuminance 0 on a hit
template <int NL>
__global__ void __launch_bounds__(DUST_THREADS, DUST_MIN_BLOCKS) gi_primary(const __grid_constant__ GiParams p)
{
    __shared__ uint32_t s_scratch[DUST_SCRATCH_WORDS * DUST_THREADS];
    // thread -> pixel: 32x32 tiles, row-major inside a tile, so that a warp is one tile row and shares the tile's candidate list
    const uint32_t item = blockIdx.x * DUST_THREADS + threadIdx.x, tile = item >> 10, q = item & 1023u;
    const uint32_t px = (tile % p.tiles_x) * 32u + (q & 31u), py = (tile / p.tiles_x) * 32u + (q >> 5);
    const bool live = px < p.width && py < p.height;
    const uint32_t pi = py * p.width + px;
    const uint32_t *tl = p.tile_lists ? p.tile_lists + (size_t)tile * DUST_TILE_LIST_WORDS : nullptr;
    f3 ro = mk3(0.0f, 0.0f, 0.0f), rd = mk3(1.0f, 1.0f, 1.0f);
    if (live) camera_ray(p.cam, p.width, p.height, px, py, 0.5f, 0.5f, ro, rd);
    Counters st = {};
    Hit h;
    trace_scene<NL, false>(p.scene, ro, rd, live ? p.cam.near : 1.0f, live ? p.cam.far : 0.0f, false, live, s_scratch + threadIdx.x, h, st, LEAF_PRECISE, tl);
    if (!live) return;
    if (h.hit) {
        f3 sn, x = ro;
        uint32_t pidx = 0;
        surface_at<false>(p.scene, h, x, rd, sn, st, &pidx);
        gbuffer_store(p.gb, p.scene, pi, true, h.t, h.inst, h.voxel, pidx, ro, rd, sn);
        p.illum[pi] = pack_illuminance(mk3(0.0f, 0.0f, 0.0f), 0.0f);
    } else {
        gbuffer_store(p.gb, p.scene, pi, false, 0.0f, 0u, 0u, 0u, ro, rd, ro);
        const f3 dn = normalize3(rd);
        f3 sk, su;
        sky_eval(p.sky, p.sun_ar2, dn.x, dn.y, dn.z, true, sk.x, sk.y, sk.z, su.x, su.y, su.z);
        p.illum[pi] = pack_illuminance(mk3((sk.x + su.x) / 3.14f, (sk.y + su.y) / 3.14f, (sk.z + su.z) / 3.14f), 100000.0f);
    }
}

// ambient_occlusion.rgen: the sun shadow ray (any hit) and the near segment [0.1, 8] of the gather ray (closest hit) of every pixel
__global__ void __launch_bounds__(256) gi_gen_ao(const __grid_constant__ GiParams p)
{
    const uint32_t pi = blockIdx.x * 256u + threadIdx.x;
    if (pi >= p.width * p.height) return;
    f3 n, x, dir;
    uint32_t s1;
    if (!gi_pixel_setup(p, pi, n, x, dir, s1)) { gi_dead_ray(p.q_any, pi, pi | 0x80000000u); gi_dead_ray(p.q_closest, pi, pi); return; }
    if (dot3(mk3(p.sky.direction[0], p.sky.direction[1], p.sky.direction[2]), n) > 0.0f) {
        p.q_any.a[pi] = make_float4(x.x, x.y, x.z, 0.1f); p.q_any.b[pi] = make_float4(p.sun_dir[0], p.sun_dir[1], p.sun_dir[2], 10000.0f);
        p.q_any.path[pi] = pi | 0x80000000u;
    } else gi_dead_ray(p.q_any, pi, pi | 0x80000000u);
    p.q_closest.a[pi] = make_float4(x.x, x.y, x.z, 0.1f); p.q_closest.b[pi] = make_float4(dir.x, dir.y, dir.z, 8.0f);
    p.q_closest.path[pi] = pi;
}

// nee.rmiss + ambient_occlusion.rchit / .rmiss
__global__ void __launch_bounds__(256) gi_ao_resolve(const __grid_constant__ GiParams p)
{
    const uint32_t pi = blockIdx.x * 256u + threadIdx.x;
    if (pi >= p.width * p.height) return;
    f3 n, x, dir;
    uint32_t s1;
    if (!gi_pixel_setup(p, pi, n, x, dir, s1)) return;
    f3 L = unpack_illuminance(p.illum[pi]);
    if (dot3(mk3(p.sky.direction[0], p.sky.direction[1], p.sky.direction[2]), n) > 0.0f && !p.vis[pi]) {
        const float ndl = dot3(n, mk3(p.sun_dir[0], p.sun_dir[1], p.sun_dir[2]));
        L.x += p.sun_nee[0] * ndl; L.y += p.sun_nee[1] * ndl; L.z += p.sun_nee[2] * ndl;
    }
    const uint4 hw = p.hits[pi];
    const bool hit = hw.y != 0xFFFFFFFFu;
    if (hit) atomicAdd(&p.stats[GI_NEAR_HITS], 1ull);
    p.illum[pi] = pack_illuminance(L, hit ? __uint_as_float(hw.x) : 0.0f);
}

// final_gather.rgen: the far segment [8, far] of the same direction, for pixels the near segment did not resolve
__global__ void __launch_bounds__(256) gi_gen_fg(const __grid_constant__ GiParams p)
{
    const uint32_t pi = blockIdx.x * 256u + threadIdx.x;
    if (pi >= p.width * p.height) return;
    f3 n, x, dir;
    uint32_t s1;
    if (!gi_pixel_setup(p, pi, n, x, dir, s1) || illuminance_w(p.illum[pi]) > 0.0f) { gi_dead_ray(p.q_closest, pi, pi); return; }
    p.q_closest.a[pi] = make_float4(x.x, x.y, x.z, 8.0f); p.q_closest.b[pi] = make_float4(dir.x, dir.y, dir.z, p.cam.far);
    p.q_closest.path[pi] = pi;
}

// final_gather.rchit / .rmiss
__global__ void __launch_bounds__(256) gi_fg_shade(const __grid_constant__ GiParams p)
{
    const uint32_t pi = blockIdx.x * 256u + threadIdx.x;
    if (pi >= p.width * p.height) return;
    p.sched[pi] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0xFFFFFFFFu));
    f3 n, x, dir;
    uint32_t s1;
    if (!gi_pixel_setup(p, pi, n, x, dir, s1)) return;
    const uint2 px = p.illum[pi];
    if (illuminance_w(px) > 0.0f) return;
    f3 L = unpack_illuminance(px);
    const uint4 hw = p.hits[pi];
    if (hw.y == 0xFFFFFFFFu) {
        const f3 sk = sky_only(p, normalize3(dir));
        p.illum[pi] = pack_illuminance(mk3(L.x + sk.x, L.y + sk.y, L.z + sk.z), 0.0f);
        atomicAdd(&p.stats[GI_FAR_MISSES], 1ull);
        return;
    }
    atomicAdd(&p.stats[GI_FAR_HITS], 1ull);
    const BlockHit b = gi_block_of_hit(p.scene, hw, x, dir);
    f3 radiance;
    uint32_t count;
    if (cache_get(p.table, p.capacity, b.key, p.frame_index, radiance, count)) atomicAdd(&p.stats[GI_CACHE_FOUND], 1ull);
    const float prob = 1.0f / (float)(count + 2u);
    if (u01(pcg(s1 + 6u)) > prob) {
        p.sched[pi] = make_float4(b.center.x, b.center.y, b.center.z, __uint_as_float(b.key.direction));
        atomicAdd(&p.stats[GI_SCHEDULED], 1ull);
    }
    const f3 out = gi_modulate(p.srgb_lut, b.avg_albedo, radiance);
    p.illum[pi] = pack_illuminance(mk3(L.x + out.x, L.y + out.y, L.z + out.z), __uint_as_float(hw.x));
}

// surfel_pool[pixel % SurfelPoolSize] = entry (final_gather.rchit:56-62): the highest pixel of a slot wins
__global__ void __launch_bounds__(256) gi_pool_update(const __grid_constant__ GiParams p)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x, npx = p.width * p.height;
    if (i >= p.pool_size || i >= npx) return;
    for (uint32_t pi = i + ((npx - 1u - i) / p.pool_size) * p.pool_size;; pi -= p.pool_size) {
        const float4 s = p.sched[pi];
        if (__float_as_uint(s.w) != 0xFFFFFFFFu) { p.pool[i] = s; return; }
        if (pi < p.pool_size) return;
    }
}

__device__ __forceinline__ bool gi_surfel_setup(const GiParams &p, uint32_t i, float4 &e, f3 &n, f3 &x, f3 &dir, uint32_t &s1)
{
    e = p.pool[i];
    const uint32_t face = __float_as_uint(e.w);
    if (face >= 6u) return false;
    n = face_id_to_normal(face);
    x = mk3(e.x + 2.01f * n.x, e.y + 2.01f * n.y, e.z + 2.01f * n.z);
    s1 = pcg(pcg((i + 0x51F15EEDu) ^ p.key));
    dir = normalize3(rotate_by_normal(n, cosine_sample(s1)));
    return true;
}

// surfel.rgen
__global__ void __launch_bounds__(256) gi_gen_surfel(const __grid_constant__ GiParams p)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= p.pool_size) return;
    float4 e;
    f3 n, x, dir;
    uint32_t s1;
    if (!gi_surfel_setup(p, i, e, n, x, dir, s1)) { gi_dead_ray(p.q_any, i, i | 0x80000000u); gi_dead_ray(p.q_closest, i, i); return; }
    if (dot3(mk3(p.sky.direction[0], p.sky.direction[1], p.sky.direction[2]), n) > 0.0f) {
        p.q_any.a[i] = make_float4(x.x, x.y, x.z, 0.1f); p.q_any.b[i] = make_float4(p.sun_dir[0], p.sun_dir[1], p.sun_dir[2], 10000.0f);
        p.q_any.path[i] = i | 0x80000000u;
    } else gi_dead_ray(p.q_any, i, i | 0x80000000u);
    p.q_closest.a[i] = make_float4(x.x, x.y, x.z, 0.1f); p.q_closest.b[i] = make_float4(dir.x, dir.y, dir.z, 10000.0f);
    p.q_closest.path[i] = i;
}

// surfel/nee.rmiss + surfel.rchit / .rmiss: look-ups now, the insert (or the re-scheduling of the own slot) recorded for later
__global__ void __launch_bounds__(256) gi_surfel_shade(const __grid_constant__ GiParams p)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= p.pool_size) return;
    p.ins_loc[i] = 0xFFFFFFFFu;
    float4 e;
    f3 n, x, dir;
    uint32_t s1;
    if (!gi_surfel_setup(p, i, e, n, x, dir, s1)) return;
    atomicAdd(&p.stats[GI_SURFEL_RAYS], 1ull);
    f3 payload = mk3(0.0f, 0.0f, 0.0f);
    if (dot3(mk3(p.sky.direction[0], p.sky.direction[1], p.sky.direction[2]), n) > 0.0f && !p.vis[i]) {
        const float ndl = dot3(n, mk3(p.sun_dir[0], p.sun_dir[1], p.sun_dir[2]));
        payload = mk3(payload.x + p.sun_nee[0] * ndl, payload.y + p.sun_nee[1] * ndl, payload.z + p.sun_nee[2] * ndl);
    }
    CacheKey own;
    own.x = (int)(e.x / 4.0f); own.y = (int)(e.y / 4.0f); own.z = (int)(e.z / 4.0f);
    own.direction = __float_as_uint(e.w);
    const uint4 hw = p.hits[i];
    f3 value;
    if (hw.y == 0xFFFFFFFFu) {
        const f3 sk = sky_only(p, normalize3(dir));
        value = mk3(sk.x + payload.x, sk.y + payload.y, sk.z + payload.z);
    } else {
        const BlockHit b = gi_block_of_hit(p.scene, hw, x, dir);
        f3 radiance;
        uint32_t count;
        if (!cache_get(p.table, p.capacity, b.key, p.frame_index, radiance, count)) {
            const float prob = 1.0f / (float)(count + 2u);
            if (u01(pcg(s1 + 6u)) > prob) {
                p.pool[i] = make_float4(b.center.x, b.center.y, b.center.z, __uint_as_float(b.key.direction));   // index = i % SurfelPoolSize = i: the own slot
                atomicAdd(&p.stats[GI_SCHEDULED], 1ull);
            }
            return;
        }
        atomicAdd(&p.stats[GI_CACHE_FOUND], 1ull);
        const f3 out = gi_modulate(p.srgb_lut, b.avg_albedo, radiance);
        value = mk3(out.x + payload.x, out.y + payload.y, out.z + payload.z);
    }
    p.ins_loc[i] = cache_location(own, p.capacity);
    p.ins_fp[i] = cache_fingerprint(own);
    p.ins_val[i] = make_float4(value.x, value.y, value.z, 0.0f);
    atomicAdd(&p.stats[GI_INSERTS], 1ull);
}

// ---- the deterministic insert ----
__global__ void __launch_bounds__(256) gi_iota(uint32_t *__restrict__ a, uint32_t n)
{
    const uint32_t k = blockIdx.x * 256u + threadIdx.x;
    if (k < n) a[k] = k;
}
// after sort 1 (by location; surfel index as payload; invalid locations sort last): head flag of every probe-window group
__global__ void __launch_bounds__(256) gi_ins_flags(const uint32_t *__restrict__ loc_sorted, uint32_t n, uint32_t *__restrict__ flag)
{
    const uint32_t k = blockIdx.x * 256u + threadIdx.x;
    if (k >= n) return;
    const uint32_t l = loc_sorted[k];
    flag[k] = (l != 0xFFFFFFFFu && (k == 0u || l - loc_sorted[k - 1u] > 2u)) ? 1u : 0u;
}
// key of sort 2: (segment, surfel index); invalid entries stay last
__global__ void __launch_bounds__(256) gi_ins_keys(const uint32_t *__restrict__ loc_sorted, const uint32_t *__restrict__ idx_sorted, const uint32_t *__restrict__ seg,
                                                  uint32_t n, unsigned long long *__restrict__ key2)
{
    const uint32_t k = blockIdx.x * 256u + threadIdx.x;
    if (k >= n) return;
    key2[k] = loc_sorted[k] == 0xFFFFFFFFu ? ~0ull : (((unsigned long long)seg[k] << 32) | idx_sorted[k]);
}
// one thread per segment head: its inserts one after the other, ascending surfel index
__global__ void __launch_bounds__(128) gi_ins_apply(const __grid_constant__ GiParams p, const unsigned long long *__restrict__ key2, uint32_t n)
{
    const uint32_t k = blockIdx.x * 128u + threadIdx.x;
    if (k >= n) return;
    const unsigned long long me = key2[k];
    if (me == ~0ull) return;
    const uint32_t seg = (uint32_t)(me >> 32);
    if (k > 0u && (uint32_t)(key2[k - 1u] >> 32) == seg) return;                 // not a head
    atomicAdd(&p.stats[GI_SEGMENTS], 1ull);
    for (uint32_t j = k; j < n; j++) {
        const unsigned long long kj = key2[j];
        if ((uint32_t)(kj >> 32) != seg || kj == ~0ull) break;
        const uint32_t i = (uint32_t)kj;
        const float4 v = p.ins_val[i];
        cache_insert(p.table, p.ins_loc[i], p.ins_fp[i], mk3(v.x, v.y, v.z), p.frame_index);
    }
}

}  // namespace dust
