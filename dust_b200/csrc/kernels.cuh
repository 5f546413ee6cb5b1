// sm_100a device code, part 3: the kernels.
//   frame_kernel    per-pixel primary ray + sun next-event estimation + diffuse GI bounces + radiance
//                   (replaces the vkCmdTraceRaysKHR passes of StandardPipeline::render,
//                   crates/render/src/pipeline/standard.rs:477-490, 564-577, 627-640)
//   (explicit ray batches, BASELINE.json config 5, run on wavefront.cuh's wf_cull / wf_trace_pt / wf_trace)
//   resolve_kernel  un-permute gathered shard buffers to a row-major frame
// Reference sources restated in frame_kernel:
//   camera ray   assets/shaders/headers/camera.glsl:1-16, primary/primary.rgen:9-21
//   shading      primary/hit.rchit:16-95, primary/miss.rmiss:7-17,
//                final_gather/{ambient_occlusion.rgen, nee.rmiss, final_gather.rgen, .rmiss, .rchit}
//   final value  assets/shaders/tone_map.comp:192-203
#pragma once
#include "trace_core.cuh"

namespace dust {

#ifndef DUST_FRAME_MIN_BLOCKS
#define DUST_FRAME_MIN_BLOCKS 5          // frame_kernel: 96 registers, no spills
#endif
#ifndef DUST_MIN_BLOCKS
#define DUST_MIN_BLOCKS 7                // wf_trace: 72 registers (6 / 7 / 8 CTAs per SM: 2 939 / 2 971 / 2 952 Mrays/s on configs[2] once the dense root table had freed registers)
#endif

struct FrameParams {
    DSceneView scene;
    DustCamera cam;
    DustSky sky;
    float sun_dir[3];          // normalize(sky.direction.xyz)                          (ambient_occlusion.rgen:45)
    float sun_ar2;             // 1 / sin^2(solar radius)                               (sky.glsl:88-89, uniform-only)
    float sun_nee[3];          // sun_radiance(sun_dir) * (1 - cos(solar radius))        (nee.rmiss:18-21, uniform-only)
    DustTraceConfig cfg;
    uint32_t tiles_x, tiles_y, n_local_tiles;
    uint32_t slot_log2;        // 2^slot_log2 threads share a pixel, each takes every 2^slot_log2-th sample
    float4 *radiance32;
    __half *radiance16;
    uint32_t ycocg;            // radiance16 holds REBLUR_FrontEnd_PackRadianceAndNormHitDist's YCoCg (DUST_FMT_YCOCG16F)
    uint32_t *hit_instance, *hit_leaf;
    uint8_t *hit_voxel;
    float *hit_t;
    DustGBuffer gb;            // optional G-buffer of the sample-0 primary hit (NULL members are skipped)
    DCounters *counters;
    const uint32_t *tile_lists;   // per local tile: DUST_TILE_LIST_WORDS words (wf_tile_cull), or NULL
    // sun grid (wf_sun_grid): candidate instances of the sun shadow rays by where their origin projects along the sun direction
    const uint32_t *sun_lists;    // per cell: DUST_TILE_LIST_WORDS words, or NULL
    float sun_u[3], sun_v[3];     // orthonormal basis of the plane perpendicular to sun_dir
    float sun_u0, sun_v0, sun_inv_du, sun_inv_dv, sun_du, sun_dv, sun_eps;   // grid origin, 1 / cell size, cell size, conservative margin
};

#ifndef DUST_SUN_GRID
#define DUST_SUN_GRID 128                // cells per side of the sun grid
#endif

// Sun shadow rays all run along sun_dir, so a ray can only touch the instances whose box, projected along that direction, covers the
// projection of its origin: a 2D look-up replaces the BVH walk (the rays then run the exact box test on the listed instances, like
// the camera rays do on their tile's list).  Returns the cell's list, or NULL when the origin projects outside the grid (the ray walks the BVH).
__device__ __forceinline__ const uint32_t *sun_list_of(const FrameParams &f, f3 o)
{
    const float pu = (o.x * f.sun_u[0] + o.y * f.sun_u[1]) + o.z * f.sun_u[2];
    const float pv = (o.x * f.sun_v[0] + o.y * f.sun_v[1]) + o.z * f.sun_v[2];
    const float fu = (pu - f.sun_u0) * f.sun_inv_du, fv = (pv - f.sun_v0) * f.sun_inv_dv;
    if (!(fu >= 0.0f && fu < (float)DUST_SUN_GRID && fv >= 0.0f && fv < (float)DUST_SUN_GRID)) return nullptr;
    const uint32_t cell = (uint32_t)fv * DUST_SUN_GRID + (uint32_t)fu;
    return f.sun_lists + (size_t)cell * DUST_TILE_LIST_WORDS;
}

enum : int { RAY_NONE = 0, RAY_PRIMARY = 1, RAY_SHADOW = 2, RAY_BOUNCE = 3 };

// headers/camera.glsl:1-16
__device__ __forceinline__ void camera_ray(const DustCamera &cam, uint32_t width, uint32_t height, uint32_t px, uint32_t py,
                                           float jx, float jy, f3 &o, f3 &d)
{
    float ndc_x = ((float)px + jx) / (float)width;
    float ndc_y = ((float)py + jy) / (float)height;
    float cx = 2.0f * ndc_x - 1.0f;
    float cy = 2.0f * ndc_y - 1.0f;
    cy = cy * -1.0f;
    cx = cx * ((float)width / (float)height);
    cx = cx * cam.tan_half_fov;
    cy = cy * cam.tan_half_fov;
    d.x = (cam.camera_view_col0[0] * cx + cam.camera_view_col1[0] * cy) + cam.camera_view_col2[0] * -1.0f;
    d.y = (cam.camera_view_col0[1] * cx + cam.camera_view_col1[1] * cy) + cam.camera_view_col2[1] * -1.0f;
    d.z = (cam.camera_view_col0[2] * cx + cam.camera_view_col1[2] * cy) + cam.camera_view_col2[2] * -1.0f;
    o = mk3(cam.position_x, cam.position_y, cam.position_z);
}

// hit.rchit:16-76 for hit `h` of world ray (ro, rd): world normal -> sn, offset hit point
// (ambient_occlusion.rgen:23) -> ro, returns SRGBToLinear(albedo)
template <bool STATS>
__device__ __forceinline__ f3 surface_at(const DSceneView &sc, const Hit &h, f3 &ro, const f3 rd, f3 &sn, Counters &st, uint32_t *palette_index = nullptr)
{
    const DInstance &in = sc.inst[h.inst];
    f3 o, d;
    object_ray(in, ro, rd, o, d);
    const uint32_t vox = h.voxel;
    const int lx = (int)(h.pos & 1023u) << 2, ly = (int)((h.pos >> 10) & 1023u) << 2, lz = (int)(h.pos >> 20) << 2;
    const float t = h.t;
    const float vx = (t * d.x + o.x) - (((float)lx + (float)((vox >> 4) & 3u)) + 0.5f);
    const float vy = (t * d.y + o.y) - (((float)ly + (float)((vox >> 2) & 3u)) + 0.5f);
    const float vz = (t * d.z + o.z) - (((float)lz + (float)(vox & 3u)) + 0.5f);
    const float ax = fabsf(vx), ay = fabsf(vy), az = fabsf(vz);
    const float mx = fmaxf(ax, fmaxf(ay, az));                                     // normal.glsl:39-43
    const f3 n_obj = mk3(signf(vx) * (ax >= mx ? 1.0f : 0.0f), signf(vy) * (ay >= mx ? 1.0f : 0.0f), signf(vz) * (az >= mx ? 1.0f : 0.0f));
    const float *W = in.o2w;
    sn = mk3((W[0] * n_obj.x + W[1] * n_obj.y) + W[2] * n_obj.z,
             (W[3] * n_obj.x + W[4] * n_obj.y) + W[5] * n_obj.z,
             (W[6] * n_obj.x + W[7] * n_obj.y) + W[8] * n_obj.z);
    const uint4 leaf = __ldg(&in.leaves[h.leaf]);
    const uint32_t off = rank64(leaf.x, leaf.y, vox);                              // hit.rchit:60-61
    const uint8_t pi = __ldg(&in.materials[leaf.z + off]);                         // :71
    const float4 c = __ldg(&in.palette_lin[pi]);                                   // :72-74 + SRGBToLinear
    if (palette_index) *palette_index = pi;
    ro = mk3((t * rd.x + ro.x) + sn.x * 0.01f, (t * rd.y + ro.y) + sn.y * 0.01f, (t * rd.z + ro.z) + sn.z * 0.01f);
    if (STATS) st.hits_shaded++;
    return mk3(c.x, c.y, c.z);
}

// Pixel store shared by both pipelines: RGBA32F / RGBA16F ACEScg, or the `img_illuminance` encoding the
// reference's post chain reads (headers/nrd.glsl:99-147: Y, Co, Cg; normHitDist is NRD-only and stored as 1).
__device__ __forceinline__ void store_radiance(float4 *radiance32, __half *radiance16, uint32_t ycocg, uint32_t index, float r, float g, float b)
{
    if (radiance32) radiance32[index] = make_float4(r, g, b, 1.0f);
    if (radiance16) {
        if (ycocg) {                                                               // _NRD_LinearToYCoCg
            const float Y = (r * 0.25f + g * 0.5f) + b * 0.25f;
            const float Co = (r * 0.5f + g * 0.0f) + b * -0.5f;
            const float Cg = (r * -0.25f + g * 0.5f) + b * -0.25f;
            r = Y; g = Co; b = Cg;
        }
        const __half2 lo = __floats2half2_rn(r, g), hi = __floats2half2_rn(b, 1.0f);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t *>(&lo); pk.y = *reinterpret_cast<const uint32_t *>(&hi);
        reinterpret_cast<uint2 *>(radiance16)[index] = pk;
    }
}

// ------------------------------------------------------------------------------------------ G-buffer

// float -> UNORM (Vulkan "round to nearest"): rint(clamp(x, 0, 1) * (2^bits - 1))
__device__ __forceinline__ uint32_t unorm(float x, float maxv) { return (uint32_t)rintf(fminf(fmaxf(x, 0.0f), 1.0f) * maxv); }
// A2R10G10B10_UNORM_PACK32: A 31..30, R 29..20, G 19..10, B 9..0
__device__ __forceinline__ uint32_t pack_a2r10g10b10(float r, float g, float b, float a)
{
    return (unorm(a, 3.0f) << 30) | (unorm(r, 1023.0f) << 20) | (unorm(g, 1023.0f) << 10) | unorm(b, 1023.0f);
}

// The image stores of hit.rchit:57-94 (hit) / miss.rmiss:13-16 (miss) for pixel `pi`.  (ro, rd) is the
// world-space primary ray BEFORE surface_at offsets its origin; sn the world normal surface_at returned.
__device__ __forceinline__ void gbuffer_store(const DustGBuffer &gb, const DSceneView &sc, uint32_t pi, bool hit, float t, uint32_t inst,
                                              uint32_t voxel, uint32_t palette_index, f3 ro, f3 rd, f3 sn)
{
    if (!hit) {
        if (gb.albedo) gb.albedo[pi] = 0xFFFFFFFFu;                                // vec4(1.0)
        if (gb.depth) gb.depth[pi] = CUDART_INF_F;                                 // 1.0 / 0.0
        if (gb.motion) reinterpret_cast<uint2 *>(gb.motion)[pi] = make_uint2(0u, 0u);
        return;
    }
    const DInstance &in = sc.inst[inst];
    if (gb.albedo) {
        const uchar4 c = __ldg(&in.palette_raw[palette_index]);                    // hit.rchit:72-76
        gb.albedo[pi] = pack_a2r10g10b10((float)c.x / 255.0f, (float)c.y / 255.0f, (float)c.z / 255.0f, 1.0f);
    }
    if (gb.depth) gb.depth[pi] = t;                                                // :81
    if (gb.normal) {
        // NRD_FrontEnd_PackNormalAndRoughness(normalWorld, 1.0, palette_index), headers/nrd.glsl:1-52 (R10G10B10A2 encoding)
        const float s = (fabsf(sn.x) + fabsf(sn.y)) + fabsf(sn.z);
        const float vx = sn.x / s, vy = sn.y / s, vz = sn.z / s;
        const float wx = (1.0f - fabsf(vy)) * ((vx < 0.0f ? 0.0f : 1.0f) * 2.0f - 1.0f);
        const float wy = (1.0f - fabsf(vx)) * ((vy < 0.0f ? 0.0f : 1.0f) * 2.0f - 1.0f);
        const float ex = vz >= 0.0f ? vx : wx, ey = vz >= 0.0f ? vy : wy;
        gb.normal[pi] = pack_a2r10g10b10(ex * 0.5f + 0.5f, ey * 0.5f + 0.5f, 1.0f, fminf(fmaxf((float)palette_index / 3.0f, 0.0f), 1.0f));
    }
    if (gb.voxel_id) gb.voxel_id[pi] = (voxel << 24) | (inst & 0xFFFFu) | (palette_index << 16);   // :87
    if (gb.motion) {
        // :90-94: world hit point -> model space -> last frame's world position
        const f3 hw = mk3(t * rd.x + ro.x, t * rd.y + ro.y, t * rd.z + ro.z);
        const float4 r0 = __ldg(reinterpret_cast<const float4 *>(in.w2o)), r1 = __ldg(reinterpret_cast<const float4 *>(in.w2o) + 1),
                     r2 = __ldg(reinterpret_cast<const float4 *>(in.w2o) + 2);
        const f3 hm = mk3(((r0.x * hw.x + r0.y * hw.y) + r0.z * hw.z) + r0.w, ((r1.x * hw.x + r1.y * hw.y) + r1.z * hw.z) + r1.w,
                          ((r2.x * hw.x + r2.y * hw.y) + r2.z * hw.z) + r2.w);
        const float4 p0 = __ldg(&sc.prev_o2w[3 * inst]), p1 = __ldg(&sc.prev_o2w[3 * inst + 1]), p2 = __ldg(&sc.prev_o2w[3 * inst + 2]);
        const f3 last = mk3(((p0.x * hm.x + p0.y * hm.y) + p0.z * hm.z) + p0.w, ((p1.x * hm.x + p1.y * hm.y) + p1.z * hm.z) + p1.w,
                            ((p2.x * hm.x + p2.y * hm.y) + p2.z * hm.z) + p2.w);
        const __half2 a = __floats2half2_rn(last.x - hw.x, last.y - hw.y), b = __floats2half2_rn(last.z - hw.z, 0.0f);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t *>(&a); pk.y = *reinterpret_cast<const uint32_t *>(&b);
        reinterpret_cast<uint2 *>(gb.motion)[pi] = pk;
    }
}

__device__ __forceinline__ bool gbuffer_wanted(const DustGBuffer &gb)
{
    return gb.albedo || gb.depth || gb.normal || gb.motion || gb.voxel_id;
}

__device__ __forceinline__ void flush_counters(DCounters *g, const Counters &st)
{
    // warp-aggregate, one atomic per counter per warp
    uint32_t v[23] = { st.primary_rays, st.bounce_rays, st.shadow_rays, st.mask_words, st.child_ptrs, st.hits_shaded };
#pragma unroll
    for (int i = 0; i < 16; i++) v[6 + i] = st.dbg[i];
    v[22] = st.mask_words_strict;
    unsigned long long *dst = reinterpret_cast<unsigned long long *>(g);
#pragma unroll
    for (int i = 0; i < 23; i++) {
        unsigned long long s = v[i];
        for (int ofs = 16; ofs > 0; ofs >>= 1) s += __shfl_xor_sync(0xffffffffu, s, ofs);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(&dst[i], s);
    }
}

__device__ __forceinline__ uint32_t even_bits(uint32_t v)      // bits 0,2,4,6,8 -> 0..4
{
    v &= 0x155u;
    v = (v | (v >> 1)) & 0x133u;
    v = (v | (v >> 2)) & 0x10Fu;
    v = (v | (v >> 4)) & 0x01Fu;
    return v;
}

// One warp per local tile: the instances whose padded world box can be touched by ANY camera ray of the tile (any jitter), by a
// conservative frustum test — four planes through the camera position spanned by the tile's corner directions (camera.glsl:1-16 is linear
// in the pixel coordinate), widened by a relative epsilon; a box is dropped only when all eight corners lie outside one plane.
// list[0] = count (0xFFFFFFFF: more than DUST_TILE_LIST_MAX, the rays walk the BVH), list[1..] = instance indices.
__global__ void __launch_bounds__(128) wf_tile_cull(const __grid_constant__ FrameParams p, uint32_t *__restrict__ lists)
{
    const unsigned full = 0xffffffffu;
    const uint32_t lt = blockIdx.x * 4u + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
    if (lt >= p.n_local_tiles) return;
    uint32_t *list = lists + (size_t)lt * DUST_TILE_LIST_WORDS;
    const DustTraceConfig &cfg = p.cfg;
    const uint32_t t = lt * cfg.shard_count + cfg.shard_index;
    const uint32_t tile_x = t % p.tiles_x, tile_y = t / p.tiles_x;
    if (tile_y >= p.tiles_y) { if (lane == 0) list[0] = 0u; return; }
    // pixel coordinate + jitter spans [x0, x0 + 32] x [y0, y0 + 32]
    const float W = (float)cfg.width, H = (float)cfg.height, tanf_ = p.cam.tan_half_fov;
    const float xa = (float)(tile_x * 32u), xb = xa + 32.0f, ya = (float)(tile_y * 32u), yb = ya + 32.0f;
    float cx0 = ((2.0f * (xa / W) - 1.0f) * (W / H)) * tanf_, cx1 = ((2.0f * (xb / W) - 1.0f) * (W / H)) * tanf_;
    float cy0 = ((2.0f * (ya / H) - 1.0f) * -1.0f) * tanf_, cy1 = ((2.0f * (yb / H) - 1.0f) * -1.0f) * tanf_;
    if (cx0 > cx1) { const float s_ = cx0; cx0 = cx1; cx1 = s_; }
    if (cy0 > cy1) { const float s_ = cy0; cy0 = cy1; cy1 = s_; }
    const float ex = 1e-4f * (fabsf(cx0) + fabsf(cx1) + 1.0f), ey = 1e-4f * (fabsf(cy0) + fabsf(cy1) + 1.0f);
    cx0 -= ex; cx1 += ex; cy0 -= ey; cy1 += ey;
    const f3 c0 = mk3(p.cam.camera_view_col0[0], p.cam.camera_view_col0[1], p.cam.camera_view_col0[2]);
    const f3 c1 = mk3(p.cam.camera_view_col1[0], p.cam.camera_view_col1[1], p.cam.camera_view_col1[2]);
    const f3 c2 = mk3(p.cam.camera_view_col2[0], p.cam.camera_view_col2[1], p.cam.camera_view_col2[2]);
    auto dir = [&](float cx, float cy) { return mk3(c0.x * cx + c1.x * cy - c2.x, c0.y * cx + c1.y * cy - c2.y, c0.z * cx + c1.z * cy - c2.z); };
    auto cross = [](f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); };
    const f3 d00 = dir(cx0, cy0), d10 = dir(cx1, cy0), d11 = dir(cx1, cy1), d01 = dir(cx0, cy1), dc = dir(0.5f * (cx0 + cx1), 0.5f * (cy0 + cy1));
    f3 n[4] = { cross(d00, d10), cross(d10, d11), cross(d11, d01), cross(d01, d00) };
    float nl[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (dot3(n[k], dc) > 0.0f) n[k] = mk3(-n[k].x, -n[k].y, -n[k].z);        // inside: dot(n, v) <= 0
        nl[k] = sqrtf(dot3(n[k], n[k]));
    }
    const f3 o = mk3(p.cam.position_x, p.cam.position_y, p.cam.position_z);
    uint32_t count = 0;
    for (uint32_t base = 0; base < p.scene.n_inst; base += 32u) {
        const uint32_t inst = base + lane;
        bool keep = false;
        if (inst < p.scene.n_inst) {
            const float4 lo = __ldg(&p.scene.inst_boxes[2u * inst]), hi = __ldg(&p.scene.inst_boxes[2u * inst + 1u]);
            const float far_ = fmaxf(fmaxf(fmaxf(fabsf(lo.x - o.x), fabsf(hi.x - o.x)), fmaxf(fabsf(lo.y - o.y), fabsf(hi.y - o.y))), fmaxf(fabsf(lo.z - o.z), fabsf(hi.z - o.z)));
            keep = true;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                // the corner that lies deepest inside (smallest dot): if even that one is outside by more than the margin, all eight are
                const float vx = (n[k].x > 0.0f ? lo.x : hi.x) - o.x, vy = (n[k].y > 0.0f ? lo.y : hi.y) - o.y, vz = (n[k].z > 0.0f ? lo.z : hi.z) - o.z;
                const float dmin = (n[k].x * vx + n[k].y * vy) + n[k].z * vz;
                if (dmin > 1e-4f * nl[k] * (far_ * 3.0f + 1.0f)) keep = false;
            }
        }
        const unsigned m = __ballot_sync(full, keep);
        if (keep) {
            const uint32_t k = count + (uint32_t)__popc(m & ((1u << lane) - 1u));
            if (k < DUST_TILE_LIST_MAX) list[1u + k] = inst;
        }
        count += (uint32_t)__popc(m);
    }
    if (lane == 0) list[0] = count > DUST_TILE_LIST_MAX ? 0xFFFFFFFFu : count;
}

// One warp per cell of the sun grid: the instances whose padded world box, projected onto the plane perpendicular to the sun direction
// (centre +- the box's half extents weighted by |u|, |v|: the bounding rectangle of the projected box), overlaps the cell widened by
// sun_eps (rounding of the rays' own projection, of the cell index, and of this test).  list[0] = count (0xFFFFFFFF: more than
// DUST_TILE_LIST_MAX, the rays of that cell walk the BVH), list[1..] = instance indices.
__global__ void __launch_bounds__(128) wf_sun_grid(const __grid_constant__ FrameParams p, uint32_t *__restrict__ lists)
{
    const unsigned full = 0xffffffffu;
    const uint32_t cell = blockIdx.x * 4u + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
    if (cell >= DUST_SUN_GRID * DUST_SUN_GRID) return;
    uint32_t *list = lists + (size_t)cell * DUST_TILE_LIST_WORDS;
    const uint32_t cu = cell % DUST_SUN_GRID, cv = cell / DUST_SUN_GRID;
    const float ulo = p.sun_u0 + (float)cu * p.sun_du - p.sun_eps, uhi = p.sun_u0 + (float)(cu + 1u) * p.sun_du + p.sun_eps;
    const float vlo = p.sun_v0 + (float)cv * p.sun_dv - p.sun_eps, vhi = p.sun_v0 + (float)(cv + 1u) * p.sun_dv + p.sun_eps;
    uint32_t count = 0;
    for (uint32_t base = 0; base < p.scene.n_inst; base += 32u) {
        const uint32_t inst = base + lane;
        bool keep = false;
        if (inst < p.scene.n_inst) {
            const float4 lo = __ldg(&p.scene.inst_boxes[2u * inst]), hi = __ldg(&p.scene.inst_boxes[2u * inst + 1u]);
            const float cx = 0.5f * (lo.x + hi.x), cy = 0.5f * (lo.y + hi.y), cz = 0.5f * (lo.z + hi.z);
            const float hx = 0.5f * (hi.x - lo.x), hy = 0.5f * (hi.y - lo.y), hz = 0.5f * (hi.z - lo.z);
            const float pu = (cx * p.sun_u[0] + cy * p.sun_u[1]) + cz * p.sun_u[2], ru = (hx * fabsf(p.sun_u[0]) + hy * fabsf(p.sun_u[1])) + hz * fabsf(p.sun_u[2]);
            const float pv = (cx * p.sun_v[0] + cy * p.sun_v[1]) + cz * p.sun_v[2], rv = (hx * fabsf(p.sun_v[0]) + hy * fabsf(p.sun_v[1])) + hz * fabsf(p.sun_v[2]);
            keep = pu - ru <= uhi && pu + ru >= ulo && pv - rv <= vhi && pv + rv >= vlo;
        }
        const unsigned m = __ballot_sync(full, keep);
        if (keep) {
            const uint32_t k = count + (uint32_t)__popc(m & ((1u << lane) - 1u));
            if (k < DUST_TILE_LIST_MAX) list[1u + k] = inst;
        }
        count += (uint32_t)__popc(m);
    }
    if (lane == 0) list[0] = count > DUST_TILE_LIST_MAX ? 0xFFFFFFFFu : count;
}

// Thread mapping: 2^slot_log2 consecutive threads share a pixel (one sample each per round); pixels
// follow a Morton curve inside 32x32 tiles, so a warp covers a compact patch (4x2 pixels x 4 samples
// at 4 spp, 8x4 pixels at 1 spp).  32x32 tiles are the sharding unit: tile t -> shard t % shard_count.
// Every thread runs ONE loop around ONE call of trace_scene: the path's next ray (primary, sun
// shadow, bounce) is generated after the previous one is consumed.
template <int NL, bool STATS>
__global__ void __launch_bounds__(DUST_THREADS, DUST_FRAME_MIN_BLOCKS) frame_kernel(const __grid_constant__ FrameParams p)
{
    __shared__ uint32_t s_scratch[DUST_SCRATCH_WORDS * DUST_THREADS];
    uint32_t *sm = s_scratch + threadIdx.x;
    const unsigned full = 0xffffffffu;
    const DustTraceConfig &cfg = p.cfg;
    const uint32_t item = blockIdx.x * DUST_THREADS + threadIdx.x;
    const uint32_t S = 1u << p.slot_log2, slot = item & (S - 1u), pix = item >> p.slot_log2;
    const uint32_t lt = pix >> 10, q = pix & 1023u;
    const uint32_t lx = even_bits(q), ly = even_bits(q >> 1);
    const uint32_t t = lt * cfg.shard_count + cfg.shard_index;
    const uint32_t tile_x = t % p.tiles_x, tile_y = t / p.tiles_x;
    const uint32_t px = tile_x * 32u + lx, py = tile_y * 32u + ly;
    const bool have_pixel = lt < p.n_local_tiles && tile_y < p.tiles_y && px < cfg.width && py < cfg.height;
    const uint32_t pixel_index = py * cfg.width + px;
    const uint32_t s0 = pcg(pixel_index ^ pcg(cfg.seed + 0x9E3779B9u * cfg.frame_index));
    const f3 sun_dir = mk3(p.sun_dir[0], p.sun_dir[1], p.sun_dir[2]);
    Counters st = {};
    f3 acc = mk3(0.0f, 0.0f, 0.0f);

#pragma unroll 1
    for (uint32_t smp0 = 0; smp0 < cfg.spp; smp0 += S) {
        const uint32_t smp = smp0 + slot;
        int kind = (have_pixel && smp < cfg.spp) ? RAY_PRIMARY : RAY_NONE;
        const uint32_t s1 = pcg(s0 + smp);
        f3 ro, rd, sn = mk3(0.0f, 0.0f, 0.0f);
        {
            float jx = 0.5f, jy = 0.5f;
            if (smp > 0) { jx = u01(pcg(s1 + 0u)); jy = u01(pcg(s1 + 1u)); }
            camera_ray(p.cam, cfg.width, cfg.height, px, py, jx, jy, ro, rd);          // primary.rgen:9-21
        }
        float tmin = p.cam.near, tmax = p.cam.far;
        f3 L = mk3(0.0f, 0.0f, 0.0f), T = mk3(1.0f, 1.0f, 1.0f);                      // linear sRGB primaries
        uint32_t v = 0;

        // the first ray of every sample is the camera ray of a pixel of this CTA's tile: the tile's candidate list replaces its BVH walk
        const uint32_t *tl = p.tile_lists ? p.tile_lists + (size_t)min(lt, p.n_local_tiles - 1u) * DUST_TILE_LIST_WORDS : nullptr;
#pragma unroll 1
        while (__any_sync(full, kind != RAY_NONE)) {
            Hit h;
            trace_scene<NL, STATS>(p.scene, ro, rd, tmin, tmax, kind == RAY_SHADOW, kind != RAY_NONE, sm, h, st, LEAF_PRECISE, tl);
            tl = nullptr;

            bool to_vertex = false, to_bounce = false, need_sky = false;
            if (kind == RAY_SHADOW) {
                if (!h.hit) {                                                          // nee.rmiss:11-22
                    const float ndl = dot3(sn, sun_dir);
                    const f3 ts = acescg2srgb(mk3(p.sun_nee[0] * ndl, p.sun_nee[1] * ndl, p.sun_nee[2] * ndl));
                    L.x += T.x * ts.x; L.y += T.y * ts.y; L.z += T.z * ts.z;
                }
                to_bounce = true;
            } else if (kind != RAY_NONE) {
                if (STATS) { if (kind == RAY_PRIMARY) st.primary_rays++; else st.bounce_rays++; }
                if (kind == RAY_PRIMARY && smp == 0 && (cfg.flags & DUST_TRACE_WRITE_HITS) && p.hit_instance) {
                    p.hit_instance[pixel_index] = h.hit ? h.inst : 0xFFFFFFFFu; p.hit_leaf[pixel_index] = h.hit ? h.leaf : 0xFFFFFFFFu;
                    p.hit_voxel[pixel_index] = h.hit ? (uint8_t)h.voxel : (uint8_t)0xFF; p.hit_t[pixel_index] = h.hit ? h.t : CUDART_INF_F;
                }
                const bool want_gb = kind == RAY_PRIMARY && smp == 0 && gbuffer_wanted(p.gb);
                const f3 ro0 = ro;
                if (want_gb && !h.hit) gbuffer_store(p.gb, p.scene, pixel_index, false, 0.0f, 0u, 0u, 0u, ro0, rd, sn);
                if (h.hit) {
                    uint32_t pidx = 0;
                    const f3 a = surface_at<STATS>(p.scene, h, ro, rd, sn, st, &pidx);
                    if (want_gb) gbuffer_store(p.gb, p.scene, pixel_index, true, h.t, h.inst, h.voxel, pidx, ro0, rd, sn);
                    if (kind == RAY_PRIMARY) {                                         // tone_map.comp:195-203 (the albedo factor;
                        if (!(cfg.flags & DUST_TRACE_DEMODULATE)) T = a;               //  left to tone_map when demodulated)
                        v = 0;
                    }
                    else { T.x = T.x * a.x; T.y = T.y * a.y; T.z = T.z * a.z; v += 1; }   // final_gather.rchit:76-80
                    to_vertex = true;
                } else need_sky = true;
            }
            if (need_sky) {
                const f3 dn = normalize3(rd);
                f3 sk, su;
                sky_eval(p.sky, p.sun_ar2, dn.x, dn.y, dn.z, kind == RAY_PRIMARY, sk.x, sk.y, sk.z, su.x, su.y, su.z);
                if (kind == RAY_PRIMARY) {                                             // miss.rmiss:8-14
                    L = acescg2srgb(mk3((sk.x + su.x) / 3.14f, (sk.y + su.y) / 3.14f, (sk.z + su.z) / 3.14f));
                } else {                                                               // final_gather.rmiss:13-17
                    const f3 ts = acescg2srgb(sk);
                    L.x += T.x * ts.x; L.y += T.y * ts.y; L.z += T.z * ts.z;
                }
                kind = RAY_NONE;
            }
            if (to_vertex) {
                if (dot3(mk3(p.sky.direction[0], p.sky.direction[1], p.sky.direction[2]), sn) > 0.0f) {   // ambient_occlusion.rgen:33-50
                    if (STATS) st.shadow_rays++;
                    rd = sun_dir; tmin = 0.1f; tmax = 10000.0f;
                    kind = RAY_SHADOW;
                } else to_bounce = true;
            }
            if (to_bounce) {
                if (v == cfg.max_bounces) kind = RAY_NONE;
                else {
                    const float u1 = u01(pcg(s1 + 2u + 2u * v)), u2 = u01(pcg(s1 + 3u + 2u * v));   // final_gather.rgen:32-34
                    const float r = sqrtf(u1), cz = sqrtf(1.0f - u1);
                    float c, sn_;
                    cossin_2pi(u2, c, sn_);
                    rd = normalize3(rotate_by_normal(sn, mk3(r * c, r * sn_, cz)));
                    tmin = 0.1f; tmax = p.cam.far;
                    kind = RAY_BOUNCE;
                }
            }
        }

        // pixel = mean over samples, accumulated in sample order (like the oracle's loop) by every thread of the pixel
        const f3 Lc = srgb2acescg(L);
        const unsigned base = (threadIdx.x & 31u) & ~(S - 1u);
        for (uint32_t s = 0; s < S; s++) {
            const float x = __shfl_sync(full, Lc.x, base + s), y = __shfl_sync(full, Lc.y, base + s), z = __shfl_sync(full, Lc.z, base + s);
            if (smp0 + s < cfg.spp) { acc.x += x; acc.y += y; acc.z += z; }
        }
    }

    if (have_pixel && slot == 0) {
        const uint32_t out_index = cfg.shard_count > 1 ? lt * 1024u + ly * 32u + lx : pixel_index;
        const float n = (float)cfg.spp;
        store_radiance(p.radiance32, p.radiance16, p.ycocg, out_index, acc.x / n, acc.y / n, acc.z / n);
    }
    if (STATS) flush_counters(p.counters, st);
}

// un-permute gathered shard buffers to a row-major frame (runs after the all-gather)
__global__ void __launch_bounds__(256) resolve_kernel(const float4 *__restrict__ gathered, float4 *__restrict__ out,
                                                      uint32_t width, uint32_t height, uint32_t shard_count,
                                                      uint32_t tiles_x, uint32_t tiles_per_shard)
{
    const uint32_t x = blockIdx.x * 32 + (threadIdx.x & 31);
    const uint32_t y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= width || y >= height) return;
    const uint32_t t = (y >> 5) * tiles_x + (x >> 5);
    const uint32_t shard = t % shard_count, lt = t / shard_count;
    const size_t src = ((size_t)shard * tiles_per_shard + lt) * 1024u + (y & 31u) * 32u + (x & 31u);
    out[(size_t)y * width + x] = __ldg(&gathered[src]);
}

}  // namespace dust
