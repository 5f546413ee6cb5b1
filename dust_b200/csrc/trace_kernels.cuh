// sm_100a device code of the hot path: camera rays, closest-hit / any-hit traversal of the
// uploaded trees, per-hit radiance.  Built with -fmad=false: every float expression is evaluated
// exactly as written (explicit __fmaf_rn only where DESIGN.md's traversal spec says "fma"), so hit
// ids and t are reproducible bit-for-bit.  Reference sources restated here:
//   camera ray   assets/shaders/headers/camera.glsl:1-16, primary/primary.rgen:9-21
//   leaf step    assets/shaders/primary/hit.rint:20-28, 43-131
//   shading      primary/hit.rchit:16-95, primary/miss.rmiss:7-17, headers/{normal,color,sky}.glsl,
//                final_gather/{ambient_occlusion.rgen, nee.rmiss, final_gather.rgen, .rmiss, .rchit}
//   final value  assets/shaders/tone_map.comp:192-203
#pragma once
#include "device_layout.h"
#include "../../include/dust_b200.h"
#include <cuda_fp16.h>
#include <math_constants.h>

namespace dust {

struct f3 { float x, y, z; };

__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ float dot3(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ f3 normalize3(f3 v)
{
    float l = sqrtf(dot3(v, v));
    return mk3(v.x / l, v.y / l, v.z / l);
}
__device__ __forceinline__ float signf(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }

// headers/spatial_hash.glsl:105-111
__device__ __forceinline__ uint32_t pcg(uint32_t v)
{
    uint32_t state = v * 747796405u + 2891336453u;
    uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}
__device__ __forceinline__ float u01(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }

// cos/sin(2*pi*u) from mul/add polynomials only (bit-reproducible; see DESIGN.md "PRNG and sampling")
__device__ __forceinline__ void cossin_2pi(float u, float &c, float &s)
{
    float x = u * 4.0f;
    float qf = floorf(x);
    int q = (int)qf & 3;
    float th = ((x - qf) - 0.5f) * 1.57079632679489661923f;
    float t2 = th * th;
    float sn = th * (1.0f + t2 * (-1.0f / 6.0f + t2 * (1.0f / 120.0f + t2 * (-1.0f / 5040.0f + t2 * (1.0f / 362880.0f)))));
    float cs = 1.0f + t2 * (-0.5f + t2 * (1.0f / 24.0f + t2 * (-1.0f / 720.0f + t2 * (1.0f / 40320.0f + t2 * (-1.0f / 3628800.0f)))));
    float c45 = (cs - sn) * 0.70710678118654752440f;
    float s45 = (cs + sn) * 0.70710678118654752440f;
    if (q == 0)      { c = c45;  s = s45; }
    else if (q == 1) { c = -s45; s = c45; }
    else if (q == 2) { c = -c45; s = -s45; }
    else             { c = s45;  s = -c45; }
}

// headers/normal.glsl:31-37
__device__ __forceinline__ f3 rotate_by_normal(f3 n, f3 v)
{
    float q0 = -n.y, q1 = n.x, q2 = 0.0f, q3 = 1.0f + n.z;
    float l = sqrtf(((q0 * q0 + q1 * q1) + q2 * q2) + q3 * q3);
    q0 = q0 / l; q1 = q1 / l; q2 = q2 / l; q3 = q3 / l;
    if (n.z < -0.99999f) { q0 = -1.0f; q1 = 0.0f; q2 = 0.0f; q3 = 0.0f; }
    f3 q = mk3(q0, q1, q2);
    float d = dot3(q, v);
    float qq = dot3(q, q);
    f3 cr = mk3(q.y * v.z - q.z * v.y, q.z * v.x - q.x * v.z, q.x * v.y - q.y * v.x);
    float k0 = 2.0f * d, k1 = q3 * q3 - qq, k2 = 2.0f * q3;
    return mk3((k0 * q.x + k1 * v.x) + k2 * cr.x, (k0 * q.y + k1 * v.y) + k2 * cr.y, (k0 * q.z + k1 * v.z) + k2 * cr.z);
}

// ------------------------------------------------------------------------------------ colour / sky

// GLSL mat3(c0,c1,c2) * v
__device__ __forceinline__ f3 mat3_mul(const float *m, f3 v)
{
    return mk3((m[0] * v.x + m[3] * v.y) + m[6] * v.z, (m[1] * v.x + m[4] * v.y) + m[7] * v.z, (m[2] * v.x + m[5] * v.y) + m[8] * v.z);
}
// headers/color.glsl:8-31
__device__ __forceinline__ f3 srgb2acescg(f3 v)
{
    const float m[9] = { 0.6031065f, 0.07011794f, 0.022178888f, 0.32633433f, 0.9199162f, 0.11607823f, 0.047995567f, 0.012763573f, 0.94101846f };
    return mat3_mul(m, v);
}
__device__ __forceinline__ f3 acescg2srgb(f3 v)
{
    const float m[9] = { 1.7312546f, -0.131619f, -0.024568284f, -0.6040432f, 1.1348418f, -0.12575036f, -0.08010775f, -0.008679431f, 1.0656371f };
    return mat3_mul(m, v);
}
__device__ __forceinline__ f3 xyz2acescg(f3 v)
{
    const float m[9] = { 1.6410228f, -0.66366285f, 0.011721907f, -0.32480323f, 1.6153315f, -0.0082844375f, -0.23642465f, 0.016756356f, 0.9883947f };
    return mat3_mul(m, v);
}

// headers/sky.glsl:1-16
__device__ __forceinline__ float sky_channel(const DustSkyChannel &ch, float cos_theta, float gamma, float cos_gamma)
{
    const float *c = ch.config;
    float expM = expf(c[4] * gamma);
    float rayM = cos_gamma * cos_gamma;
    float mieM = (1.0f + rayM) / powf((1.0f + c[8] * c[8]) - (2.0f * c[8]) * cos_gamma, 1.5f);
    float zenith = sqrtf(cos_theta);
    return (1.0f + c[0] * expf(c[1] / (cos_theta + 0.01f))) *
           ((((c[2] + c[3] * expM) + c[5] * rayM) + c[6] * mieM) + c[7] * zenith);
}

// headers/sky.glsl:19-78 (cos_gamma clamped before acos, like the oracle)
__device__ __forceinline__ f3 sky_radiance(const DustSky &sky, f3 dir)
{
    if (sky.direction[1] <= 0.0f) return mk3(0.0f, 0.0f, 0.0f);
    float cos_theta = fminf(fmaxf(dir.y, 0.0f), 1.0f);
    float cos_gamma = (dir.x * sky.direction[0] + dir.y * sky.direction[1]) + dir.z * sky.direction[2];
    float gamma = acosf(fminf(fmaxf(cos_gamma, -1.0f), 1.0f));
    f3 xyz;
    xyz.x = sky_channel(sky.r, cos_theta, gamma, cos_gamma) * sky.r.radiance;
    xyz.y = sky_channel(sky.g, cos_theta, gamma, cos_gamma) * sky.g.radiance;
    xyz.z = sky_channel(sky.b, cos_theta, gamma, cos_gamma) * sky.b.radiance;
    xyz.x = xyz.x * 683.0f; xyz.y = xyz.y * 683.0f; xyz.z = xyz.z * 683.0f;
    return xyz2acescg(xyz);
}

// headers/sky.glsl:80-113; ar2 = 1/sin^2(solar radius) is a uniform-only term, hoisted to the host
__device__ __forceinline__ f3 sun_radiance(const DustSky &sky, float ar2, f3 dir)
{
    float cos_gamma = (dir.x * sky.direction[0] + dir.y * sky.direction[1]) + dir.z * sky.direction[2];
    if (cos_gamma < 0.0f || dir.y < 0.0f) return mk3(0.0f, 0.0f, 0.0f);
    float singamma = 1.0f - (cos_gamma * cos_gamma);
    float sc2 = 1.0f - ar2 * singamma * singamma;
    if (sc2 <= 0.0f) return mk3(0.0f, 0.0f, 0.0f);
    float sc = sqrtf(sc2);
    f3 dark = mk3(sky.r.ld_coefficient0, sky.g.ld_coefficient0, sky.b.ld_coefficient0);
    dark.x += sky.r.ld_coefficient1 * sc; dark.y += sky.g.ld_coefficient1 * sc; dark.z += sky.b.ld_coefficient1 * sc;
    float cur = sc;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        cur *= sc;
        dark.x += sky.r.ld_coefficient2[i] * cur; dark.y += sky.g.ld_coefficient2[i] * cur; dark.z += sky.b.ld_coefficient2[i] * cur;
    }
    return xyz2acescg(mk3(sky.sunlight_intensity[0] * dark.x, sky.sunlight_intensity[1] * dark.y, sky.sunlight_intensity[2] * dark.z));
}

// --------------------------------------------------------------------------------------- LEAF STEP

// assets/shaders/primary/hit.rint:43-131 (scale = 1).  origin is leaf-local.  Returns true with
// (t, voxelId) when the DDA reaches a set bit; the caller applies the [tmin, tmax] acceptance.
// `inv` must be exactly (1/dir.x, 1/dir.y, 1/dir.z) (t_coef, hit.rint:87): callers that walk many leaves
// with one ray compute it once.
__device__ __forceinline__ bool leaf_dda_inv(f3 origin, f3 dir, f3 inv, uint64_t grid, float ray_tmin, float &t_out, uint32_t &voxel)
{
    // intersectAABB vs [0,4]^3, hit.rint:20-28 (true divisions)
    float ax = (0.0f - origin.x) / dir.x, bx = (4.0f - origin.x) / dir.x;
    float ay = (0.0f - origin.y) / dir.y, by = (4.0f - origin.y) / dir.y;
    float az = (0.0f - origin.z) / dir.z, bz = (4.0f - origin.z) / dir.z;
    float t_min = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    float t_max = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    if (t_min >= t_max) return false;                                   // :53
    if (t_max <= 0.0f) return false;                                    // :57
    float hitDistance = fmaxf(t_min, ray_tmin);                         // :71
    int px = (int)floorf(origin.x + dir.x * hitDistance);               // :72-74
    int py = (int)floorf(origin.y + dir.y * hitDistance);
    int pz = (int)floorf(origin.z + dir.z * hitDistance);
    px = min(max(px, 0), 3); py = min(max(py, 0), 3); pz = min(max(pz, 0), 3);
    float sx = signf(dir.x), sy = signf(dir.y), sz = signf(dir.z);      // :79
    float cx = inv.x, cy = inv.y, cz = inv.z;                           // :87
    float bx_ = cx * origin.x, by_ = cy * origin.y, bz_ = cz * origin.z;   // :88
    float mx = ((float)px + fmaxf(sx, 0.0f)) * cx - bx_;                // :90
    float my = ((float)py + fmaxf(sy, 0.0f)) * cy - by_;
    float mz = ((float)pz + fmaxf(sz, 0.0f)) * cz - bz_;
    float dx = (1.0f * cx) * sx, dy = (1.0f * cy) * sy, dz = (1.0f * cz) * sz;   // :96
    uint32_t hit = ((uint32_t)px << 4) | ((uint32_t)py << 2) | (uint32_t)pz;     // :99
    while (((grid >> hit) & 1ull) == 0) {                               // :101
        float kx = (mz >= mx ? 1.0f : 0.0f) * (my >= mx ? 1.0f : 0.0f);   // :102
        float ky = (mx >= my ? 1.0f : 0.0f) * (mz >= my ? 1.0f : 0.0f);
        float kz = (my >= mz ? 1.0f : 0.0f) * (mx >= mz ? 1.0f : 0.0f);
        px += (int)(sx * kx); py += (int)(sy * ky); pz += (int)(sz * kz);   // :103-104
        hitDistance = fminf(fminf(mx, my), mz);                         // :106
        if (hitDistance + 0.001f >= t_max) return false;                // :107
        mx += dx * kx; my += dy * ky; mz += dz * kz;                    // :110
        if ((px | py | pz) & ~3) return false;                          // deviation (2), see oracle
        hit = ((uint32_t)px << 4) | ((uint32_t)py << 2) | (uint32_t)pz;   // :112
    }
    t_out = hitDistance / 1.0f;                                         // :115
    voxel = hit;
    return true;
}

__device__ __forceinline__ bool leaf_dda(f3 origin, f3 dir, uint64_t grid, float ray_tmin, float &t_out, uint32_t &voxel)
{
    return leaf_dda_inv(origin, dir, mk3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z), grid, ray_tmin, t_out, voxel);
}

// ----------------------------------------------------------------------------- instance traversal

struct InstHit {
    float t;
    uint32_t leaf, voxel;
    int lx, ly, lz;          // leaf origin (object voxel coordinates)
    f3 o, d;                 // object-space ray
    uint32_t material_ptr;
    uint64_t mask;
};

struct Counters {
    uint32_t primary_rays, bounce_rays, shadow_rays, mask_words, child_ptrs, hits_shaded;
};

#define DUST_DIR_EPS 9.094947017729282e-13f   // 2^-40

__device__ __forceinline__ uint64_t u64_of(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// HDDA over the internal levels (DESIGN.md "Traversal semantics").  NL = 3: (4,2,2); NL = 4: (4,4,2,2).
template <int NL, bool STATS>
__device__ __forceinline__ bool trace_instance(const DInstance &in, f3 ow, f3 dw, float tmin, float tmax,
                                               InstHit &out, Counters &st)
{
    const float *M = in.w2o;
    f3 o, d;
    o.x = ((M[0] * ow.x + M[1] * ow.y) + M[2] * ow.z) + M[3];
    o.y = ((M[4] * ow.x + M[5] * ow.y) + M[6] * ow.z) + M[7];
    o.z = ((M[8] * ow.x + M[9] * ow.y) + M[10] * ow.z) + M[11];
    d.x = (M[0] * dw.x + M[1] * dw.y) + M[2] * dw.z;
    d.y = (M[4] * dw.x + M[5] * dw.y) + M[6] * dw.z;
    d.z = (M[8] * dw.x + M[9] * dw.y) + M[10] * dw.z;
    if (fabsf(d.x) < DUST_DIR_EPS) d.x = copysignf(DUST_DIR_EPS, d.x);
    if (fabsf(d.y) < DUST_DIR_EPS) d.y = copysignf(DUST_DIR_EPS, d.y);
    if (fabsf(d.z) < DUST_DIR_EPS) d.z = copysignf(DUST_DIR_EPS, d.z);
    const float ix_ = 1.0f / d.x, iy_ = 1.0f / d.y, iz_ = 1.0f / d.z;
    // clip against the occupied bounds [bmin, bmax) of the tree, not the whole [0, extent)^3
    const int bx0 = in.bmin[0], by0 = in.bmin[1], bz0 = in.bmin[2], bx1 = in.bmax[0], by1 = in.bmax[1], bz1 = in.bmax[2];
    // plane times are (plane - o) * inv: difference first, plane*inv - o*inv would cancel (|o/d| >> t)
    const float t0x = ((float)bx0 - o.x) * ix_, t0y = ((float)by0 - o.y) * iy_, t0z = ((float)bz0 - o.z) * iz_;
    const float t1x = ((float)bx1 - o.x) * ix_, t1y = ((float)by1 - o.y) * iy_, t1z = ((float)bz1 - o.z) * iz_;
    float t_enter = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), tmin));
    // walk a hair past tmax (acceptance stays exact): DESIGN.md "Traversal semantics"
    const float tmax_pad = tmax + (fabsf(tmax) * 3.814697265625e-06f + 9.5367431640625e-07f);
    float t_exit = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), tmax_pad));
    if (!(t_enter <= t_exit) || bx1 <= bx0) return false;
    int ix = (int)floorf(__fmaf_rn(d.x, t_enter, o.x));
    int iy = (int)floorf(__fmaf_rn(d.y, t_enter, o.y));
    int iz = (int)floorf(__fmaf_rn(d.z, t_enter, o.z));
    ix = min(max(ix, bx0), bx1 - 1); iy = min(max(iy, by0), by1 - 1); iz = min(max(iz, bz0), bz1 - 1);
    const bool px_ = d.x > 0.0f, py_ = d.y > 0.0f, pz_ = d.z > 0.0f;
    const f3 inv = mk3(ix_, iy_, iz_);

    // leave the cell of edge L holding (ix,iy,iz); returns false when the walk ends
    auto leave_cell = [&](int L) -> bool {
        const int cx = ix & ~(L - 1), cy = iy & ~(L - 1), cz = iz & ~(L - 1);
        const float tx = ((float)(px_ ? cx + L : cx) - o.x) * ix_;
        const float ty = ((float)(py_ ? cy + L : cy) - o.y) * iy_;
        const float tz = ((float)(pz_ ? cz + L : cz) - o.z) * iz_;
        const float t_next = fminf(fminf(tx, ty), tz);
        bool ok = (t_next <= t_exit);            // inclusive, DESIGN.md "Traversal semantics"
        int nx, ny, nz;
        if (tx <= t_next) { nx = px_ ? cx + L : cx - 1; ok &= (nx >= bx0 && nx < bx1); }
        else { const int ip = (int)floorf(__fmaf_rn(d.x, t_next, o.x)); nx = min(max(ip, cx), cx + L - 1); }
        if (ty <= t_next) { ny = py_ ? cy + L : cy - 1; ok &= (ny >= by0 && ny < by1); }
        else { const int ip = (int)floorf(__fmaf_rn(d.y, t_next, o.y)); ny = min(max(ip, cy), cy + L - 1); }
        if (tz <= t_next) { nz = pz_ ? cz + L : cz - 1; ok &= (nz >= bz0 && nz < bz1); }
        else { const int ip = (int)floorf(__fmaf_rn(d.z, t_next, o.z)); nz = min(max(ip, cz), cz + L - 1); }
        ix = nx; iy = ny; iz = nz;
        return ok;
    };

    // "while-while": every lane first walks empty cells until it stands on an occupied leaf (cheap
    // steps, lanes in lockstep), THEN the lanes that found one run the expensive leaf step together.
    uint32_t mid_key = 0, mid_lo = 0, mid_hi = 0, mid_base = 0;     // the 4^3 node of the current 16-cell
    for (;;) {
        uint32_t leaf_i = 0;
        bool walking = true;
        for (;;) {
            int L;
            bool at_leaf = false;
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
                bool have_mid = mid_key == key;
                L = 16;
                if (!have_mid) {
                    const uint64_t m = __ldg(&n16->mask[ridx >> 6]);
                    const uint32_t b = ridx & 63u;
                    if ((m >> b) & 1ull) {
                        const uint32_t mid_i = __ldg(&n16->child[ridx >> 6]) + __popcll(m & ((1ull << b) - 1ull));
                        const uint4 mid = __ldg(&in.mids[mid_i]);
                        mid_key = key; mid_lo = mid.x; mid_hi = mid.y; mid_base = mid.z;
                        have_mid = true;
                    }
                }
                if (have_mid) {
                    if (STATS) { st.child_ptrs++; st.mask_words++; }
                    const uint64_t mm = u64_of(mid_lo, mid_hi);
                    const uint32_t midx = ((uint32_t)((ix >> 2) & 3) << 4) | ((uint32_t)((iy >> 2) & 3) << 2) | (uint32_t)((iz >> 2) & 3);
                    L = 4;
                    if ((mm >> midx) & 1ull) {
                        if (STATS) { st.child_ptrs++; st.mask_words++; }
                        leaf_i = mid_base + __popcll(mm & ((1ull << midx) - 1ull));
                        at_leaf = true;
                    }
                }
            }
            if (at_leaf) break;
            if (!leave_cell(L)) { walking = false; break; }
        }
        if (!walking) return false;
        // leaf step: one 128-bit load (occupancy + material_ptr), then hit.rint's DDA
        const uint4 leaf = __ldg(&in.leaves[leaf_i]);
        const int cx = ix & ~3, cy = iy & ~3, cz = iz & ~3;
        const uint64_t grid = u64_of(leaf.x, leaf.y);
        const f3 lo = mk3(o.x - (float)cx, o.y - (float)cy, o.z - (float)cz);   // hit.rint:140
        float th; uint32_t vox;
        if (leaf_dda_inv(lo, d, inv, grid, tmin, th, vox) && tmin <= th && th <= tmax) {
            out.t = th; out.leaf = leaf_i; out.voxel = vox;
            out.lx = cx; out.ly = cy; out.lz = cz;
            out.o = o; out.d = d;
            out.material_ptr = leaf.z; out.mask = grid;
            return true;
        }
        if (!leave_cell(4)) return false;
    }
}

template <bool STATS>
__device__ __forceinline__ bool trace_instance_any_levels(const DInstance &in, f3 ow, f3 dw, float tmin, float tmax,
                                                          InstHit &out, Counters &st)
{
    if (in.n_levels == 3) return trace_instance<3, STATS>(in, ow, dw, tmin, tmax, out, st);
    return trace_instance<4, STATS>(in, ow, dw, tmin, tmax, out, st);
}

// ------------------------------------------------------------------------------------ scene level

struct SceneHit {
    bool hit;
    uint32_t instance;
    InstHit ih;
};

// conservative slab test of a (padded) world box; returns entry t or +inf
__device__ __forceinline__ float box_entry(const DBvhNode &n, f3 o, f3 inv, float tmin, float tmax)
{
    float ax = (n.lo[0] - o.x) * inv.x, bx = (n.hi[0] - o.x) * inv.x;
    float ay = (n.lo[1] - o.y) * inv.y, by = (n.hi[1] - o.y) * inv.y;
    float az = (n.lo[2] - o.z) * inv.z, bz = (n.hi[2] - o.z) * inv.z;
    float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), tmin));
    float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
    return t0 <= t1 ? t0 : CUDART_INF_F;
}

#define DUST_BVH_STACK 24
#define DUST_BVH_CAND 16

// Closest hit = lexicographic minimum of (t, instance) over all instances.  The BVH only culls
// (boxes are padded, see build_bvh), so the result equals the oracle's linear instance loop.
template <bool STATS, bool ANY>
__device__ __forceinline__ void trace_scene(const DSceneView &sc, f3 o, f3 d, float tmin, float tmax, SceneHit &best, Counters &st)
{
    best.hit = false; best.instance = 0xFFFFFFFFu;
    float cur = tmax;
    if (sc.n_inst == 0) return;
    // box tests only: a zero component gets a huge FINITE inverse so (lo - o) * inv is never 0 * inf
    f3 inv = mk3(fabsf(d.x) < 1e-30f ? copysignf(1e30f, d.x) : 1.0f / d.x,
                 fabsf(d.y) < 1e-30f ? copysignf(1e30f, d.y) : 1.0f / d.y,
                 fabsf(d.z) < 1e-30f ? copysignf(1e30f, d.z) : 1.0f / d.z);
    int stack[DUST_BVH_STACK];
    int sp = 0;
    int node = 0;
    {
        DBvhNode root = sc.bvh[0];
        if (box_entry(root, o, inv, tmin, cur) == CUDART_INF_F) return;
    }
    // Two phases instead of walking an instance the moment the BVH reaches it: (1) cull — collect the
    // instances whose padded box the ray touches, sorted by box entry; (2) walk them front to back.
    // All lanes of a warp are then inside the (long) tree walk at the same time instead of taking
    // turns, which is what kept the walk at ~3.5 active lanes (profiles/).  The result does not depend
    // on the visiting order (lexicographic min of (t, instance)), only the pruning does.
    float ct[DUST_BVH_CAND];
    int ci[DUST_BVH_CAND];
    bool more = true;
    while (more) {
        int nc = 0;
        for (;;) {                                                   // ---- phase 1: cull
            const DBvhNode n = sc.bvh[node];
            bool descend = false;
            if (n.left < 0) {
                const float te = box_entry(n, o, inv, tmin, cur);
                if (te != CUDART_INF_F) {
                    int k = nc++;                                    // insertion sort by (entry t, instance)
                    const int id = ~n.left;
                    while (k > 0 && (ct[k - 1] > te || (ct[k - 1] == te && ci[k - 1] > id))) { ct[k] = ct[k - 1]; ci[k] = ci[k - 1]; --k; }
                    ct[k] = te; ci[k] = id;
                }
            } else {
                float tl = box_entry(sc.bvh[n.left], o, inv, tmin, cur);
                float tr = box_entry(sc.bvh[n.right], o, inv, tmin, cur);
                int near_ = n.left, far_ = n.right;
                if (tr < tl) { near_ = n.right; far_ = n.left; const float s = tl; tl = tr; tr = s; }
                if (tl != CUDART_INF_F) {
                    if (tr != CUDART_INF_F && sp < DUST_BVH_STACK) stack[sp++] = far_;
                    node = near_;
                    descend = true;
                }
            }
            if (descend) continue;
            if (sp == 0) { more = false; break; }
            node = stack[--sp];
            if (nc == DUST_BVH_CAND) break;                          // list full: walk these, then resume the cull
        }
        for (int j = 0; j < nc; ++j) {                               // ---- phase 2: walk, front to back
            if (ct[j] > cur) break;                                  // sorted: nothing further can beat the best hit
            const uint32_t k = (uint32_t)ci[j];
            InstHit ih;
            if (trace_instance_any_levels<STATS>(sc.inst[k], o, d, tmin, cur, ih, st)) {
                if (ANY) { best.hit = true; best.instance = k; return; }
                if (!best.hit || ih.t < best.ih.t || (ih.t == best.ih.t && k < best.instance)) {
                    best.hit = true; best.instance = k; best.ih = ih; cur = ih.t;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ frame

struct FrameParams {
    DSceneView scene;
    DustCamera cam;
    DustSky sky;
    float sun_dir[3];          // normalize(sky.direction.xyz)
    float sun_cone;            // 1 - cos(solar radius)          (nee.rmiss:18-19, uniform-only)
    float sun_ar2;             // 1 / sin^2(solar radius)        (sky.glsl:88-89, uniform-only)
    DustTraceConfig cfg;
    uint32_t tiles_x, tiles_y, n_local_tiles;
    float4 *radiance32;
    __half *radiance16;
    uint32_t *hit_instance, *hit_leaf;
    uint8_t *hit_voxel;
    float *hit_t;
    DCounters *counters;
};

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

struct Surface {
    f3 x, n, albedo_lin;
};

// primary/hit.rchit:16-76 + ambient_occlusion.rgen:23
template <bool STATS>
__device__ __forceinline__ void surface_of(const DSceneView &sc, const SceneHit &h, f3 o, f3 d, Surface &sf, Counters &st)
{
    const DInstance &in = sc.inst[h.instance];
    const InstHit &ih = h.ih;
    float vx = (ih.t * ih.d.x + ih.o.x) - (((float)ih.lx + (float)((ih.voxel >> 4) & 3u)) + 0.5f);
    float vy = (ih.t * ih.d.y + ih.o.y) - (((float)ih.ly + (float)((ih.voxel >> 2) & 3u)) + 0.5f);
    float vz = (ih.t * ih.d.z + ih.o.z) - (((float)ih.lz + (float)(ih.voxel & 3u)) + 0.5f);
    float ax = fabsf(vx), ay = fabsf(vy), az = fabsf(vz);
    float mx = fmaxf(ax, fmaxf(ay, az));
    f3 n_obj = mk3(signf(vx) * (ax >= mx ? 1.0f : 0.0f), signf(vy) * (ay >= mx ? 1.0f : 0.0f), signf(vz) * (az >= mx ? 1.0f : 0.0f));
    const float *W = in.o2w;
    sf.n = mk3((W[0] * n_obj.x + W[1] * n_obj.y) + W[2] * n_obj.z,
               (W[3] * n_obj.x + W[4] * n_obj.y) + W[5] * n_obj.z,
               (W[6] * n_obj.x + W[7] * n_obj.y) + W[8] * n_obj.z);
    uint32_t off = __popcll(ih.mask & ((1ull << ih.voxel) - 1ull));              // hit.rchit:60-61
    uint8_t pi = __ldg(&in.materials[ih.material_ptr + off]);                   // :71
    float4 c = __ldg(&in.palette_lin[pi]);                                      // :72-74 + SRGBToLinear
    sf.albedo_lin = mk3(c.x, c.y, c.z);
    sf.x = mk3((ih.t * d.x + o.x) + sf.n.x * 0.01f, (ih.t * d.y + o.y) + sf.n.y * 0.01f, (ih.t * d.z + o.z) + sf.n.z * 0.01f);
    if (STATS) st.hits_shaded++;
}

// One pixel: spp samples of primary + NEE + up to max_bounces diffuse bounces.
template <bool STATS>
__device__ __forceinline__ void shade_pixel(const FrameParams &p, uint32_t px, uint32_t py, uint32_t out_index, Counters &st)
{
    const DustTraceConfig &cfg = p.cfg;
    const uint32_t pixel_index = py * cfg.width + px;
    const uint32_t key = pcg(cfg.seed + 0x9E3779B9u * cfg.frame_index);
    const uint32_t s0 = pcg(pixel_index ^ key);
    const f3 sun_dir = mk3(p.sun_dir[0], p.sun_dir[1], p.sun_dir[2]);
    const f3 sky_dir = mk3(p.sky.direction[0], p.sky.direction[1], p.sky.direction[2]);
    f3 acc = mk3(0.0f, 0.0f, 0.0f);
    for (uint32_t smp = 0; smp < cfg.spp; smp++) {
        const uint32_t s1 = pcg(s0 + smp);
        float jx = 0.5f, jy = 0.5f;
        if (smp > 0) { jx = u01(pcg(s1 + 0u)); jy = u01(pcg(s1 + 1u)); }
        f3 o, d;
        camera_ray(p.cam, cfg.width, cfg.height, px, py, jx, jy, o, d);
        SceneHit h;
        trace_scene<STATS, false>(p.scene, o, d, p.cam.near, p.cam.far, h, st);
        if (STATS) st.primary_rays++;
        f3 L = mk3(0.0f, 0.0f, 0.0f);
        const bool write_hits = smp == 0 && (cfg.flags & DUST_TRACE_WRITE_HITS) && p.hit_instance;
        if (!h.hit) {
            f3 dn = normalize3(d);                                               // miss.rmiss:8-14
            f3 sk = sky_radiance(p.sky, dn);
            f3 su = sun_radiance(p.sky, p.sun_ar2, dn);
            L = acescg2srgb(mk3((sk.x + su.x) / 3.14f, (sk.y + su.y) / 3.14f, (sk.z + su.z) / 3.14f));
            if (write_hits) {
                p.hit_instance[pixel_index] = 0xFFFFFFFFu; p.hit_leaf[pixel_index] = 0xFFFFFFFFu;
                p.hit_voxel[pixel_index] = 0xFF; p.hit_t[pixel_index] = CUDART_INF_F;
            }
        } else {
            Surface sf;
            surface_of<STATS>(p.scene, h, o, d, sf, st);
            if (write_hits) {
                p.hit_instance[pixel_index] = h.instance; p.hit_leaf[pixel_index] = h.ih.leaf;
                p.hit_voxel[pixel_index] = (uint8_t)h.ih.voxel; p.hit_t[pixel_index] = h.ih.t;
            }
            f3 T = sf.albedo_lin;                                                // tone_map.comp:195-203
            for (uint32_t v = 0;; v++) {
                if (dot3(sky_dir, sf.n) > 0.0f) {                                // ambient_occlusion.rgen:33-50
                    if (STATS) st.shadow_rays++;
                    SceneHit sh;
                    trace_scene<STATS, true>(p.scene, sf.x, sun_dir, 0.1f, 10000.0f, sh, st);
                    if (!sh.hit) {                                               // nee.rmiss:11-22
                        f3 su = sun_radiance(p.sky, p.sun_ar2, normalize3(sun_dir));
                        float ndl = dot3(sf.n, sun_dir);
                        f3 ts = acescg2srgb(mk3((su.x * p.sun_cone) * ndl, (su.y * p.sun_cone) * ndl, (su.z * p.sun_cone) * ndl));
                        L.x += T.x * ts.x; L.y += T.y * ts.y; L.z += T.z * ts.z;
                    }
                }
                if (v == cfg.max_bounces) break;
                float u1 = u01(pcg(s1 + 2u + 2u * v)), u2 = u01(pcg(s1 + 3u + 2u * v));   // final_gather.rgen:32-34
                float r = sqrtf(u1), cz = sqrtf(1.0f - u1), c, sn;
                cossin_2pi(u2, c, sn);
                f3 w = normalize3(rotate_by_normal(sf.n, mk3(r * c, r * sn, cz)));
                SceneHit bh;
                f3 xo = sf.x;
                trace_scene<STATS, false>(p.scene, xo, w, 0.1f, p.cam.far, bh, st);
                if (STATS) st.bounce_rays++;
                if (!bh.hit) {                                                   // final_gather.rmiss:13-17
                    f3 ts = acescg2srgb(sky_radiance(p.sky, normalize3(w)));
                    L.x += T.x * ts.x; L.y += T.y * ts.y; L.z += T.z * ts.z;
                    break;
                }
                surface_of<STATS>(p.scene, bh, xo, w, sf, st);                   // final_gather.rchit:76-80
                T.x = T.x * sf.albedo_lin.x; T.y = T.y * sf.albedo_lin.y; T.z = T.z * sf.albedo_lin.z;
            }
        }
        f3 Lc = srgb2acescg(L);
        acc.x += Lc.x; acc.y += Lc.y; acc.z += Lc.z;
    }
    float inv_n = (float)cfg.spp;
    float4 outv = make_float4(acc.x / inv_n, acc.y / inv_n, acc.z / inv_n, 1.0f);
    if (p.radiance32) p.radiance32[out_index] = outv;
    if (p.radiance16) {
        __half2 a = __floats2half2_rn(outv.x, outv.y), b = __floats2half2_rn(outv.z, outv.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t *>(&a); pk.y = *reinterpret_cast<uint32_t *>(&b);
        reinterpret_cast<uint2 *>(p.radiance16)[out_index] = pk;
    }
}

__device__ __forceinline__ void flush_counters(DCounters *g, const Counters &st)
{
    // warp-aggregate, one atomic per counter per warp
    uint32_t v[6] = { st.primary_rays, st.bounce_rays, st.shadow_rays, st.mask_words, st.child_ptrs, st.hits_shaded };
    unsigned long long *dst = reinterpret_cast<unsigned long long *>(g);
#pragma unroll
    for (int i = 0; i < 6; i++) {
        unsigned long long s = v[i];
        for (int ofs = 16; ofs > 0; ofs >>= 1) s += __shfl_xor_sync(0xffffffffu, s, ofs);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(&dst[i], s);
    }
}

// v1 mapping: a warp owns an 8x4 pixel patch; a 128-thread block owns a 16x8 patch; 32x32 tiles
// are the sharding unit (tile t -> shard t % shard_count).
template <bool STATS>
__global__ void __launch_bounds__(128) frame_kernel(const __grid_constant__ FrameParams p)
{
    const uint32_t blocks_per_tile = 8;                 // (32*32) / 128
    const uint32_t lt = blockIdx.x / blocks_per_tile;   // local tile
    const uint32_t sub = blockIdx.x % blocks_per_tile;  // 16x8 patch inside the tile: 2 across, 4 down
    const uint32_t t = lt * p.cfg.shard_count + p.cfg.shard_index;
    const uint32_t tile_x = t % p.tiles_x, tile_y = t / p.tiles_x;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lx = (sub & 1) * 16 + (warp & 1) * 8 + (lane & 7);
    const uint32_t ly = (sub >> 1) * 8 + (warp >> 1) * 4 + (lane >> 3);
    const uint32_t px = tile_x * 32 + lx, py = tile_y * 32 + ly;
    Counters st = { 0, 0, 0, 0, 0, 0 };
    if (tile_y < p.tiles_y && px < p.cfg.width && py < p.cfg.height) {
        const uint32_t out_index = p.cfg.shard_count > 1 ? lt * 1024u + ly * 32u + lx : py * p.cfg.width + px;
        shade_pixel<STATS>(p, px, py, out_index, st);
    }
    if (STATS) flush_counters(p.counters, st);
}

// ------------------------------------------------------------------------------------- ray batches

struct RayParams {
    DSceneView scene;
    const DustRay *rays;
    DustHit *hits;
    unsigned long long n;
    DCounters *counters;
};

template <bool STATS>
__global__ void __launch_bounds__(128) rays_kernel(const __grid_constant__ RayParams p)
{
    Counters st = { 0, 0, 0, 0, 0, 0 };
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        // 32-byte ray = two coalesced 128-bit loads
        const float4 a = __ldg(reinterpret_cast<const float4 *>(p.rays) + 2 * i);
        const float4 b = __ldg(reinterpret_cast<const float4 *>(p.rays) + 2 * i + 1);
        SceneHit h;
        trace_scene<STATS, false>(p.scene, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), a.w, b.w, h, st);
        if (STATS) st.primary_rays++;
        uint4 o;
        if (h.hit) { o.x = __float_as_uint(h.ih.t); o.y = h.instance; o.z = h.ih.leaf; o.w = h.ih.voxel; }
        else { o.x = __float_as_uint(CUDART_INF_F); o.y = o.z = o.w = 0xFFFFFFFFu; }
        reinterpret_cast<uint4 *>(p.hits)[i] = o;
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
