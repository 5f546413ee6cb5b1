// sm_100a device code, part 1: scalar helpers, PRNG / sampling, colour, sky model and the LEAF STEP.
// Built with -fmad=false: every float expression is evaluated exactly as written (explicit
// __fmaf_rn only where DESIGN.md's traversal spec says "fma"), so hit ids and t are reproducible
// bit-for-bit against the CPU checker.  Reference sources restated here:
//   PRNG         assets/shaders/headers/spatial_hash.glsl:105-111
//   normals      assets/shaders/headers/normal.glsl:31-43
//   colour       assets/shaders/headers/color.glsl:1-31
//   sky / sun    assets/shaders/headers/sky.glsl:1-113
//   leaf step    assets/shaders/primary/hit.rint:20-28, 43-131
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
__device__ __forceinline__ uint64_t u64_of(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// bit `i` (0..63) of the 64-bit mask (lo, hi): one 64-bit funnel shift (SHF.R.U64) + one LOP3
__device__ __forceinline__ bool bit64(uint32_t lo, uint32_t hi, uint32_t i) { return (uint32_t)(u64_of(lo, hi) >> i) & 1u; }
// number of set bits below bit `i`, which MUST itself be set (every caller has just tested it): shifting
// bit i up to bit 63 leaves exactly the bits <= i, so the count is popc - 1 (2 SHF + 2 POPC + 1 IADD3,
// where the select-by-word form took 14 instructions on the saturated ALU pipe, profiles/r01_history.md)
__device__ __forceinline__ uint32_t rank64(uint32_t lo, uint32_t hi, uint32_t i)
{
    return (uint32_t)__popcll(u64_of(lo, hi) << (63u - i)) - 1u;
}

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
    const float cc = (q & 1) ? s45 : c45, ss = (q & 1) ? c45 : s45;     // quadrant rotation as selects: the four-way branch split every warp
    c = (q == 1 || q == 2) ? -cc : cc;
    s = (q >= 2) ? -ss : ss;
}

// x / l for a finite l > 0.  A zero numerator — every axis-aligned normal has two — is answered without the divider: FCHK sends
// x == 0 to the out-of-line slow path of the IEEE division (profiles/r02_history.md 9), and +-0 / l is +-0 bit for bit.  The
// substitute numerator goes through inline PTX: written in C, the compiler folds it away again and divides x itself.
__device__ __forceinline__ float div_pos(float x, float l)
{
    float num;
    asm("{ .reg .pred z; setp.eq.f32 z, %1, 0f00000000; selp.f32 %0, 0f3F800000, %1, z; }" : "=f"(num) : "f"(x));
    const float r = num / l;
    return x == 0.0f ? x : r;
}

// headers/normal.glsl:31-37
__device__ __forceinline__ f3 rotate_by_normal(f3 n, f3 v)
{
    float q0 = -n.y, q1 = n.x, q2 = 0.0f, q3 = 1.0f + n.z;
    float l = sqrtf(((q0 * q0 + q1 * q1) + q2 * q2) + q3 * q3);
    q0 = div_pos(q0, l); q1 = div_pos(q1, l); q2 = div_pos(q2, l); q3 = div_pos(q3, l);   // l == 0 only for n = -z, replaced below
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

// exp / acos as FIXED mul / add / sqrt expressions (deviation (6)) — the same operations in the same order as
// the CPU checker's, so that sky radiance (and everything accumulated or cached from it) is
// bit-identical on both sides; CUDA's expf / acosf and glibc's round differently.
__device__ __forceinline__ float exp2_fixed(float x)
{
    const float nf = floorf(x + 0.5f);
    const float f = (x - nf) * 0.693147181f;
    const float p = 1.0f + f * (1.0f + f * (0.5f + f * (0.166666667f + f * (0.0416666667f + f * (0.00833333333f + f * (0.00138888889f + f * 0.000198412698f))))));
    if (!(nf >= -126.0f)) return 0.0f;
    if (nf > 127.0f) return CUDART_INF_F;
    return p * __uint_as_float((uint32_t)((int32_t)nf + 127) << 23);
}
__device__ __forceinline__ float exp_fixed(float x) { return exp2_fixed(x * 1.44269504f); }
__device__ __forceinline__ float acos_fixed(float x)
{
    const float a = fabsf(x);
    const float p = 1.5707963050f + a * (-0.2145988016f + a * (0.0889789874f + a * (-0.0501743046f + a * (0.0308918810f + a * (-0.0170881256f + a * (0.0066700901f + a * -0.0012624911f))))));
    const float r = sqrtf(1.0f - a) * p;
    return x < 0.0f ? 3.14159265358979f - r : r;
}

// headers/sky.glsl:1-16
__device__ __forceinline__ float sky_channel(const DustSkyChannel &ch, float cos_theta, float gamma, float cos_gamma)
{
    const float *c = ch.config;
    float expM = exp_fixed(c[4] * gamma);
    float rayM = cos_gamma * cos_gamma;
    float mb = (1.0f + c[8] * c[8]) - (2.0f * c[8]) * cos_gamma;
    float mieM = (1.0f + rayM) / (mb * sqrtf(mb));                       // pow(x, 1.5) = x sqrt(x)
    float zenith = sqrtf(cos_theta);
    return (1.0f + c[0] * exp_fixed(c[1] / (cos_theta + 0.01f))) *
           ((((c[2] + c[3] * expM) + c[5] * rayM) + c[6] * mieM) + c[7] * zenith);
}

// Sky and sun radiance of direction `dir` (unit), ACEScg.  headers/sky.glsl:19-78 (cos_gamma clamped
// before acos, like the oracle) and :80-113 (ar2 = 1/sin^2(solar radius) is a uniform-only term hoisted
// to the host).  ONE out-of-line copy with a rolled channel loop: expf/powf/acosf are a few hundred
// instructions each and this is the coldest code of the frame kernel — inlined at every miss site it
// tripled the kernel and pushed the traversal loops out of the instruction cache (profiles/).
__device__ __noinline__ void sky_eval(const DustSky &sky, float ar2, float dx, float dy, float dz, bool with_sun,
                                      float &sk_x, float &sk_y, float &sk_z, float &su_x, float &su_y, float &su_z)
{
    f3 skyv = mk3(0.0f, 0.0f, 0.0f), sunv = mk3(0.0f, 0.0f, 0.0f);
    const float cos_gamma = (dx * sky.direction[0] + dy * sky.direction[1]) + dz * sky.direction[2];
    if (sky.direction[1] > 0.0f) {
        const float cos_theta = fminf(fmaxf(dy, 0.0f), 1.0f);
        const float gamma = acos_fixed(fminf(fmaxf(cos_gamma, -1.0f), 1.0f));
        float xyz[3];
        const DustSkyChannel *ch = &sky.r;
#pragma unroll 1
        for (int i = 0; i < 3; i++) xyz[i] = (sky_channel(ch[i], cos_theta, gamma, cos_gamma) * ch[i].radiance) * 683.0f;
        skyv = xyz2acescg(mk3(xyz[0], xyz[1], xyz[2]));
    }
    if (with_sun && !(cos_gamma < 0.0f || dy < 0.0f)) {
        const float singamma = 1.0f - (cos_gamma * cos_gamma);
        const float sc2 = 1.0f - ar2 * singamma * singamma;
        if (sc2 > 0.0f) {
            const float sc = sqrtf(sc2);
            float dark[3];
            const DustSkyChannel *ch = &sky.r;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                float dk = ch[i].ld_coefficient0;
                dk += ch[i].ld_coefficient1 * sc;
                float cur = sc;
#pragma unroll
                for (int k = 0; k < 4; k++) { cur *= sc; dk += ch[i].ld_coefficient2[k] * cur; }
                dark[i] = dk;
            }
            sunv = xyz2acescg(mk3(sky.sunlight_intensity[0] * dark[0], sky.sunlight_intensity[1] * dark[1], sky.sunlight_intensity[2] * dark[2]));
        }
    }
    sk_x = skyv.x; sk_y = skyv.y; sk_z = skyv.z;
    su_x = sunv.x; su_y = sunv.y; su_z = sunv.z;
}

// --------------------------------------------------------------------------------------- LEAF STEP

// assets/shaders/primary/hit.rint:43-131 (scale = 1).  origin is leaf-local.  Returns true with
// (t, voxelId) when the DDA reaches a set bit; the caller applies the [tmin, tmax] acceptance.
// `inv` must be exactly (1/dir.x, 1/dir.y, 1/dir.z) (t_coef, hit.rint:87): callers that walk many
// leaves with one ray compute it once.  dir has no zero component (deviation (1)), so step = sign(dir)
// is +-1 and is carried as a predicate; the 0/1 `comp` factors of hit.rint:102-110 are carried as
// predicates too: x + d * 1.0f == x + d and x + d * 0.0f == x + 0.0f for finite d, bit for bit.
// AO: assets/shaders/final_gather/ambient_occlusion.rint:46-134, the intersection shader of the ambient-occlusion pass' two rays — the
// same DDA plus one rule (:62-73): when AMBIENT_OCCLUSION_THRESHOLD (8.0) lies inside this leaf's box along the ray, a non-empty leaf
// is reported as a solid block at its entry (voxelId 0xFF) and an empty one is skipped, without a DDA.
template <bool COUNT, bool AO = false>
__device__ __forceinline__ bool leaf_dda_inv(f3 origin, f3 dir, f3 inv, uint32_t grid_lo, uint32_t grid_hi, float ray_tmin,
                                             float &t_out, uint32_t &voxel, uint32_t &iterations)
{
    // intersectAABB vs [0,4]^3, hit.rint:20-28 (true divisions)
    float ax = (0.0f - origin.x) / dir.x, bx = (4.0f - origin.x) / dir.x;
    float ay = (0.0f - origin.y) / dir.y, by = (4.0f - origin.y) / dir.y;
    float az = (0.0f - origin.z) / dir.z, bz = (4.0f - origin.z) / dir.z;
    float t_min = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    float t_max = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    if (t_min >= t_max) return false;                                   // :53
    if (t_max <= 0.0f) return false;                                    // :57
    if (AO && t_min <= 8.0f && 8.0f <= t_max) {
        if ((grid_lo | grid_hi) == 0u) return false;
        t_out = t_min + 0.0f;
        voxel = 0xFFu;
        return true;
    }
    float hitDistance = fmaxf(t_min, ray_tmin);                         // :71
    int px = __float2int_rd(origin.x + dir.x * hitDistance);            // :72-74
    int py = __float2int_rd(origin.y + dir.y * hitDistance);
    int pz = __float2int_rd(origin.z + dir.z * hitDistance);
    px = min(max(px, 0), 3); py = min(max(py, 0), 3); pz = min(max(pz, 0), 3);
    const bool posx = dir.x > 0.0f, posy = dir.y > 0.0f, posz = dir.z > 0.0f;      // :79 (step = +-1)
    float cx = inv.x, cy = inv.y, cz = inv.z;                           // :87
    float bx_ = cx * origin.x, by_ = cy * origin.y, bz_ = cz * origin.z;   // :88
    float mx = ((float)px + (posx ? 1.0f : 0.0f)) * cx - bx_;           // :90
    float my = ((float)py + (posy ? 1.0f : 0.0f)) * cy - by_;
    float mz = ((float)pz + (posz ? 1.0f : 0.0f)) * cz - bz_;
    const float dx = (1.0f * cx) * (posx ? 1.0f : -1.0f), dy = (1.0f * cy) * (posy ? 1.0f : -1.0f), dz = (1.0f * cz) * (posz ? 1.0f : -1.0f);   // :96
    // The cell is carried as x << 6 | y << 3 | z: 3-bit fields whose top bit is a GUARD.  A step that leaves
    // 0..3 on an axis sets a guard bit (+1 from 3 -> 4; -1 from 0 -> 7, borrowing from the field above,
    // which can only turn that field's 4 into 3 or its 7 into 6: whenever any axis leaves, at least the
    // lowest leaving field keeps its guard).  So "did any stepped axis leave" is ONE mask test, with no
    // per-axis branches (they ran with ~3 of 32 lanes, profiles/r01_v4_wf_trace_pt_lines.txt).
    const int isx = posx ? 64 : -64, isy = posy ? 8 : -8, isz = posz ? 1 : -1;
    int cell = (px << 6) | (py << 3) | pz;
    int hit = (px << 4) | (py << 2) | pz;                               // :99
#pragma unroll 1
    while (!bit64(grid_lo, grid_hi, (uint32_t)hit)) {                   // :101
        if (COUNT) iterations++;
        const bool kx = (mz >= mx) && (my >= mx);                       // :102
        const bool ky = (mx >= my) && (mz >= my);
        const bool kz = (my >= mz) && (mx >= mz);
        hitDistance = fminf(fminf(mx, my), mz);                         // :106
        cell += (kx ? isx : 0) + (ky ? isy : 0) + (kz ? isz : 0);       // :103-104, :112
        // :107, and deviation (2), see oracle: a tie-step that leaves 0..3 on any axis ends the walk — one exit for both
        if ((hitDistance + 0.001f >= t_max) | ((cell & 0x124) != 0)) return false;
        hit = ((cell >> 2) & 0x30) | ((cell >> 1) & 0xC) | (cell & 3);
        mx += kx ? dx : 0.0f; my += ky ? dy : 0.0f; mz += kz ? dz : 0.0f;          // :110
    }
    t_out = hitDistance / 1.0f + 0.0f;                                  // :115; + 0.0f: canonical +0 (see the CPU checker)
    voxel = (uint32_t)hit;
    return true;
}

// assets/shaders/final_gather/rough.rint:46-58 — the far-field intersection shader: a non-empty leaf is one solid 4x4x4 box,
// hit where the ray enters it (no DDA); voxel id 0.  The caller applies the [tmin, tmax] acceptance.
__device__ __forceinline__ bool leaf_rough(f3 origin, f3 dir, uint32_t grid_lo, uint32_t grid_hi, float &t_out, uint32_t &voxel)
{
    float ax = (0.0f - origin.x) / dir.x, bx = (4.0f - origin.x) / dir.x;
    float ay = (0.0f - origin.y) / dir.y, by = (4.0f - origin.y) / dir.y;
    float az = (0.0f - origin.z) / dir.z, bz = (4.0f - origin.z) / dir.z;
    const float t_min = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    const float t_max = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    if (t_min >= t_max || (grid_lo | grid_hi) == 0u) return false;      // :47-53
    t_out = t_min + 0.0f;                                               // :54; canonical +0
    voxel = 0u;
    return true;
}

}  // namespace dust
