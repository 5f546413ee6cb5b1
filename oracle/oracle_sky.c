/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_sky.h).
 * CPU restatement of crates/render/src/pipeline/sky.rs and assets/shaders/headers/{sky,color}.glsl.
 */
#include "oracle_sky.h"
#include "hosek_wilkie_xyz_tables.h"
#include <math.h>
#include <string.h>

#define O_PI_F 3.14159265358979323846f
#define O_FRAC_PI_2_F 1.57079632679489661923f

static float powi_f(float x, int n) { float r = 1.0f; for (int i = 0; i < n; i++) r *= x; return r; }

/* sky.rs:135-143 — quintic Bezier in the (cube-rooted, normalised) solar elevation */
static void coefficient(const float *elev_matrix /* [6][3] */, float se, float out[3])
{
    float rev = 1.0f - se;
    float w[6] = {
        powi_f(rev, 5),
        5.0f * powi_f(rev, 4) * se,
        10.0f * powi_f(rev, 3) * powi_f(se, 2),
        10.0f * powi_f(rev, 2) * powi_f(se, 3),
        5.0f * rev * powi_f(se, 4),
        powi_f(se, 5),
    };
    for (int c = 0; c < 3; c++) {
        float acc = w[0] * elev_matrix[c];
        for (int k = 1; k < 6; k++) acc = acc + w[k] * elev_matrix[3 * k + c];
        out[c] = acc;
    }
}

/* shared blend of sky.rs:145-179 and :181-227: low/high albedo x low/high turbidity */
static void blend(const float *low, const float *high, uint32_t stride, float turbidity,
                  const float albedo[3], float se, float out[3])
{
    uint32_t it = (uint32_t)turbidity;
    float rem = turbidity - (float)it;
    float c[3];
    coefficient(low + (it - 1) * stride, se, c);
    for (int k = 0; k < 3; k++) out[k] = (1.0f - albedo[k]) * (1.0f - rem) * c[k];
    coefficient(high + (it - 1) * stride, se, c);
    for (int k = 0; k < 3; k++) out[k] += albedo[k] * (1.0f - rem) * c[k];
    if (it < 10) {
        coefficient(low + it * stride, se, c);
        for (int k = 0; k < 3; k++) out[k] += (1.0f - albedo[k]) * rem * c[k];
        coefficient(high + it * stride, se, c);
        for (int k = 0; k < 3; k++) out[k] += albedo[k] * rem * c[k];
    }
}

/* sky.rs:229-253 */
static void sr_internal_xyz(uint32_t turbidity, float elevation, float out[3])
{
    const uint32_t PIECES = 45, ORDER = 4;
    uint32_t pos = (uint32_t)(powf(2.0f * elevation / O_PI_F, 1.0f / 3.0f) * (float)PIECES);
    if (pos > PIECES - 1) pos = PIECES - 1;
    float break_x = powi_f((float)pos / (float)PIECES, 3) * O_FRAC_PI_2_F;
    float x = elevation - break_x, x_exp = 1.0f;
    uint32_t base = ORDER * PIECES * turbidity + ORDER * pos;
    out[0] = out[1] = out[2] = 0.0f;
    for (int k = (int)ORDER - 1; k >= 0; k--) {
        const float *coef = &HW_SOLAR_DATASET[3 * (base + (uint32_t)k)];
        for (int c = 0; c < 3; c++) out[c] += coef[c] * x_exp;
        x_exp *= x;
    }
}

/* sky.rs:254-268 */
static void solar_direct_radiance_xyz(float elevation, float turbidity, float out[3])
{
    uint32_t turb_low = (uint32_t)turbidity - 1;
    float turb_frac = turbidity - (float)(turb_low + 1);
    if (turb_low == 9) { turb_low = 8; turb_frac = 1.0f; }
    float a[3], b[3];
    sr_internal_xyz(turb_low, elevation, a);
    sr_internal_xyz(turb_low + 1, elevation, b);
    for (int c = 0; c < 3; c++) out[c] = (1.0f - turb_frac) * a[c] + turb_frac * b[c];
}

/* sky.rs:90-133 */
void o_sunlight_bake(float turbidity, const float albedo[3], const float direction[3], OSky *out)
{
    memset(out, 0, sizeof *out);
    float theta = asinf(direction[1]);
    float se = powf(theta / O_FRAC_PI_2_F, 1.0f / 3.0f);
    OSkyChannel *ch[3] = { &out->r, &out->g, &out->b };
    for (int i = 0; i < 9; i++) {
        float res[3];
        /* CONFIG_*[turbidity][i][6] -> stride between turbidities = 9*6 Vec3 */
        blend(HW_CONFIG_LOW_ALBEDO + (size_t)i * 18, HW_CONFIG_HIGH_ALBEDO + (size_t)i * 18, 9 * 18,
              turbidity, albedo, se, res);
        for (int c = 0; c < 3; c++) ch[c]->config[i] = res[c];
    }
    float rad[3];
    blend(HW_RAD_LOW_ALBEDO, HW_RAD_HIGH_ALBEDO, 18, turbidity, albedo, se, rad);
    float solar[3];
    solar_direct_radiance_xyz(theta, turbidity, solar);
    for (int c = 0; c < 3; c++) {
        ch[c]->radiance = rad[c];
        ch[c]->ld_coefficient0 = HW_SOLAR_LD_COEFS[0 * 3 + c];
        ch[c]->ld_coefficient1 = HW_SOLAR_LD_COEFS[1 * 3 + c];
        for (int k = 0; k < 4; k++) ch[c]->ld_coefficient2[k] = HW_SOLAR_LD_COEFS[(2 + k) * 3 + c];
    }
    out->direction[0] = direction[0]; out->direction[1] = direction[1]; out->direction[2] = direction[2];
    out->direction[3] = 0.0f;
    out->sunlight_intensity[0] = solar[0]; out->sunlight_intensity[1] = solar[1]; out->sunlight_intensity[2] = solar[2];
    out->sunlight_intensity[3] = (0.51f * (O_PI_F / 180.0f)) / 2.0f;
}

/* ------------------------------------------------------------------ colour */

/* color.glsl:1-5 */
float o_srgb_to_linear(float c)
{
    return c < 0.04045f ? c / 12.92f : powf(fabsf(c + 0.055f) / 1.055f, 2.4f);
}

/* GLSL mat3(c0, c1, c2) * v = c0*v.x + c1*v.y + c2*v.z */
static void mat3_mul(const float m[9], const float v[3], float out[3])
{
    float r[3];
    for (int i = 0; i < 3; i++) r[i] = (m[i] * v[0] + m[3 + i] * v[1]) + m[6 + i] * v[2];
    out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}

/* color.glsl:8-15 */
void o_srgb2acescg(const float in[3], float out[3])
{
    static const float m[9] = { 0.6031065f, 0.07011794f, 0.022178888f, 0.32633433f, 0.9199162f, 0.11607823f,
                                0.047995567f, 0.012763573f, 0.94101846f };
    mat3_mul(m, in, out);
}
/* color.glsl:16-23 */
void o_acescg2srgb(const float in[3], float out[3])
{
    static const float m[9] = { 1.7312546f, -0.131619f, -0.024568284f, -0.6040432f, 1.1348418f, -0.12575036f,
                                -0.08010775f, -0.008679431f, 1.0656371f };
    mat3_mul(m, in, out);
}
/* color.glsl:24-31 */
void o_xyz2acescg(const float in[3], float out[3])
{
    static const float m[9] = { 1.6410228f, -0.66366285f, 0.011721907f, -0.32480323f, 1.6153315f, -0.0082844375f,
                                -0.23642465f, 0.016756356f, 0.9883947f };
    mat3_mul(m, in, out);
}

/* --------------------------------------------------------------------- sky */

/* exp / acos as FIXED mul / add / sqrt expressions (deviation (6)): GLSL leaves their precision to the implementation,
 * glibc and CUDA round differently, and the radiance cache (oracle_gi.c) quantises what the sky returns — so that cache tables
 * and frames can be compared bit for bit, both sides evaluate these two functions with the same operations in the same order
 * (relative error < 1e-7, below the float32 rounding of the surrounding expression).  o_sky_exp2 is oracle_cache.c's exp2_fixed. */
float o_sky_exp2(float x)
{
    float nf = floorf(x + 0.5f);
    float f = (x - nf) * 0.693147181f;
    float p = 1.0f + f * (1.0f + f * (0.5f + f * (0.166666667f + f * (0.0416666667f + f * (0.00833333333f + f * (0.00138888889f + f * 0.000198412698f))))));
    if (!(nf >= -126.0f)) return 0.0f;                                  /* also NaN */
    if (nf > 127.0f) return INFINITY;
    int32_t n = (int32_t)nf;
    uint32_t bits = (uint32_t)(n + 127) << 23;
    float scale;
    memcpy(&scale, &bits, 4);
    return p * scale;
}
float o_sky_exp(float x) { return o_sky_exp2(x * 1.44269504f); }
/* Abramowitz & Stegun 4.4.46: acos(x) = sqrt(1 - x) * P7(x) on [0, 1] (|error| <= 2e-8), acos(-x) = pi - acos(x) */
float o_sky_acos(float x)
{
    float a = fabsf(x);
    float p = 1.5707963050f + a * (-0.2145988016f + a * (0.0889789874f + a * (-0.0501743046f + a * (0.0308918810f + a * (-0.0170881256f + a * (0.0066700901f + a * -0.0012624911f))))));
    float r = sqrtf(1.0f - a) * p;
    return x < 0.0f ? 3.14159265358979f - r : r;
}

/* sky.glsl:1-16 */
static float radiance_internal(const float c[9], float cos_theta, float gamma, float cos_gamma)
{
    float expM = o_sky_exp(c[4] * gamma);
    float rayM = cos_gamma * cos_gamma;
    float mb = (1.0f + c[8] * c[8]) - (2.0f * c[8]) * cos_gamma;
    float mieM = (1.0f + rayM) / (mb * sqrtf(mb));                       /* pow(x, 1.5) = x sqrt(x): exact operations */
    float zenith = sqrtf(cos_theta);
    return (1.0f + c[0] * o_sky_exp(c[1] / (cos_theta + 0.01f))) *
           ((((c[2] + c[3] * expM) + c[5] * rayM) + c[6] * mieM) + c[7] * zenith);
}

/* sky.glsl:19-78.  Deviation: cos_gamma is clamped to [-1,1] before acos (GLSL leaves acos
 * undefined outside that range; |dir|,|sun| are 1 only to rounding). */
void o_sky_radiance(const OSky *sky, const float dir[3], float out[3])
{
    if (sky->direction[1] <= 0.0f) { out[0] = out[1] = out[2] = 0.0f; return; }
    float cos_theta = fminf(fmaxf(dir[1], 0.0f), 1.0f);
    float cos_gamma = (dir[0] * sky->direction[0] + dir[1] * sky->direction[1]) + dir[2] * sky->direction[2];
    float gamma = o_sky_acos(fminf(fmaxf(cos_gamma, -1.0f), 1.0f));
    float xyz[3];
    xyz[0] = radiance_internal(sky->r.config, cos_theta, gamma, cos_gamma) * sky->r.radiance;
    xyz[1] = radiance_internal(sky->g.config, cos_theta, gamma, cos_gamma) * sky->g.radiance;
    xyz[2] = radiance_internal(sky->b.config, cos_theta, gamma, cos_gamma) * sky->b.radiance;
    for (int c = 0; c < 3; c++) xyz[c] = xyz[c] * 683.0f;
    o_xyz2acescg(xyz, out);
}

/* sky.glsl:80-113 */
void o_sun_radiance(const OSky *sky, const float dir[3], float out[3])
{
    out[0] = out[1] = out[2] = 0.0f;
    float cos_gamma = (dir[0] * sky->direction[0] + dir[1] * sky->direction[1]) + dir[2] * sky->direction[2];
    if (cos_gamma < 0.0f || dir[1] < 0.0f) return;
    float sol_rad_sin = sinf(sky->sunlight_intensity[3]);
    float ar2 = 1.0f / (sol_rad_sin * sol_rad_sin);
    float singamma = 1.0f - (cos_gamma * cos_gamma);
    float sc2 = 1.0f - ar2 * singamma * singamma;
    if (sc2 <= 0.0f) return;
    float sampleCosine = sqrtf(sc2);
    const OSkyChannel *ch[3] = { &sky->r, &sky->g, &sky->b };
    float dark[3];
    for (int c = 0; c < 3; c++) dark[c] = ch[c]->ld_coefficient0;
    for (int c = 0; c < 3; c++) dark[c] += ch[c]->ld_coefficient1 * sampleCosine;
    float cur = sampleCosine;
    for (int i = 0; i < 4; i++) {
        cur *= sampleCosine;
        for (int c = 0; c < 3; c++) dark[c] += ch[c]->ld_coefficient2[i] * cur;
    }
    float xyz[3];
    for (int c = 0; c < 3; c++) xyz[c] = sky->sunlight_intensity[c] * dark[c];
    o_xyz2acescg(xyz, out);
}
