/* ORACLE — test infrastructure only.  See oracle_post.h for the reference files restated here. */
#include "oracle_post.h"
#include "oracle_sky.h"
#include <math.h>
#include <string.h>

float o_half_to_float(uint16_t h)
{
    uint32_t s = (uint32_t)(h >> 15) << 31, e = (h >> 10) & 31u, m = h & 1023u, bits;
    if (e == 0) {
        if (m == 0) bits = s;
        else {                                        /* subnormal half: value m * 2^-24, normalise */
            int k = 0;
            while (!(m & 1024u)) { m <<= 1; k++; }
            bits = s | ((uint32_t)(113 - k) << 23) | ((m & 1023u) << 13);
        }
    } else if (e == 31) bits = s | 0x7F800000u | (m << 13);
    else bits = s | ((e + 112u) << 23) | (m << 13);
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

uint16_t o_float_to_half(float f)
{
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t s = (x >> 16) & 0x8000u, a = x & 0x7FFFFFFFu;
    if (a >= 0x7F800000u) return (uint16_t)(s | 0x7C00u | (a > 0x7F800000u ? 0x200u : 0u));
    if (a >= 0x477FF000u) return (uint16_t)(s | 0x7C00u);                  /* >= 65520 rounds to inf */
    if (a < 0x33000001u) return (uint16_t)s;                               /* <= 2^-25 rounds to zero */
    int32_t e = (int32_t)(a >> 23) - 127;
    uint32_t m = (a & 0x7FFFFFu) | 0x800000u;
    uint32_t shift = e < -14 ? (uint32_t)(13 + (-14 - e)) : 13u;           /* bits dropped */
    uint32_t q = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) q++;
    uint32_t h = e < -14 ? q : (((uint32_t)(e + 15) << 10) + (q - 1024u)); /* carry out of the mantissa bumps the exponent */
    return (uint16_t)(s | h);
}

/* log2 of a positive normal float from mul / add / div only (identical on the GPU):
 * x = m * 2^e, m in [sqrt(1/2), sqrt(2));  log2(m) = 2/ln2 * atanh(s),  s = (m - 1) / (m + 1), |s| < 0.172 */
float o_post_log2(float x)
{
    uint32_t bits;
    memcpy(&bits, &x, 4);
    int32_t e = (int32_t)(bits >> 23) - 127;
    uint32_t mb = (bits & 0x7FFFFFu) | 0x3F800000u;
    float m;
    memcpy(&m, &mb, 4);
    if (m > 1.41421356f) { m = m * 0.5f; e += 1; }
    float s = (m - 1.0f) / (m + 1.0f), s2 = s * s;
    float p = 1.0f + s2 * (0.333333333f + s2 * (0.2f + s2 * (0.142857143f + s2 * 0.111111111f)));
    return (float)e + (s * 2.88539008f) * p;
}

static void ycocg_to_linear(const float c[3], float r[3])          /* _NRD_YCoCgToLinear */
{
    float t = c[0] - c[2];
    r[1] = fmaxf(c[0] + c[2], 0.0f);
    r[0] = fmaxf(t + c[1], 0.0f);
    r[2] = fmaxf(t - c[1], 0.0f);
}

uint32_t o_color_to_bin(const float rgb[3], const OExposureParams *p)          /* auto_exposure.comp:22-38 */
{
    float lum = (rgb[0] * 0.299f + rgb[1] * 0.587f) + rgb[2] * 0.114f;
    if (lum < 0.005f) return 0;
    float v = (o_post_log2(lum) - p->min_log_lum) * (1.0f / p->log_lum_range);
    float logLum = fminf(fmaxf(v, 0.0f), 1.0f);
    return (uint32_t)(logLum * 254.0f + 1.0f);
}

void o_exposure_histogram(const uint16_t *img, uint32_t width, uint32_t height, const OExposureParams *p, OHistogram *h)
{
    for (uint64_t i = 0; i < (uint64_t)width * height; i++) {
        float c[3] = { o_half_to_float(img[4 * i]), o_half_to_float(img[4 * i + 1]), o_half_to_float(img[4 * i + 2]) }, rgb[3];
        ycocg_to_linear(c, rgb);
        h->histogram[o_color_to_bin(rgb, p)]++;
    }
}

void o_exposure_average(OHistogram *h, uint32_t width, uint32_t height, const OExposureParams *p)      /* auto_exposure_avg.comp:19-51 */
{
    uint32_t sum = 0;                                                   /* uint arithmetic, like the shared-memory reduction */
    for (uint32_t i = 0; i < 256; i++) { sum += h->histogram[i] * i; h->histogram[i] = 0; }
    float num_pixels = (float)(width * height);
    float weightedLogAverage = ((float)sum / fmaxf(num_pixels, 1.0f)) - 1.0f;
    float weightedAvgLum = exp2f(((weightedLogAverage / 254.0f) * p->log_lum_range) + p->min_log_lum);
    float last = h->avg;
    h->avg = last + (weightedAvgLum - last) * p->time_coeff;
}

static float linear_to_srgb(float c) { return c <= 0.0031308f ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f; }
static float rrt_odt(float v)                                                       /* tone_map.comp:36-40 */
{
    float a = v * (v + 0.0245786f) - 0.000090537f;
    float b = v * (0.983729f * v + 0.4329510f) + 0.238081f;
    return a / b;
}
static float dotc(const float v[3], float a, float b, float c) { return (v[0] * a + v[1] * b) + v[2] * c; }

static float oetf1(float c, uint32_t f)                                             /* tone_map.comp:70-167 */
{
    switch (f) {
    case 1: return linear_to_srgb(c);
    case 2: return c <= -0.0031308f ? -1.055f * powf(-c, 1.0f / 2.4f) + 0.055f : linear_to_srgb(c);
    case 3: return powf(c / 52.37f, 1.0f / 2.6f);
    case 4: return c < 0.0030186f ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
    case 5: return c < 0.0181f ? 4.5f * c : 1.0993f * powf(c, 0.45f) - (1.0993f - 1.0f);
    case 6: {
        const float m1 = 2610.0f / 16384.0f, m2 = (2523.0f / 4096.0f) * 128.0f, c2 = (2413.0f / 4096.0f) * 32.0f, c3 = (2392.0f / 4096.0f) * 32.0f;
        const float c1 = c3 - c2 + 1.0f;
        float Lm = powf(c, m1);
        return powf((c1 + c2 * Lm) / (1.0f + c3 * Lm), m2);
    }
    case 7: {
        const float a = 0.17883277f, b = 1.0f - 4.0f * a, cc = 0.55991073f;
        return c < (1.0f / 12.0f) ? sqrtf(3.0f * c) : a * logf(12.0f * c - b) + cc;
    }
    case 8: return powf(c, 256.0f / 563.0f);
    default: return c;
    }
}

void o_tone_map(const uint16_t *img, const uint32_t *albedo, uint32_t width, uint32_t height, float avg, const float csc[9],
                uint32_t oetf, float *dst)
{
    float exposure = 1.0f / (9.6f * avg);                                           /* tone_map.comp:196-197 */
    exposure = exposure * 9.6f;
    for (uint64_t i = 0; i < (uint64_t)width * height; i++) {
        float c[3] = { o_half_to_float(img[4 * i]), o_half_to_float(img[4 * i + 1]), o_half_to_float(img[4 * i + 2]) }, in[3], s[3], m[3], v[3];
        ycocg_to_linear(c, in);
        uint32_t a = albedo[i];
        float al[3] = { (float)((a >> 20) & 1023u) / 1023.0f, (float)((a >> 10) & 1023u) / 1023.0f, (float)(a & 1023u) / 1023.0f };
        o_acescg2srgb(in, s);
        for (int k = 0; k < 3; k++) s[k] = s[k] * o_srgb_to_linear(al[k])                    /* color.glsl:1-5 */;
        o_srgb2acescg(s, m);
        for (int k = 0; k < 3; k++) m[k] = m[k] * exposure;                         /* :200 */
        /* mat3(c0..c8) * v, column-major (:202-209) */
        v[0] = (csc[0] * m[0] + csc[3] * m[1]) + csc[6] * m[2];
        v[1] = (csc[1] * m[0] + csc[4] * m[1]) + csc[7] * m[2];
        v[2] = (csc[2] * m[0] + csc[5] * m[1]) + csc[8] * m[2];
        /* ACESFitted (:42-68): v *= M is the row vector times the matrix: component j = dot(v, column j) */
        float r[3] = { dotc(v, 0.59719f, 0.35458f, 0.04823f), dotc(v, 0.07600f, 0.90834f, 0.01566f), dotc(v, 0.02840f, 0.13383f, 0.83777f) };
        for (int k = 0; k < 3; k++) r[k] = rrt_odt(r[k]);
        float o[3] = { dotc(r, 1.60475f, -0.53108f, -0.07367f), dotc(r, -0.10208f, 1.10813f, -0.00605f), dotc(r, -0.00327f, -0.07276f, 1.07602f) };
        for (int k = 0; k < 3; k++) dst[4 * i + k] = oetf1(o[k], oetf);
        dst[4 * i + 3] = 1.0f;
    }
}
