/* ORACLE — test infrastructure only (see oracle_trace.h).  CPU restatement of the reference's post chain
 * (SURVEY.md 8f rank 3):
 *   luminance histogram   assets/shaders/auto_exposure.comp:22-72
 *   adapted luminance     assets/shaders/auto_exposure_avg.comp:19-51   (buffer layout: auto_exposure.playout)
 *   tone map + OETF       assets/shaders/tone_map.comp:36-219
 * Images: illuminance = R16G16B16A16_SFLOAT holding REBLUR_FrontEnd_PackRadianceAndNormHitDist's YCoCg
 * (headers/nrd.glsl:99-147), albedo = A2R10G10B10_UNORM_PACK32 (standard.rs:880-917), dst = float RGBA.
 * Parity status: PARITY UNPINNED — the reference holds no test, golden image or fixture for these passes and cannot be
 * run here; tests/test_post.py cross-checks this restatement against independent float64 formulas of the same shaders.
 * One documented deviation: log2() of the histogram is a fixed mul/add/div expression (o_post_log2) so
 * bins are bit-reproducible on the GPU; GLSL leaves log2's precision to the implementation. */
#ifndef ORACLE_POST_H
#define ORACLE_POST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct OExposureParams { float min_log_lum, log_lum_range, time_coeff; } OExposureParams;   /* auto_exposure.comp:7-11 */
typedef struct OHistogram { uint32_t histogram[256]; float avg; } OHistogram;                       /* auto_exposure.playout */

float    o_half_to_float(uint16_t h);
uint16_t o_float_to_half(float f);                 /* round to nearest even */
float    o_post_log2(float x);
uint32_t o_color_to_bin(const float rgb[3], const OExposureParams *p);
void     o_exposure_histogram(const uint16_t *illuminance_rgba16f, uint32_t width, uint32_t height, const OExposureParams *p, OHistogram *h);
void     o_exposure_average(OHistogram *h, uint32_t width, uint32_t height, const OExposureParams *p);
/* csc: the nine COLOR_SPACE_CONVERSION_n specialisation constants in order; oetf: OUTPUT_TRANSFER_FUNCTION */
void     o_tone_map(const uint16_t *illuminance_rgba16f, const uint32_t *albedo_a2r10g10b10, uint32_t width, uint32_t height,
                    float avg, const float csc[9], uint32_t oetf, float *dst_rgba32f);

#ifdef __cplusplus
}
#endif
#endif
