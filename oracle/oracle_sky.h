/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under dust_b200/ may include, link or call this.
 *
 * CPU restatement (plain C) of the reference's Hosek-Wilkie sky: host bake
 * crates/render/src/pipeline/sky.rs:90-268 and shader evaluation
 * assets/shaders/headers/sky.glsl:1-113 + headers/color.glsl.
 * Parity status: PARITY UNPINNED (no reference test touches the sky); the coefficient tables
 * are the reference's own dataset.bin / datasetSolar.bin (data/hosek_wilkie_xyz_tables.h).
 */
#ifndef ORACLE_SKY_H
#define ORACLE_SKY_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* SkyModelChannelState, sky.rs:66-75 (std430: 16 floats) */
typedef struct OSkyChannel {
    float config[9];       /* config1.xyzw, config2.xyzw, config3 */
    float radiance;
    float ld_coefficient0;
    float ld_coefficient1;
    float ld_coefficient2[4];
} OSkyChannel;

/* SkyModelState, sky.rs:78-85 (224 bytes) */
typedef struct OSky {
    OSkyChannel r, g, b;
    float direction[4];
    float sunlight_intensity[4];   /* xyz = solar radiance (CIE XYZ), w = solar angular radius */
} OSky;

/* Sunlight::bake, sky.rs:90-133 */
void o_sunlight_bake(float turbidity, const float albedo[3], const float direction[3], OSky *out);
/* sky.glsl:21-78 — `dir` normalised; result in ACEScg */
/* fixed-expression exp2 / exp / acos (deviation (6)): the kernel evaluates the same operations in the same order */
float o_sky_exp2(float x);
float o_sky_exp(float x);
float o_sky_acos(float x);
void o_sky_radiance(const OSky *sky, const float dir[3], float out[3]);
/* sky.glsl:80-113 */
void o_sun_radiance(const OSky *sky, const float dir[3], float out[3]);

/* color.glsl */
float o_srgb_to_linear(float c);
void  o_srgb2acescg(const float in[3], float out[3]);
void  o_acescg2srgb(const float in[3], float out[3]);
void  o_xyz2acescg(const float in[3], float out[3]);

#ifdef __cplusplus
}
#endif
#endif
