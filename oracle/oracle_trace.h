/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under dust_b200/ may include, link or call this.
 *
 * CPU restatement (plain C, pthreads over pixel rows) of the hot path: camera rays, closest-hit /
 * any-hit traversal of instanced `crates/vdb` trees, and per-hit radiance accumulation.
 *
 * What is restated literally from the reference (file:line in the .c):
 *   camera ray         assets/shaders/headers/camera.glsl:1-16, primary/primary.rgen:9-21
 *   LEAF STEP          assets/shaders/primary/hit.rint:20-28,43-131 (4x4x4 Amanatides-Woo DDA)
 *   normal / albedo    assets/shaders/primary/hit.rchit:16-95, headers/normal.glsl:31-43
 *   sky / sun / colour headers/sky.glsl, headers/color.glsl, primary/miss.rmiss:7-17
 *   bounce + NEE       final_gather/ambient_occlusion.rgen:14-66, nee.rmiss:11-22,
 *                      final_gather.rgen:14-52, final_gather.rmiss:12-24, final_gather.rchit:65-89
 *   final radiance     assets/shaders/tone_map.comp:192-203
 * What the reference does NOT contain and is therefore DEFINED here (DESIGN.md "Traversal
 * semantics"): everything above the leaf — the reference hands root/internal nodes to the Vulkan
 * driver's BVH.  The HDDA over InternalNode masks, the instance loop and the
 * PRNG (pcg, headers/spatial_hash.glsl:105-111, instead of STBN textures) are this oracle's own
 * operational definition; the CUDA path must reproduce it bit-for-bit.
 *
 * Parity status: PARITY UNPINNED for traversal / hit ids / t / radiance — no reference test,
 * golden image or fixture exists for them and the reference cannot be built here
 * (Rust nightly + Vulkan RT).  Independent cross-check: oracle_dense.c (brute-force dense DDA).
 *
 * Documented deviations from the GLSL (DESIGN.md 2 has all eight; the ones that live in this file): (1) object-space direction components with
 * |d| < 2^-40 are replaced by copysign(2^-40, d) (GLSL: undefined inf/NaN); (2) the leaf DDA
 * returns "no hit" when a tie-step leaves the 0..3 cell range (GLSL: garbage index);
 * (3) counter-based pcg PRNG instead of blue-noise textures; (6) exp / acos / pow(x, 1.5) of the sky model are fixed mul / add / sqrt
 * expressions (oracle_sky.c), identical in the kernels, so radiance is bit-identical on both sides.
 * Independent evidence beyond the dense marcher's single model: tests/test_oracle_trace.py traces 10^5 rays through sheared multi-instance
 * scenes and a 4-level tree and classifies every disagreement (coplanar-instance tie / float tie / none unexplained).
 */
#ifndef ORACLE_TRACE_H
#define ORACLE_TRACE_H
#include "oracle_vdb.h"
#include "oracle_sky.h"

#ifdef __cplusplus
extern "C" {
#endif

/* CameraSettings, crates/render/src/pipeline/standard.rs:813-827 (std430, 192 bytes) */
typedef struct OCamera {
    float view_proj[16];
    float inverse_view_proj[16];
    float camera_view_col0[3]; float position_x;
    float camera_view_col1[3]; float position_y;
    float camera_view_col2[3]; float position_z;
    float tan_half_fov, far, near, _padding;
} OCamera;

typedef struct OScene OScene;

typedef struct OTraceConfig {
    uint32_t width, height;
    uint32_t spp;
    uint32_t max_bounces;
    uint32_t seed, frame_index;
    uint32_t y0, y1;             /* rows rendered: [y0, y1) */
    uint32_t flags;              /* O_TRACE_* */
} OTraceConfig;
#define O_TRACE_DEMODULATE 1u    /* radiance without the primary-hit albedo factor (img_illuminance; tone_map.comp:192-203 multiplies it back) */

typedef struct OStats {
    uint64_t primary_rays, bounce_rays, shadow_rays;
    uint64_t mask_words;         /* W of SURVEY.md §8d */
    uint64_t child_ptrs;         /* P */
    uint64_t hits_shaded;        /* H */
} OStats;

typedef struct ORay { float o[3]; float tmin; float d[3]; float tmax; } ORay;        /* 32 bytes */
typedef struct OHitRec { float t; uint32_t instance, leaf, voxel; } OHitRec;         /* 16 bytes */

#define O_MISS 0xFFFFFFFFu

/* trees[i] may repeat; materials[i] / palettes[i] belong to trees[i] */
OScene *o_scene_create(uint32_t n, const float *obj2world /* n*12 */, const OTree *const *trees,
                       const uint8_t *const *materials, const uint8_t *const *palettes /* 255*4 RGBA */);
void    o_scene_destroy(OScene *s);
/* 1 (default): skip instances whose padded world box the ray cannot touch; 0: test every instance */
void    o_scene_set_cull(OScene *s, int enabled);

/* Full path: fills radiance (float4 per pixel, row-major W*H) and, when non-NULL, the sample-0
 * primary hit record arrays.  Only rows [y0,y1) are touched. */
void o_render(const OScene *s, const OCamera *cam, const OSky *sky, const OTraceConfig *cfg,
              float *radiance, uint32_t *hit_instance, uint32_t *hit_leaf, uint8_t *hit_voxel, float *hit_t,
              OStats *stats, uint32_t n_threads);
/* The G-buffer images of the sample-0 primary hit (assets/shaders/primary/hit.rchit:57-94, miss.rmiss:13-16;
 * formats crates/render/src/pipeline/standard.rs:880-917).  NULL members are skipped; on a miss normal and
 * voxel_id stay untouched, like in the reference. */
typedef struct OGBuffer {
    uint32_t *albedo;    /* A2R10G10B10_UNORM_PACK32 */
    float *depth;        /* R32_SFLOAT */
    uint32_t *normal;    /* A2R10G10B10_UNORM_PACK32, NRD_FrontEnd_PackNormalAndRoughness */
    float *motion;       /* xyz0 as float32 (the image is R16G16B16A16_SFLOAT: the caller rounds to half) */
    uint32_t *voxel_id;  /* R32_UINT */
} OGBuffer;

/* last frame's object->world matrices (n x 12), the `instances[]` buffer of hit.rchit:92; default: this frame's */
void o_scene_set_previous(OScene *s, const float *obj2world_prev);

void o_render_g(const OScene *s, const OCamera *cam, const OSky *sky, const OTraceConfig *cfg,
                const uint32_t *rows, uint32_t n_rows,
                float *radiance, uint32_t *hit_instance, uint32_t *hit_leaf, uint8_t *hit_voxel, float *hit_t,
                const OGBuffer *gbuffer, OStats *stats, uint32_t n_threads);

/* the same for an explicit list of rows (bounded CPU samples spread over the frame) */
void o_render_rows(const OScene *s, const OCamera *cam, const OSky *sky, const OTraceConfig *cfg,
                   const uint32_t *rows, uint32_t n_rows,
                   float *radiance, uint32_t *hit_instance, uint32_t *hit_leaf, uint8_t *hit_voxel, float *hit_t,
                   OStats *stats, uint32_t n_threads);

/* Explicit ray batch, closest hit (BASELINE config 5) */
/* rough.rint: leaves are solid 4x4x4 boxes (hit = entry into the box, voxel 0).  Process-wide; affects every later trace call. */
void o_set_rough_blocks(int enabled);
void o_trace_rays(const OScene *s, const ORay *rays, uint64_t n, OHitRec *hits, OStats *stats, uint32_t n_threads);

/* debug aid: every ray of one pixel's paths + the oracle's answers (voxel bit 31 = any-hit ray) */
uint32_t o_debug_pixel(const OScene *s, const OCamera *cam, const OSky *sky, const OTraceConfig *cfg,
                       uint32_t px, uint32_t py, ORay *rays, OHitRec *hits, uint32_t cap, float rgba[4]);

/* building blocks exposed for unit tests */
int   o_leaf_dda(const float origin[3], const float dir[3], uint64_t grid, float ray_tmin, float *t, uint32_t *voxel);
int   o_leaf_ao(const float origin[3], const float dir[3], uint64_t grid, float ray_tmin, float *t, uint32_t *voxel);
void  o_camera_ray(const OCamera *cam, uint32_t width, uint32_t height, uint32_t px, uint32_t py,
                   float jx, float jy, float o[3], float d[3]);
uint32_t o_pcg(uint32_t v);
void  o_cossin_2pi(float u, float *c, float *s);
void  o_rotate_by_normal(const float n[3], const float v[3], float out[3]);
void  o_world2obj(const float obj2world[12], float w2o[12]);

#ifdef __cplusplus
}
#endif
#endif
