/* ORACLE — test infrastructure only.  CPU restatement of the reference's own global illumination for one frame:
 *   primary pass        assets/shaders/primary/{primary.rgen, hit.rchit, miss.rmiss}                 (oracle_trace.c's G-buffer)
 *   ambient occlusion   final_gather/ambient_occlusion.{rgen:14-66, rint:46-134, rchit, rmiss}, final_gather/nee.rmiss:11-22
 *   final gather        final_gather/final_gather.{rgen:14-52, rchit:35-91, rmiss:12-24}, final_gather/rough.rint:42-59
 *   surfel pass         surfel/{surfel.rgen:12-67, surfel.rchit:35-102, surfel.rmiss:14-26, nee.rmiss:15-27}
 *   radiance cache      headers/spatial_hash.glsl (oracle_cache.c), persistent state of crates/render/src/pipeline/standard.rs:334-358
 * The passes talk to each other through the RGBA16F `img_illuminance` image exactly as in the reference: YCoCg + normHitDist,
 * rounded to half floats between passes (headers/nrd.glsl:99-147).
 *
 * The reference runs thousands of invocations that read and write the cache and the surfel pool concurrently (racy read-modify-
 * write, last writer wins).  This file fixes ONE of the interleavings the reference admits, chosen so that a GPU can reproduce it:
 *   - a pass first traces all of its rays and performs all of its cache LOOK-UPS against the table as the previous pass left it
 *     (a look-up only stamps `last_accessed_frame` with this frame's index: every writer stores the same value),
 *   - then its cache INSERTS are applied one after the other in ascending invocation index (surfel index),
 *   - surfel-pool writes of a pass are applied in ascending invocation index (the highest pixel that maps to a slot wins).
 * Documented deviations: (3) pcg streams instead of the STBN blue-noise textures, (5)/(6) fixed log2 / exp2 / exp / acos
 * expressions, as everywhere else.  PARITY UNPINNED: no reference test or fixture touches these passes. */
#ifndef ORACLE_GI_H
#define ORACLE_GI_H
#include "oracle_trace.h"
#include "oracle_cache.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct OSurfel { float position[3]; uint32_t direction; } OSurfel;      /* headers/layout.playout: SurfelEntry */

typedef struct OGiState {
    OSpatialHashEntry *table;      /* capacity + 2 entries (a probe window may run over the end) */
    uint32_t capacity;
    OSurfel *pool;                 /* pool_size entries; direction >= 6 marks an empty slot (the reference fills the buffer with 0xFF) */
    uint32_t pool_size;
} OGiState;

typedef struct OGiStats {
    uint64_t near_hits;            /* pixels whose ambient-occlusion ray hit within AMBIENT_OCCLUSION_THRESHOLD */
    uint64_t far_hits, far_misses; /* final-gather rays */
    uint64_t cache_found;          /* final-gather + surfel look-ups that found their key */
    uint64_t surfels_scheduled;    /* pool slots written this frame */
    uint64_t surfel_rays;          /* valid pool entries traced */
    uint64_t cache_inserts;
    uint64_t insert_segments;      /* (device only: independent groups of inserts; the oracle leaves 0) */
} OGiStats;

/* One frame.  cfg: width, height, seed (push_constants.rand), frame_index.  `illuminance`: width*height*4 half floats
 * (img_illuminance for hit pixels; img_illuminance_denoised = pack((sky + sun) / 3.14, 100000) for miss pixels).
 * `gb`: G-buffer of the primary hit; normal and depth are REQUIRED (the later passes read them back, like the reference). */
void o_gi_frame(const OScene *s, const OCamera *cam, const OSky *sky, const OTraceConfig *cfg, OGiState *state,
                uint16_t *illuminance, const OGBuffer *gb, OGiStats *stats, uint32_t n_threads);

/* building blocks for unit tests */
void     o_gi_unpack_normal(uint32_t packed_a2r10g10b10, float n[3]);     /* NRD_FrontEnd_UnpackNormalAndRoughness(...).xyz */
int      o_leaf_ao(const float origin[3], const float dir[3], uint64_t grid, float ray_tmin, float *t, uint32_t *voxel);   /* ambient_occlusion.rint:46-134 */
#ifdef __cplusplus
}
#endif
#endif
