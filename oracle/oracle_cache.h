/* ORACLE — test infrastructure only.  CPU restatement of the reference's world-space radiance cache, the data structure behind
 * SURVEY.md 8f rank 2 (the far-field final gather): assets/shaders/headers/spatial_hash.glsl
 *   LogLuv radiance codec         :21-93      (EncodeRGBToLogLuv / DecodeLogLuvToRGB)
 *   key fingerprint / location    :105-145    (pcg, xxhash32, SpatialHashKeyGetFingerprint / GetLocation)
 *   insert / get                  :150-220    (3 probes, running mean capped at MAX_SAMPLE_COUNT = 404, LRU eviction)
 * and headers/normal.glsl:9-26 (normal2FaceID / faceId2Normal).  Entry layout: SpatialHashEntry of headers/layout.playout
 * (scalar layout, 12 bytes).
 *
 * STATUS: groundwork for the next round — NO CUDA counterpart exists yet and nothing in the product uses a radiance cache.
 * The reference's insert is a racy read-modify-write executed by thousands of invocations; this file fixes the SEQUENTIAL
 * semantics (inserts applied one after the other, in call order) that a deterministic device-side insert will have to
 * reproduce.  Parity status: PARITY UNPINNED (no reference test or fixture touches the cache); tests/test_oracle_cache.py
 * cross-checks every function against an independent numpy restatement.
 * Two documented deviations, as elsewhere: log2 / exp2 of the codec are fixed mul/add/div expressions (bit-reproducible on a
 * GPU); `last_accessed_frame` / `sample_count` are the reference's uint16 fields with its wrap-around. */
#ifndef ORACLE_CACHE_H
#define ORACLE_CACHE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct OSpatialHashEntry {        /* headers/layout.playout: SpatialHashEntry */
    uint32_t fingerprint;
    uint32_t radiance;                    /* LogLuv */
    uint16_t last_accessed_frame;
    uint16_t sample_count;
} OSpatialHashEntry;

typedef struct OSpatialHashKey { int32_t position[3]; uint32_t direction; } OSpatialHashKey;   /* spatial_hash.glsl:98-101 */

uint32_t o_xxhash32(uint32_t p);
uint32_t o_cache_fingerprint(const OSpatialHashKey *key);
uint32_t o_cache_location(const OSpatialHashKey *key, uint32_t capacity);
uint32_t o_logluv_encode(const float rgb[3]);
void     o_logluv_decode(uint32_t packed, float rgb[3]);
uint32_t o_normal_to_face_id(const float n[3]);
void     o_face_id_to_normal(uint32_t face, float n[3]);
/* key of a far hit: block centre in world space / 4, face of the larger voxel (final_gather.rchit:48-50) */
void     o_cache_key_of(const float block_center_world[3], const float normal_world[3], OSpatialHashKey *key);
/* `table` has capacity + 2 entries (a probe window may run over the end, like the reference's unchecked location + i) */
void     o_cache_insert(OSpatialHashEntry *table, uint32_t capacity, const OSpatialHashKey *key, const float value[3], uint32_t frame_index);
int      o_cache_get(OSpatialHashEntry *table, uint32_t capacity, const OSpatialHashKey *key, uint32_t frame_index, float value[3], uint32_t *sample_count);

#ifdef __cplusplus
}
#endif
#endif
