/* ORACLE — test infrastructure only.  See oracle_cache.h: groundwork for SURVEY.md 8f rank 2, no CUDA counterpart yet. */
#include "oracle_cache.h"
#include "oracle_post.h"      /* o_post_log2: the fixed-expression log2 */
#include "oracle_sky.h"       /* o_sky_exp2: the fixed-expression exp2 */
#include <math.h>
#include <string.h>

/* headers/spatial_hash.glsl:105-111 (same function as the path tracer's PRNG) */
static uint32_t pcg(uint32_t v)
{
    uint32_t state = v * 747796405u + 2891336453u;
    uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}

uint32_t o_xxhash32(uint32_t p)                                         /* :115-126 */
{
    const uint32_t PRIME32_2 = 2246822519u, PRIME32_3 = 3266489917u, PRIME32_4 = 668265263u, PRIME32_5 = 374761393u;
    uint32_t h32 = p + PRIME32_5;
    h32 = PRIME32_4 * ((h32 << 17) | (h32 >> (32 - 17)));
    h32 = PRIME32_2 * (h32 ^ (h32 >> 15));
    h32 = PRIME32_3 * (h32 ^ (h32 >> 13));
    return h32 ^ (h32 >> 16);
}

uint32_t o_cache_fingerprint(const OSpatialHashKey *key)                /* :128-135 */
{
    uint32_t hash = o_xxhash32((uint32_t)key->position[0]);
    hash = o_xxhash32((uint32_t)key->position[1] + hash);
    hash = o_xxhash32((uint32_t)key->position[2] + hash);
    hash = o_xxhash32(key->direction + hash);
    return hash > 1u ? hash : 1u;                                       /* max(1, hash): 0 marks an empty entry */
}

uint32_t o_cache_location(const OSpatialHashKey *key, uint32_t capacity)    /* :136-142 */
{
    uint32_t hash = pcg((uint32_t)key->position[0]);
    hash = pcg((uint32_t)key->position[1] + hash);
    hash = pcg((uint32_t)key->position[2] + hash);
    hash = pcg(key->direction + hash);
    return hash % capacity;
}

/* 2^x from mul / add only: oracle_sky.c's o_sky_exp2 (x = n + f, f in [-0.5, 0.5], 2^f by a degree-7 Taylor polynomial in f ln 2) */
#define exp2_fixed o_sky_exp2

static void mat3_mul_cols(const float m[9], const float v[3], float out[3])     /* GLSL mat3(c0, c1, c2) * v */
{
    for (int r = 0; r < 3; r++) out[r] = (m[r] * v[0] + m[3 + r] * v[1]) + m[6 + r] * v[2];
}
static const float ACESCG2XYZ[9] = { 0.66245437f, 0.2722288f, -0.0055746622f, 0.13400422f, 0.6740818f, 0.00406073f, 0.15618773f, 0.05368953f, 1.0103393f };
static const float XYZ2ACESCG[9] = { 1.6410228f, -0.66366285f, 0.011721907f, -0.32480323f, 1.6153315f, -0.0082844375f, -0.23642465f, 0.016756356f, 0.9883947f };

uint32_t o_logluv_encode(const float rgb[3])                            /* :27-58 */
{
    float XYZ[3];
    mat3_mul_cols(ACESCG2XYZ, rgb, XYZ);
    if (!(XYZ[1] > 0.0f)) return 0;                                     /* log2(Y <= 0) = -inf / NaN -> Le = 0 */
    float logY = 409.6f * (o_post_log2(XYZ[1]) + 20.0f);
    float cl = logY < 0.0f ? 0.0f : (logY > 16383.0f ? 16383.0f : logY);
    uint32_t Le = (uint32_t)cl;
    if (Le == 0) return 0;
    float invDenom = 1.0f / ((-2.0f * XYZ[0] + 12.0f * XYZ[1]) + 3.0f * ((XYZ[0] + XYZ[1]) + XYZ[2]));
    float u = (4.0f * XYZ[0]) * invDenom, v = (9.0f * XYZ[1]) * invDenom;
    float su = 820.0f * u, sv = 820.0f * v;
    uint32_t ue = (uint32_t)(su < 0.0f ? 0.0f : (su > 511.0f ? 511.0f : su)), ve = (uint32_t)(sv < 0.0f ? 0.0f : (sv > 511.0f ? 511.0f : sv));
    return (Le << 18) | (ue << 9) | ve;
}

void o_logluv_decode(uint32_t packed, float rgb[3])                     /* :62-93 */
{
    uint32_t Le = packed >> 18;
    if (Le == 0) { rgb[0] = rgb[1] = rgb[2] = 0.0f; return; }
    float logY = ((float)Le + 0.5f) / 409.6f - 20.0f;
    float Y = exp2_fixed(logY);
    float u = ((float)((packed >> 9) & 0x1ffu) + 0.5f) / 820.0f, v = ((float)(packed & 0x1ffu) + 0.5f) / 820.0f;
    float invDenom = 1.0f / ((6.0f * u - 16.0f * v) + 12.0f);
    float x = (9.0f * u) * invDenom, y = (4.0f * v) * invDenom;
    float s = Y / y;
    float XYZ[3] = { s * x, Y, s * ((1.0f - x) - y) }, out[3];
    mat3_mul_cols(XYZ2ACESCG, XYZ, out);
    for (int k = 0; k < 3; k++) rgb[k] = out[k] > 0.0f ? out[k] : 0.0f;
}

uint32_t o_normal_to_face_id(const float n[3])                          /* headers/normal.glsl:9-18 */
{
    float sum = (n[0] + n[1]) + n[2];
    float s = sum < 0.0f ? 0.0f : (sum > 1.0f ? 1.0f : sum);
    uint32_t face = (uint32_t)rintf(s);
    face += (uint32_t)rintf(fabsf(n[2])) * 4u + (uint32_t)rintf(fabsf(n[1])) * 2u;
    return face & 0xFFu;
}

void o_face_id_to_normal(uint32_t face, float n[3])                     /* :20-26 */
{
    n[0] = n[1] = n[2] = 0.0f;
    n[(face >> 1) % 3u] = (float)(face & 1u) * 2.0f - 1.0f;
}

void o_cache_key_of(const float c[3], const float normal_world[3], OSpatialHashKey *key)   /* final_gather.rchit:48-50 */
{
    for (int a = 0; a < 3; a++) key->position[a] = (int32_t)(c[a] / 4.0f);      /* ivec3(): truncation toward zero */
    key->direction = o_normal_to_face_id(normal_world);
}

#define MAX_SAMPLE_COUNT 404u

void o_cache_insert(OSpatialHashEntry *t, uint32_t capacity, const OSpatialHashKey *key, const float value[3], uint32_t frame_index)   /* :150-197 */
{
    const uint32_t fingerprint = o_cache_fingerprint(key), location = o_cache_location(key, capacity);
    uint32_t i_min = 0, min_frame = 0;
    for (uint32_t i = 0; i < 3; i++) {
        OSpatialHashEntry *e = &t[location + i];
        uint32_t current = e->fingerprint;                              /* atomicCompSwap(fingerprint, 0, new): sequential form */
        if (current == 0) e->fingerprint = fingerprint;
        uint32_t frame = e->last_accessed_frame;
        if (i == 0 || frame < min_frame) { i_min = i; min_frame = frame; }
        if (current == fingerprint || current == 0) {
            float cur[3] = { 0.0f, 0.0f, 0.0f };
            uint32_t count = 0;
            if (current == fingerprint) { count = e->sample_count; o_logluv_decode(e->radiance, cur); }
            if (count > MAX_SAMPLE_COUNT - 1u) count = MAX_SAMPLE_COUNT - 1u;
            const uint32_t next = count + 1u;
            const float a = 1.0f / (float)next;
            float mixed[3];
            for (int k = 0; k < 3; k++) mixed[k] = cur[k] * (1.0f - a) + value[k] * a;      /* mix(x, y, a) = x (1 - a) + y a */
            e->radiance = o_logluv_encode(mixed);
            e->last_accessed_frame = (uint16_t)frame_index;
            e->sample_count = (uint16_t)next;
            return;
        }
    }
    OSpatialHashEntry *e = &t[location + i_min];                        /* evict the least recently accessed of the three */
    e->fingerprint = fingerprint;
    e->radiance = o_logluv_encode(value);
    e->last_accessed_frame = (uint16_t)frame_index;
    e->sample_count = 1;
}

int o_cache_get(OSpatialHashEntry *t, uint32_t capacity, const OSpatialHashKey *key, uint32_t frame_index, float value[3], uint32_t *sample_count)   /* :202-220 */
{
    const uint32_t fingerprint = o_cache_fingerprint(key), location = o_cache_location(key, capacity);
    value[0] = value[1] = value[2] = 0.0f;
    *sample_count = 0;
    for (uint32_t i = 0; i < 3; i++) {
        OSpatialHashEntry *e = &t[location + i];
        if (e->fingerprint == 0) return 0;                              /* an empty entry ends the probe */
        if (e->fingerprint == fingerprint) {
            e->last_accessed_frame = (uint16_t)frame_index;
            o_logluv_decode(e->radiance, value);
            *sample_count = e->sample_count;
            return 1;
        }
    }
    return 0;
}
