/* ORACLE — test infrastructure only.  See oracle_gi.h (scope, citations, the fixed interleaving, parity status).
 * Textually included at the end of oracle_trace.c: the passes below use its static traversal (trace_closest / trace_any with the
 * leaf mode switch) and helpers. */
#include "oracle_gi.h"
#include "oracle_post.h"

/* half floats (the storage format of img_illuminance, R16G16B16A16_SFLOAT, round to nearest even): oracle_post.c's o_float_to_half / o_half_to_float */

/* headers/nrd.glsl:130-147 REBLUR_FrontEnd_PackRadianceAndNormHitDist + the image store */
static void store_illuminance(uint16_t *img, uint32_t pi, const float radiance[3], float norm_hit_dist)
{
    if (norm_hit_dist != 0.0f) norm_hit_dist = fmaxf(norm_hit_dist, 1e-7f);
    const float Y = (radiance[0] * 0.25f + radiance[1] * 0.5f) + radiance[2] * 0.25f;          /* _NRD_LinearToYCoCg */
    const float Co = (radiance[0] * 0.5f + radiance[1] * 0.0f) + radiance[2] * -0.5f;
    const float Cg = (radiance[0] * -0.25f + radiance[1] * 0.5f) + radiance[2] * -0.25f;
    uint16_t *p = img + 4 * (size_t)pi;
    p[0] = o_float_to_half(Y); p[1] = o_float_to_half(Co); p[2] = o_float_to_half(Cg); p[3] = o_float_to_half(norm_hit_dist);
}
/* the image load + headers/nrd.glsl:110-126 REBLUR_BackEnd_UnpackRadianceAndNormHitDist */
static float load_illuminance(const uint16_t *img, uint32_t pi, float radiance[3])
{
    const uint16_t *p = img + 4 * (size_t)pi;
    const float Y = o_half_to_float(p[0]), Co = o_half_to_float(p[1]), Cg = o_half_to_float(p[2]);
    const float t = Y - Cg;
    radiance[1] = fmaxf(Y + Cg, 0.0f);
    radiance[0] = fmaxf(t + Co, 0.0f);
    radiance[2] = fmaxf(t - Co, 0.0f);
    return o_half_to_float(p[3]);
}

/* texel load of the A2R10G10B10_UNORM normal image + headers/nrd.glsl:54-94 NRD_FrontEnd_UnpackNormalAndRoughness(...).xyz */
void o_gi_unpack_normal(uint32_t packed, float n[3])
{
    float px = (float)((packed >> 20) & 1023u) / 1023.0f, py = (float)((packed >> 10) & 1023u) / 1023.0f;
    px = px * 2.0f - 1.0f; py = py * 2.0f - 1.0f;                                            /* _NRD_DecodeUnitVector, bSigned = false */
    float v[3] = { px, py, (1.0f - fabsf(px)) - fabsf(py) };
    const float t = fminf(fmaxf(-v[2], 0.0f), 1.0f);
    v[0] -= t * ((v[0] < 0.0f ? 0.0f : 1.0f) * 2.0f - 1.0f);
    v[1] -= t * ((v[1] < 0.0f ? 0.0f : 1.0f) * 2.0f - 1.0f);
    normalize3(v, n);                                                                        /* :85 */
}

/* the cosine-weighted unit vector the reference fetches from its blue-noise texture (deviation (3): two pcg uniforms) */
static void cosine_sample(uint32_t s1, float out[3])
{
    const float u1 = u01(o_pcg(s1 + 2u)), u2 = u01(o_pcg(s1 + 3u));
    const float r = sqrtf(u1), cz = sqrtf(1.0f - u1);
    float c, sn;
    o_cossin_2pi(u2, &c, &sn);
    out[0] = r * c; out[1] = r * sn; out[2] = cz;
}

/* Block.avg_albedo of a leaf: crates/vox/src/geometry.rs:86-113 (mean palette colour, sRGB-encoded, R10G10B10A2) */
static uint32_t leaf_avg_albedo(const OInstance *in, uint32_t leaf_ptr)
{
    const uint8_t *leaf = o_pool_get(&in->tree->pool[0], leaf_ptr);
    uint64_t mask; uint32_t material_ptr;
    memcpy(&mask, leaf, 8); memcpy(&material_ptr, leaf + 16, 4);
    uint32_t col[4] = { 0, 0, 0, 0 };
    const uint32_t nv = (uint32_t)__builtin_popcountll(mask);
    for (uint32_t i = 0; i < nv; i++) {
        const uint8_t *c = in->palette + 4 * (size_t)in->materials[material_ptr + i];
        for (int k = 0; k < 4; k++) col[k] += c[k];
    }
    const float denom = (float)nv * 255.0f;
    float e[3];
    for (int k = 0; k < 3; k++) {
        const float c = (float)col[k] / denom;
        e[k] = c <= 0.0031308f ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
    }
    const float cw = (float)col[3] / denom;
    return ((uint32_t)(e[0] * 1023.0f) << 22) | ((uint32_t)(e[1] * 1023.0f) << 12) | ((uint32_t)(e[2] * 1023.0f) << 2) | (uint32_t)(cw * 3.0f);
}

/* final_gather.rchit:38-47 / surfel.rchit:36-47: the block a block-granular hit landed on -> world centre, face, cache key */
typedef struct BlockHit { float center[3]; uint32_t face; OSpatialHashKey key; uint32_t avg_albedo; } BlockHit;

static void block_of_hit(const OScene *s, const SceneHit *h, BlockHit *b)
{
    const OInstance *in = &s->inst[h->instance];
    const InstHit *ih = &h->ih;
    float center_obj[3], normal_obj[3], nw[3], av[3], normal_world[3];
    for (int a = 0; a < 3; a++) {
        center_obj[a] = (float)ih->leaf_min[a] + 2.0f;                                       /* :38 */
        normal_obj[a] = (ih->t * ih->d[a] + ih->o[a]) - center_obj[a];                        /* :39-40 */
    }
    const float *W = in->obj2world;
    for (int a = 0; a < 3; a++) {
        nw[a] = (W[4 * a] * normal_obj[0] + W[4 * a + 1] * normal_obj[1]) + W[4 * a + 2] * normal_obj[2];       /* gl_ObjectToWorldEXT * vec4(n, 0) */
        b->center[a] = ((W[4 * a] * center_obj[0] + W[4 * a + 1] * center_obj[1]) + W[4 * a + 2] * center_obj[2]) + W[4 * a + 3];   /* :42 */
        av[a] = fabsf(nw[a]);
    }
    const float mx = fmaxf(av[0], fmaxf(av[1], av[2]));                                      /* CubedNormalize, normal.glsl:39-43 */
    for (int a = 0; a < 3; a++) normal_world[a] = signf(nw[a]) * (av[a] >= mx ? 1.0f : 0.0f);
    o_cache_key_of(b->center, normal_world, &b->key);                                        /* :45-47 */
    b->face = b->key.direction;
    b->avg_albedo = leaf_avg_albedo(in, ih->leaf_ptr);
}

/* final_gather.rchit:68-80 / surfel.rchit:60-71: incoming radiance at the block -> outgoing radiance towards the ray */
static void modulate_by_block_albedo(uint32_t packed_albedo, const float radiance[3], float out[3])
{
    float albedo[3] = { (float)((packed_albedo >> 22) & 1023u) / 1023.0f, (float)((packed_albedo >> 12) & 1023u) / 1023.0f,
                        (float)((packed_albedo >> 2) & 1023u) / 1023.0f };
    float srgb[3];
    o_acescg2srgb(radiance, srgb);
    for (int a = 0; a < 3; a++) srgb[a] = srgb[a] * o_srgb_to_linear(albedo[a]);
    o_srgb2acescg(srgb, out);
}

/* ---- the passes.  Each phase that traces is a parallel loop over independent items writing per-item records; everything that
 * has an ORDER (pool writes, cache inserts) is applied afterwards in ascending index ---- */

typedef struct GiFrame {
    const OScene *s; const OCamera *cam; const OSky *sky; const OTraceConfig *cfg; OGiState *st;
    uint16_t *img; const OGBuffer *gb;
    float sun_dir[3], sun_nee[3];           /* normalize(sunDir); sun_radiance(normalize(sunDir)) * (1 - cos(solar radius)) */
    uint32_t key;
    /* per-pixel / per-surfel records */
    OSurfel *sched;                         /* surfel a pixel (or surfel invocation) wants to write; direction 0xFFFFFFFF: none */
    uint8_t *ins_valid; OSpatialHashKey *ins_key; float *ins_val;       /* surfel pass: the insert of invocation i */
    OGiStats *stats;
} GiFrame;

typedef struct GiJob { GiFrame *F; void (*fn)(GiFrame *, uint32_t, OStats *, OGiStats *); uint32_t n, tid, nthreads; OGiStats gs; } GiJob;

static void *gi_worker(void *arg)
{
    GiJob *j = (GiJob *)arg;
    OStats st;
    memset(&st, 0, sizeof st);
    for (uint32_t base = j->tid * 64u; base < j->n; base += j->nthreads * 64u)
        for (uint32_t i = base; i < j->n && i < base + 64u; i++) j->fn(j->F, i, &st, &j->gs);
    return NULL;
}

static void gi_parallel(GiFrame *F, uint32_t n, void (*fn)(GiFrame *, uint32_t, OStats *, OGiStats *), uint32_t n_threads, int leaf_mode)
{
    g_rough_blocks = leaf_mode;
    GiJob *jobs = (GiJob *)calloc(n_threads, sizeof(GiJob));
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    for (uint32_t i = 0; i < n_threads; i++) { jobs[i].F = F; jobs[i].fn = fn; jobs[i].n = n; jobs[i].tid = i; jobs[i].nthreads = n_threads; }
    for (uint32_t i = 1; i < n_threads; i++) pthread_create(&th[i], NULL, gi_worker, &jobs[i]);
    gi_worker(&jobs[0]);
    for (uint32_t i = 1; i < n_threads; i++) pthread_join(th[i], NULL);
    for (uint32_t i = 0; i < n_threads; i++) {
        OGiStats *a = F->stats, *b = &jobs[i].gs;
        a->near_hits += b->near_hits; a->far_hits += b->far_hits; a->far_misses += b->far_misses; a->cache_found += b->cache_found;
        a->surfel_rays += b->surfel_rays;
    }
    free(th); free(jobs);
    g_rough_blocks = 0;
}

/* what ambient_occlusion.rgen:15-30 and final_gather.rgen:15-34 both derive from the G-buffer of a pixel */
static int pixel_setup(const GiFrame *F, uint32_t pi, float n[3], float x[3], float dir[3], uint32_t *s1_out)
{
    const OTraceConfig *cfg = F->cfg;
    const float hit_t = F->gb->depth[pi];
    if (hit_t == INFINITY) return 0;                                                         /* :16-19 */
    o_gi_unpack_normal(F->gb->normal[pi], n);                                                 /* :22 */
    float o[3], d[3];
    o_camera_ray(F->cam, cfg->width, cfg->height, pi % cfg->width, pi / cfg->width, 0.5f, 0.5f, o, d);
    for (int a = 0; a < 3; a++) x[a] = (hit_t * d[a] + o[a]) + n[a] * 0.01f;                 /* :23 */
    const uint32_t s1 = o_pcg(o_pcg(pi ^ F->key) + 0u);                                       /* the path tracer's stream of sample 0 */
    float local[3], rot[3];
    cosine_sample(s1, local);                                                                 /* :28 (blue noise -> pcg) */
    o_rotate_by_normal(n, local, rot);                                                        /* :30 */
    normalize3(rot, dir);                                                                     /* :62 / final_gather.rgen:48 */
    *s1_out = s1;
    return 1;
}

/* primary pass leftovers: hit.rchit:57 writes illuminance 0, miss.rmiss:8-13 writes the denoised image */
static void pass_primary_fill(GiFrame *F, uint32_t pi, OStats *st, OGiStats *gs)
{
    (void)st; (void)gs;
    const OTraceConfig *cfg = F->cfg;
    if (F->gb->depth[pi] != INFINITY) { const float z[3] = { 0.0f, 0.0f, 0.0f }; store_illuminance(F->img, pi, z, 0.0f); return; }
    float o[3], d[3], dn[3], sk[3], su[3], v[3];
    o_camera_ray(F->cam, cfg->width, cfg->height, pi % cfg->width, pi / cfg->width, 0.5f, 0.5f, o, d);
    normalize3(d, dn);
    o_sky_radiance(F->sky, dn, sk);
    o_sun_radiance(F->sky, dn, su);
    for (int a = 0; a < 3; a++) v[a] = (sk[a] + su[a]) / 3.14f;
    store_illuminance(F->img, pi, v, 100000.0f);
}

/* ambient_occlusion.rgen:14-66 + nee.rmiss + ambient_occlusion.rchit / .rmiss */
static void pass_ambient_occlusion(GiFrame *F, uint32_t pi, OStats *st, OGiStats *gs)
{
    float n[3], x[3], dir[3], L[3];
    uint32_t s1;
    if (!pixel_setup(F, pi, n, x, dir, &s1)) return;
    load_illuminance(F->img, pi, L);                                                          /* :25-26, :32 */
    if (dot3(F->sky->direction, n) > 0.0f) {                                                  /* :34-49 */
        if (!trace_any(F->s, x, F->sun_dir, 0.1f, 10000.0f, st)) {
            const float ndl = dot3(n, F->sun_dir);                                            /* nee.rmiss:21 */
            for (int a = 0; a < 3; a++) L[a] += F->sun_nee[a] * ndl;
        }
    }
    SceneHit h;
    trace_closest(F->s, x, dir, 0.1f, 8.0f, &h, st);                                          /* :53-65 */
    if (h.hit) gs->near_hits++;
    store_illuminance(F->img, pi, L, h.hit ? h.ih.t : 0.0f);                                  /* ambient_occlusion.rchit:11-12 / .rmiss:11-12 */
}

/* final_gather.rgen:14-52 + final_gather.rchit / .rmiss */
static void pass_final_gather(GiFrame *F, uint32_t pi, OStats *st, OGiStats *gs)
{
    F->sched[pi].direction = 0xFFFFFFFFu;
    float n[3], x[3], dir[3], L[3];
    uint32_t s1;
    if (!pixel_setup(F, pi, n, x, dir, &s1)) return;
    if (load_illuminance(F->img, pi, L) > 0.0f) return;                                       /* :21-26: resolved by the ambient-occlusion pass */
    SceneHit h;
    trace_closest(F->s, x, dir, 8.0f, F->cam->far, &h, st);                                   /* :39-51 */
    if (!h.hit) {                                                                             /* final_gather.rmiss:13-21 */
        float sk[3], dn[3];
        normalize3(dir, dn);
        o_sky_radiance(F->sky, dn, sk);
        for (int a = 0; a < 3; a++) L[a] += sk[a];
        store_illuminance(F->img, pi, L, 0.0f);
        gs->far_misses++;
        return;
    }
    gs->far_hits++;
    BlockHit b;
    block_of_hit(F->s, &h, &b);
    float radiance[3], out[3];
    uint32_t count;
    const int found = o_cache_get(F->st->table, F->st->capacity, &b.key, F->cfg->frame_index, radiance, &count);   /* rchit:49-51 */
    if (found) gs->cache_found++;
    const float prob = 1.0f / (float)(count + 2u);                                            /* :52 */
    if (u01(o_pcg(s1 + 6u)) > prob) {                                                         /* :53-63 (blue noise -> pcg) */
        memcpy(F->sched[pi].position, b.center, sizeof b.center);
        F->sched[pi].direction = b.face;
    }
    modulate_by_block_albedo(b.avg_albedo, radiance, out);                                    /* :68-80 */
    for (int a = 0; a < 3; a++) L[a] += out[a];                                               /* :82-85 */
    store_illuminance(F->img, pi, L, h.ih.t);                                                 /* :86-89 */
}

/* surfel.rgen:12-67 + surfel/nee.rmiss + surfel.rchit / .rmiss: traces and look-ups of invocation i; its insert / pool write are recorded */
static void pass_surfel(GiFrame *F, uint32_t i, OStats *st, OGiStats *gs)
{
    F->ins_valid[i] = 0;
    F->sched[i].direction = 0xFFFFFFFFu;
    const OSurfel e = F->st->pool[i];
    if (e.direction >= 6u) return;                                                            /* :14-16 */
    gs->surfel_rays++;
    float n[3], x[3], local[3], rot[3], dir[3], payload[3] = { 0.0f, 0.0f, 0.0f };
    o_face_id_to_normal(e.direction, n);                                                      /* :18 */
    for (int a = 0; a < 3; a++) x[a] = e.position[a] + 2.01f * n[a];                          /* :27 */
    const uint32_t s1 = o_pcg(o_pcg((i + 0x51F15EEDu) ^ F->key));
    cosine_sample(s1, local);                                                                 /* :30 */
    o_rotate_by_normal(n, local, rot);                                                        /* :32 */
    normalize3(rot, dir);                                                                     /* :63 */
    if (dot3(F->sky->direction, n) > 0.0f) {                                                  /* :38-52 */
        if (!trace_any(F->s, x, F->sun_dir, 0.1f, 10000.0f, st)) {
            const float ndl = dot3(n, F->sun_dir);                                            /* surfel/nee.rmiss:22-26 */
            for (int a = 0; a < 3; a++) payload[a] += F->sun_nee[a] * ndl;
        }
    }
    OSpatialHashKey own;
    for (int a = 0; a < 3; a++) own.position[a] = (int32_t)(e.position[a] / 4.0f);            /* surfel.rchit:77-79 / .rmiss:19-21 */
    own.direction = e.direction;
    SceneHit h;
    trace_closest(F->s, x, dir, 0.1f, 10000.0f, &h, st);                                      /* :54-66 */
    float value[3];
    if (!h.hit) {                                                                             /* surfel.rmiss:15-23 */
        float sk[3], dn[3];
        normalize3(dir, dn);
        o_sky_radiance(F->sky, dn, sk);
        for (int a = 0; a < 3; a++) value[a] = sk[a] + payload[a];
    } else {
        BlockHit b;
        block_of_hit(F->s, &h, &b);
        float radiance[3], out[3];
        uint32_t count;
        if (!o_cache_get(F->st->table, F->st->capacity, &b.key, F->cfg->frame_index, radiance, &count)) {       /* surfel.rchit:49-52 */
            const float prob = 1.0f / (float)(count + 2u);                                    /* :91 */
            if (u01(o_pcg(s1 + 6u)) > prob) {                                                 /* :58, :92-100 */
                memcpy(F->sched[i].position, b.center, sizeof b.center);
                F->sched[i].direction = b.face;
            }
            return;
        }
        gs->cache_found++;
        modulate_by_block_albedo(b.avg_albedo, radiance, out);                                /* :60-71 */
        for (int a = 0; a < 3; a++) value[a] = out[a] + payload[a];                           /* :83 */
    }
    F->ins_valid[i] = 1;
    F->ins_key[i] = own;
    memcpy(F->ins_val + 3 * (size_t)i, value, sizeof value);
}

void o_gi_frame(const OScene *s, const OCamera *cam, const OSky *sky, const OTraceConfig *cfg, OGiState *state,
                uint16_t *illuminance, const OGBuffer *gb, OGiStats *stats, uint32_t n_threads)
{
    if (n_threads < 1) n_threads = 1;
    const uint32_t npx = cfg->width * cfg->height, S = state->pool_size;
    OGiStats local;
    if (!stats) stats = &local;
    memset(stats, 0, sizeof *stats);
    /* primary pass: the G-buffer of the sample-0 primary hit (hit.rchit / miss.rmiss), written by the path tracer's own code */
    {
        OTraceConfig pc = *cfg;
        pc.spp = 1; pc.max_bounces = 0; pc.y0 = 0; pc.y1 = cfg->height; pc.flags = 0;
        float *scratch = (float *)malloc(sizeof(float) * 4 * (size_t)npx);
        o_render_g(s, cam, sky, &pc, NULL, 0, scratch, NULL, NULL, NULL, NULL, gb, NULL, n_threads);
        free(scratch);
    }
    GiFrame F;
    memset(&F, 0, sizeof F);
    F.s = s; F.cam = cam; F.sky = sky; F.cfg = cfg; F.st = state; F.img = illuminance; F.gb = gb; F.stats = stats;
    F.key = o_pcg(cfg->seed + 0x9E3779B9u * cfg->frame_index);
    normalize3(sky->direction, F.sun_dir);                              /* ambient_occlusion.rgen:45 */
    {
        float dn[3], su[3];
        normalize3(F.sun_dir, dn);                                      /* nee.rmiss:18: normalize(gl_WorldRayDirectionEXT) */
        o_sun_radiance(sky, dn, su);
        const float cone = 1.0f - cosf(sky->sunlight_intensity[3]);
        for (int a = 0; a < 3; a++) F.sun_nee[a] = su[a] * cone;
    }
    const uint32_t nrec = npx > S ? npx : S;
    F.sched = (OSurfel *)malloc(sizeof(OSurfel) * (nrec ? nrec : 1));
    F.ins_valid = (uint8_t *)calloc(S ? S : 1, 1);
    F.ins_key = (OSpatialHashKey *)malloc(sizeof(OSpatialHashKey) * (S ? S : 1));
    F.ins_val = (float *)malloc(sizeof(float) * 3 * (S ? S : 1));

    gi_parallel(&F, npx, pass_primary_fill, n_threads, 0);
    gi_parallel(&F, npx, pass_ambient_occlusion, n_threads, 2);         /* ambient_occlusion.rint for both of its rays */
    gi_parallel(&F, npx, pass_final_gather, n_threads, 1);              /* rough.rint */
    /* pool writes of the final-gather pass, ascending pixel index: surfel_pool[pixel % SurfelPoolSize] (final_gather.rchit:56-62) */
    for (uint32_t pi = 0; pi < npx && S; pi++)
        if (F.sched[pi].direction != 0xFFFFFFFFu) { state->pool[pi % S] = F.sched[pi]; stats->surfels_scheduled++; }
    gi_parallel(&F, S, pass_surfel, n_threads, 1);                      /* rough.rint for both of its rays */
    for (uint32_t i = 0; i < S; i++) {                                  /* ascending invocation index */
        if (F.sched[i].direction != 0xFFFFFFFFu) { state->pool[i] = F.sched[i]; stats->surfels_scheduled++; }   /* surfel.rchit:93-99: index = i % S = i */
        if (F.ins_valid[i]) {
            o_cache_insert(state->table, state->capacity, &F.ins_key[i], F.ins_val + 3 * (size_t)i, cfg->frame_index);
            stats->cache_inserts++;
        }
    }
    free(F.sched); free(F.ins_valid); free(F.ins_key); free(F.ins_val);
}
