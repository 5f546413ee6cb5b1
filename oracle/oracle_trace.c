/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_trace.h for scope, citations and parity status).
 * Build with -O2 -ffp-contract=off (no FMA contraction, no fast-math): every float expression
 * below is evaluated in the written order so that hit ids / t are the bit-exact reference for the
 * CUDA path (which is built with -fmad=false).  `fmaf()` appears only where the traversal spec
 * explicitly asks for a fused multiply-add.
 */
#include "oracle_trace.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct OInstance {
    float obj2world[12];
    float prev_obj2world[12];      /* last frame (hit.rchit:92) */
    float w2o[12];
    const OTree *tree;
    const uint8_t *materials;
    const uint8_t *palette;        /* RGBA8 * 255 */
    const uint32_t *leaf_order;    /* pool pointer -> index in Tree::iter_leaf order (= gl_PrimitiveID) */
    /* occupied bounds of the tree in object voxels: [bmin, bmax) spans every leaf (multiples of 4).
     * The cell walk is clipped to this box instead of the full [0, extent)^3. */
    int32_t bmin[3], bmax[3];
} OInstance;

struct OScene {
    uint32_t n;
    OInstance *inst;
    uint32_t n_orders;
    uint32_t **orders;
    const OTree **order_trees;
    /* optional early-out: padded world box of every instance's [0,E]^3 object box.  Purely a
     * speed-up of the linear instance loop (o_scene_set_cull; tests check on == off). */
    int cull;
    float *boxes;     /* n * 6: lo xyz, hi xyz */
};

/* ---------------------------------------------------------------- helpers */

static float dot3(const float a[3], const float b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

/* GLSL normalize(): v / sqrt(dot(v,v)) with IEEE sqrt and divide */
static void normalize3(const float v[3], float out[3])
{
    float l = sqrtf(dot3(v, v));
    out[0] = v[0] / l; out[1] = v[1] / l; out[2] = v[2] / l;
}

static float signf(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }

/* assets/shaders/headers/spatial_hash.glsl:105-111 */
uint32_t o_pcg(uint32_t v)
{
    uint32_t state = v * 747796405u + 2891336453u;
    uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}

static float u01(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }

/* cos/sin of 2*pi*u for u in [0,1) from plain mul/add polynomials (bit-reproducible on any IEEE
 * machine).  Quadrant q = floor(4u); theta = (4u - q - 0.5) * pi/2 in [-pi/4, pi/4); Taylor to
 * theta^9 / theta^10; then rotate by pi/4 and by the quadrant. */
void o_cossin_2pi(float u, float *c, float *s)
{
    float x = u * 4.0f;
    float qf = floorf(x);
    int q = (int)qf & 3;
    float th = ((x - qf) - 0.5f) * 1.57079632679489661923f;
    float t2 = th * th;
    float sn = th * (1.0f + t2 * (-1.0f / 6.0f + t2 * (1.0f / 120.0f + t2 * (-1.0f / 5040.0f + t2 * (1.0f / 362880.0f)))));
    float cs = 1.0f + t2 * (-0.5f + t2 * (1.0f / 24.0f + t2 * (-1.0f / 720.0f + t2 * (1.0f / 40320.0f + t2 * (-1.0f / 3628800.0f)))));
    float c45 = (cs - sn) * 0.70710678118654752440f;   /* cos(theta + pi/4) */
    float s45 = (cs + sn) * 0.70710678118654752440f;   /* sin(theta + pi/4) */
    switch (q) {
    case 0: *c = c45;  *s = s45;  break;
    case 1: *c = -s45; *s = c45;  break;
    case 2: *c = -c45; *s = -s45; break;
    default: *c = s45; *s = -c45; break;
    }
}

/* assets/shaders/headers/normal.glsl:31-37 */
void o_rotate_by_normal(const float n[3], const float v[3], float out[3])
{
    float q[4] = { -n[1], n[0], 0.0f, 1.0f + n[2] };
    float l = sqrtf(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
    q[0] = q[0] / l; q[1] = q[1] / l; q[2] = q[2] / l; q[3] = q[3] / l;
    if (n[2] < -0.99999f) { q[0] = -1.0f; q[1] = 0.0f; q[2] = 0.0f; q[3] = 0.0f; }
    float d = dot3(q, v);
    float qq = dot3(q, q);
    float cr[3] = { q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0] };
    for (int i = 0; i < 3; i++)
        out[i] = ((2.0f * d) * q[i] + (q[3] * q[3] - qq) * v[i]) + (2.0f * q[3]) * cr[i];
}

/* inverse of a row-major 3x4 affine, evaluated in double then rounded (the Vulkan driver's
 * gl_WorldToObjectEXT; arithmetic unpinned — SURVEY.md §8c(2)) */
void o_world2obj(const float m[12], float w[12])
{
    double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    double A = e * i - f * h, B = c * h - b * i, C = b * f - c * e;
    double D = f * g - d * i, E = a * i - c * g, F = c * d - a * f;
    double G = d * h - e * g, H = b * g - a * h, I = a * e - b * d;
    double det = a * A + b * D + c * G;
    double inv[9] = { A / det, B / det, C / det, D / det, E / det, F / det, G / det, H / det, I / det };
    double t[3] = { m[3], m[7], m[11] };
    for (int r = 0; r < 3; r++) {
        w[4 * r + 0] = (float)inv[3 * r + 0];
        w[4 * r + 1] = (float)inv[3 * r + 1];
        w[4 * r + 2] = (float)inv[3 * r + 2];
        w[4 * r + 3] = (float)(-(inv[3 * r + 0] * t[0] + inv[3 * r + 1] * t[1] + inv[3 * r + 2] * t[2]));
    }
}

/* ------------------------------------------------------------------ scene */

OScene *o_scene_create(uint32_t n, const float *obj2world, const OTree *const *trees,
                       const uint8_t *const *materials, const uint8_t *const *palettes)
{
    OScene *s = (OScene *)calloc(1, sizeof *s);
    s->n = n;
    s->inst = (OInstance *)calloc(n ? n : 1, sizeof(OInstance));
    s->orders = (uint32_t **)calloc(n ? n : 1, sizeof(uint32_t *));
    s->order_trees = (const OTree **)calloc(n ? n : 1, sizeof(OTree *));
    for (uint32_t k = 0; k < n; k++) {
        OInstance *in = &s->inst[k];
        memcpy(in->obj2world, obj2world + 12 * (size_t)k, sizeof in->obj2world);
        memcpy(in->prev_obj2world, in->obj2world, sizeof in->prev_obj2world);
        o_world2obj(in->obj2world, in->w2o);
        in->tree = trees[k]; in->materials = materials[k]; in->palette = palettes[k];
        if (in->tree->fanout_log2[0] != 2) { o_scene_destroy(s); return NULL; }   /* 64-bit leaf masks only */
        uint32_t j;
        for (j = 0; j < s->n_orders; j++) if (s->order_trees[j] == trees[k]) break;
        if (j == s->n_orders) {
            const OTree *t = trees[k];
            uint64_t nl = o_tree_num_leaves(t);
            uint32_t top = t->n_levels > 1 ? t->pool[0].top : 1;
            uint32_t *ptrs = (uint32_t *)malloc(sizeof(uint32_t) * (nl ? nl : 1));
            uint32_t *order = (uint32_t *)malloc(sizeof(uint32_t) * (top ? top : 1));
            memset(order, 0xFF, sizeof(uint32_t) * (top ? top : 1));
            o_tree_iter_leaf(t, NULL, ptrs, nl);
            for (uint64_t l = 0; l < nl; l++) order[ptrs[l]] = (uint32_t)l;
            free(ptrs);
            s->orders[j] = order; s->order_trees[j] = t; s->n_orders++;
        }
        in->leaf_order = s->orders[j];
    }
    s->cull = 1;
    s->boxes = (float *)calloc(n ? n : 1, 6 * sizeof(float));
    for (uint32_t k = 0; k < n; k++) {
        OInstance *in = &s->inst[k];
        /* object-space bounds of the occupied leaves (a leaf spans 4 voxels) */
        uint64_t nl = o_tree_num_leaves(in->tree);
        uint32_t *pos = (uint32_t *)malloc(sizeof(uint32_t) * 3 * (nl ? nl : 1));
        o_tree_iter_leaf(in->tree, pos, NULL, nl);
        double omin[3] = { 0, 0, 0 }, omax[3] = { 0, 0, 0 };
        for (uint64_t l = 0; l < nl; l++)
            for (int a = 0; a < 3; a++) {
                double v = (double)pos[3 * l + a];
                if (l == 0 || v < omin[a]) omin[a] = v;
                if (l == 0 || v + 4.0 > omax[a]) omax[a] = v + 4.0;
            }
        free(pos);
        for (int a = 0; a < 3; a++) { in->bmin[a] = (int32_t)omin[a]; in->bmax[a] = (int32_t)omax[a]; }   /* empty tree: [0,0) */
        double lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY }, big = 0.0;
        for (int c = 0; c < 8; c++) {
            double p[3] = { (c & 1) ? omax[0] : omin[0], (c & 2) ? omax[1] : omin[1], (c & 4) ? omax[2] : omin[2] };
            for (int a = 0; a < 3; a++) {
                const float *M = in->obj2world + 4 * a;
                double w = (double)M[0] * p[0] + (double)M[1] * p[1] + (double)M[2] * p[2] + (double)M[3];
                if (w < lo[a]) lo[a] = w;
                if (w > hi[a]) hi[a] = w;
                if (fabs(w) > big) big = fabs(w);
            }
        }
        double pad = 0.05 + 1e-4 * big;
        for (int a = 0; a < 3; a++) { s->boxes[6 * k + a] = (float)(lo[a] - pad); s->boxes[6 * k + 3 + a] = (float)(hi[a] + pad); }
    }
    return s;
}

void o_scene_set_cull(OScene *s, int enabled) { s->cull = enabled; }

/* Block-granular leaves (assets/shaders/final_gather/rough.rint:46-58, the reference's far-field intersection shader): a
 * non-empty leaf is hit where the ray enters its 4x4x4 box, no DDA.  Process-wide switch, set around a call by the tests. */
static int g_rough_blocks;      /* leaf mode: 0 = hit.rint (precise DDA), 1 = rough.rint (solid block), 2 = ambient_occlusion.rint */
void o_set_rough_blocks(int enabled) { g_rough_blocks = enabled; }

void o_scene_set_previous(OScene *s, const float *obj2world_prev)
{
    for (uint32_t k = 0; k < s->n; k++) memcpy(s->inst[k].prev_obj2world, obj2world_prev + 12 * (size_t)k, 12 * sizeof(float));
}

/* 1/d for the box tests only; a zero component becomes a huge FINITE value so that (b - o) * inv is
 * never 0 * inf */
static void cull_inv(const float d[3], float inv[3])
{
    for (int a = 0; a < 3; a++) inv[a] = fabsf(d[a]) < 1e-30f ? copysignf(1e30f, d[a]) : 1.0f / d[a];
}

/* conservative: returns 0 only when the ray certainly misses the padded box within [tmin, tmax] */
static int box_may_hit(const float *b, const float o[3], const float inv[3], float tmin, float tmax)
{
    float t0 = tmin, t1 = tmax;
    for (int a = 0; a < 3; a++) {
        float ta = (b[a] - o[a]) * inv[a], tb = (b[3 + a] - o[a]) * inv[a];
        t0 = fmaxf(t0, fminf(ta, tb));     /* inv[] is finite (cull_inv), so no 0 * inf = NaN here */
        t1 = fminf(t1, fmaxf(ta, tb));
    }
    return t0 <= t1;
}

void o_scene_destroy(OScene *s)
{
    if (!s) return;
    for (uint32_t j = 0; j < s->n_orders; j++) free(s->orders[j]);
    free(s->orders); free((void *)s->order_trees); free(s->inst); free(s->boxes); free(s);
}

/* --------------------------------------------------------------- LEAF STEP */

/* assets/shaders/primary/hit.rint:20-28 */
static void intersect_aabb(const float origin[3], const float dir[3], float box_min, float box_max, float out[2])
{
    float t1[3], t2[3];
    for (int a = 0; a < 3; a++) {
        float tMin = (box_min - origin[a]) / dir[a];
        float tMax = (box_max - origin[a]) / dir[a];
        t1[a] = fminf(tMin, tMax);
        t2[a] = fmaxf(tMin, tMax);
    }
    out[0] = fmaxf(fmaxf(t1[0], t1[1]), t1[2]);
    out[1] = fminf(fminf(t2[0], t2[1]), t2[2]);
}

/* assets/shaders/primary/hit.rint:43-131 with scale = 1.0.  `origin` is leaf-local
 * (gl_ObjectRayOriginEXT - block.position, :139-141).  Returns 1 and (t, voxelId) when the DDA
 * reaches a set bit; acceptance against [tmin, tmax] is the caller's (reportIntersectionEXT). */
static int leaf_dda_mode(const float origin[3], const float dir[3], uint64_t grid, float ray_tmin, float *t_out, uint32_t *voxel_out, int ao)
{
    float tb[2];
    intersect_aabb(origin, dir, 0.0f, 4.0f, tb);                       /* :52 */
    if (tb[0] >= tb[1]) return 0;                                       /* :53-56 */
    if (tb[1] <= 0.0f) return 0;                                        /* :57-59 */
    if (ao && tb[0] <= 8.0f && 8.0f <= tb[1]) {
        /* ambient_occlusion.rint:62-73 — AMBIENT_OCCLUSION_THRESHOLD lies inside this leaf's box: a non-empty leaf is reported as
         * a solid block at its entry (voxelId 0xFF, hit kind 1), an empty one is skipped; no DDA either way */
        if (grid == 0) return 0;
        *t_out = tb[0] + 0.0f;
        *voxel_out = 0xFFu;
        return 1;
    }
    float hitDistance = fmaxf(tb[0], ray_tmin);                         /* :71 */
    int pos[3];
    float step[3], t_coef[3], t_bias[3], tMax[3], tDelta[3];
    for (int a = 0; a < 3; a++) {
        float p = origin[a] + dir[a] * hitDistance;                     /* :72 */
        int ip = (int)floorf(p);
        pos[a] = ip < 0 ? 0 : (ip > 3 ? 3 : ip);                        /* :74 */
        step[a] = signf(dir[a]);                                        /* :79 */
        t_coef[a] = 1.0f / dir[a];                                      /* :87 */
        t_bias[a] = t_coef[a] * origin[a];                              /* :88 */
        tMax[a] = ((float)pos[a] + fmaxf(step[a], 0.0f)) * t_coef[a] - t_bias[a];   /* :90 */
        tDelta[a] = (1.0f * t_coef[a]) * step[a];                       /* :96 */
    }
    uint32_t hit = ((uint32_t)pos[0] << 4) | ((uint32_t)pos[1] << 2) | (uint32_t)pos[2];   /* :99 */
    while (((grid >> hit) & 1u) == 0) {                                 /* :101 */
        float comp[3];
        comp[0] = (tMax[2] >= tMax[0] ? 1.0f : 0.0f) * (tMax[1] >= tMax[0] ? 1.0f : 0.0f);   /* :102 */
        comp[1] = (tMax[0] >= tMax[1] ? 1.0f : 0.0f) * (tMax[2] >= tMax[1] ? 1.0f : 0.0f);
        comp[2] = (tMax[1] >= tMax[2] ? 1.0f : 0.0f) * (tMax[0] >= tMax[2] ? 1.0f : 0.0f);
        for (int a = 0; a < 3; a++) pos[a] += (int)(step[a] * comp[a]);                      /* :103-104 */
        hitDistance = fminf(fminf(tMax[0], tMax[1]), tMax[2]);          /* :106 */
        if (hitDistance + 0.001f >= tb[1]) return 0;                    /* :107-109 */
        for (int a = 0; a < 3; a++) tMax[a] += tDelta[a] * comp[a];     /* :110 */
        /* deviation (2): a tie-step may leave 0..3 one iteration before :107 fires */
        if ((pos[0] | pos[1] | pos[2]) & ~3) return 0;
        hit = ((uint32_t)pos[0] << 4) | ((uint32_t)pos[1] << 2) | (uint32_t)pos[2];          /* :112 */
    }
    /* :115.  "+ 0.0f": a ray that starts inside a set voxel exactly on a leaf plane reports max(-0, +0), whose
     * sign IEEE leaves open (glibc's fmaxf and PTX max.f32 disagree); -0 + +0 = +0 makes t canonical. */
    *t_out = hitDistance / 1.0f + 0.0f;
    *voxel_out = hit;
    return 1;
}

int o_leaf_dda(const float origin[3], const float dir[3], uint64_t grid, float ray_tmin, float *t_out, uint32_t *voxel_out)
{
    return leaf_dda_mode(origin, dir, grid, ray_tmin, t_out, voxel_out, 0);
}

/* assets/shaders/final_gather/ambient_occlusion.rint:46-134: hit.rint's DDA plus the threshold rule above (the intersection
 * shader of the ambient-occlusion pass' two rays) */
int o_leaf_ao(const float origin[3], const float dir[3], uint64_t grid, float ray_tmin, float *t_out, uint32_t *voxel_out)
{
    return leaf_dda_mode(origin, dir, grid, ray_tmin, t_out, voxel_out, 1);
}

/* ------------------------------------------------------- instance traversal */

typedef struct InstHit {
    float t;
    uint32_t leaf_ptr, voxel;
    int32_t leaf_min[3];
    float o[3], d[3];             /* object-space ray */
} InstHit;

#define O_DIR_EPS 9.094947017729282e-13f   /* 2^-40 */

/*
 * HDDA over the InternalNode levels (DESIGN.md "Traversal semantics"; defined here, the
 * reference has none).  State = integer voxel coordinate `iv` of the cell being examined.
 * Per iteration: descend from the root with the InternalNode::get index arithmetic
 * (crates/vdb/src/node/internal.rs:77-95) until a clear child_mask bit (empty cell of edge L),
 * an empty 4x4x4 block of children of a 16^3 node (empty cell of edge 4L), or a leaf (L = 4,
 * LEAF STEP); then leave the cell through its exit planes
 * t_a = (plane_a - o_a) * (1/d_a)  (the difference first: plane*inv - o*inv would cancel
 * catastrophically, |o/d| >> t); axes whose t_a equals the minimum step one cell
 * (ties step together, like hit.rint:102-104), the others are re-derived from the ray at that t
 * and clamped into the cell's own range.
 */
static int trace_instance(const OInstance *in, const float ow[3], const float dw[3], float tmin, float tmax,
                          InstHit *out, OStats *st)
{
    const OTree *t = in->tree;
    const float *M = in->w2o;
    float o[3], d[3], inv[3];
    for (int a = 0; a < 3; a++) {
        /* gl_ObjectRayOriginEXT / gl_ObjectRayDirectionEXT (hit.rint:140-141) */
        o[a] = ((M[4 * a] * ow[0] + M[4 * a + 1] * ow[1]) + M[4 * a + 2] * ow[2]) + M[4 * a + 3];
        d[a] = (M[4 * a] * dw[0] + M[4 * a + 1] * dw[1]) + M[4 * a + 2] * dw[2];
        if (fabsf(d[a]) < O_DIR_EPS) d[a] = copysignf(O_DIR_EPS, d[a]);        /* deviation (1) */
        inv[a] = 1.0f / d[a];
    }
    uint32_t root_level = t->n_levels - 1;
    /* clip against the tree's occupied bounds [bmin, bmax) rather than the whole [0, extent)^3:
     * everything outside holds no leaf, so the walk may start at, and stop on leaving, that box */
    const int32_t *bmin = in->bmin, *bmax = in->bmax;
    float lo[3], hi[3];
    for (int a = 0; a < 3; a++) {
        float t0 = ((float)bmin[a] - o[a]) * inv[a];
        float t1 = ((float)bmax[a] - o[a]) * inv[a];
        lo[a] = fminf(t0, t1); hi[a] = fmaxf(t0, t1);
    }
    float t_enter = fmaxf(fmaxf(lo[0], lo[1]), fmaxf(lo[2], tmin));
    /* The cell walk stops a hair AFTER tmax (2^-18 relative + 2^-20): plane times and leaf
     * hit times (hit.rint's divisions) round differently, so a hit at t == tmax can sit in a cell the
     * walk enters at tmax + 1 ulp.  Acceptance itself stays exact (tmin <= t <= tmax); the pad only
     * makes the result independent of how far tmax was already shrunk by other instances. */
    float tmax_pad = tmax + (fabsf(tmax) * 3.814697265625e-06f + 9.5367431640625e-07f);
    float t_box = fminf(fminf(hi[0], hi[1]), hi[2]);
    float t_exit = fminf(t_box, tmax_pad);
    if (!(t_enter <= t_exit) || bmax[0] <= bmin[0]) return 0;
    int32_t iv[3];
    for (int a = 0; a < 3; a++) {
        int32_t ip = (int32_t)floorf(fmaf(d[a], t_enter, o[a]));
        iv[a] = ip < bmin[a] ? bmin[a] : (ip > bmax[a] - 1 ? bmax[a] - 1 : ip);
    }
    for (;;) {
        const uint8_t *node = t->root;
        uint32_t lv = root_level, ptr = 0;
        int32_t L;
        for (;;) {
            if (lv == 0) { L = 4; break; }
            uint32_t f = t->fanout_log2[lv], c = t->extent_log2[lv - 1], m = (1u << f) - 1u;
            uint32_t ccx = ((uint32_t)iv[0] >> c) & m, ccy = ((uint32_t)iv[1] >> c) & m, ccz = ((uint32_t)iv[2] >> c) & m;
            if (f == 4) {
                /* BLOCK level (defined here, DESIGN.md "Traversal semantics"): a 16^3 node is first asked per
                 * 4x4x4 block of children; a block without any set child_mask bit is one empty cell of 4x the
                 * child edge.  Child bit = x<<8 | y<<4 | z  =>  word x*4 + (y>>2), bit (y&3)<<4 | z. */
                st->mask_words++;
                const uint64_t *words = (const uint64_t *)node;
                uint64_t sel = 0x000F000F000F000Full << (4u * (ccz >> 2)), any = 0;
                for (uint32_t x = ccx & ~3u; x < (ccx & ~3u) + 4u; x++) any |= words[x * 4u + (ccy >> 2)] & sel;
                if (!any) { L = (int32_t)4 << c; break; }
            }
            uint32_t idx = (ccx << (2 * f)) | (ccy << f) | ccz;
            st->mask_words++;
            if (!o_bitmask_get((const uint64_t *)node, idx)) { L = (int32_t)1 << c; break; }
            st->child_ptrs++;
            ptr = ((const uint32_t *)(node + 8u * t->mask_words[lv]))[idx];
            node = o_pool_get(&t->pool[lv - 1], ptr);
            lv--;
        }
        int32_t cmin[3] = { iv[0] & ~(L - 1), iv[1] & ~(L - 1), iv[2] & ~(L - 1) };
        if (lv == 0) {
            uint64_t grid;
            memcpy(&grid, node, 8);
            st->mask_words++;
            float lo_[3] = { o[0] - (float)cmin[0], o[1] - (float)cmin[1], o[2] - (float)cmin[2] };   /* hit.rint:140 */
            float th; uint32_t vox;
            int reported;
            if (g_rough_blocks == 2) reported = o_leaf_ao(lo_, d, grid, tmin, &th, &vox);
            else if (g_rough_blocks) {
                float tb[2];
                intersect_aabb(lo_, d, 0.0f, 4.0f, tb);                 /* rough.rint:46 */
                reported = !(tb[0] >= tb[1]) && grid != 0;              /* :47-53 */
                th = tb[0] + 0.0f; vox = 0;                             /* :54 reportIntersectionEXT(initialIntersectionT.x, 0); canonical +0 */
            } else reported = o_leaf_dda(lo_, d, grid, tmin, &th, &vox);
            if (reported && tmin <= th && th <= tmax) {                 /* reportIntersectionEXT */
                out->t = th; out->leaf_ptr = ptr; out->voxel = vox;
                memcpy(out->leaf_min, cmin, sizeof cmin);
                memcpy(out->o, o, sizeof o); memcpy(out->d, d, sizeof d);
                return 1;
            }
        }
        float tx[3];
        for (int a = 0; a < 3; a++) {
            int32_t plane = d[a] > 0.0f ? cmin[a] + L : cmin[a];
            tx[a] = ((float)plane - o[a]) * inv[a];
        }
        float t_next = fminf(fminf(tx[0], tx[1]), tx[2]);
        /* inclusive: a cell entered exactly at tmax may still hold a hit at t == tmax (accepted by
         * reportIntersectionEXT); keeps the closest-hit result independent of instance order */
        if (!(t_next <= t_exit)) return 0;
        /* Leaving the occupied bounds: a stepping axis whose plane is a bounds plane (or lies beyond it,
         * for a cell that sticks out of the bounds) has exactly / at least that plane's time, and the
         * stepping axes carry the minimum, so "some stepping axis leaves [bmin, bmax)" implies
         * t_next >= t_box — the same float expressions on both sides.  The walk therefore ends on this ONE
         * comparison (no per-axis range checks), and since non-stepping axes stay inside their cell, the
         * coordinates can never leave the tree's domain either. */
        if (!(t_next < t_box)) return 0;
        for (int a = 0; a < 3; a++) {
            if (tx[a] <= t_next) {
                iv[a] = d[a] > 0.0f ? cmin[a] + L : cmin[a] - 1;
            } else {
                int32_t ip = (int32_t)floorf(fmaf(d[a], t_next, o[a]));
                iv[a] = ip < cmin[a] ? cmin[a] : (ip > cmin[a] + L - 1 ? cmin[a] + L - 1 : ip);
            }
        }
    }
}

typedef struct SceneHit {
    int hit;
    uint32_t instance;
    InstHit ih;
} SceneHit;

/* debug recorder (o_debug_pixel): every ray a pixel's paths trace, with the oracle's answer */
static __thread ORay *g_rec_rays;
static __thread OHitRec *g_rec_hits;
static __thread uint32_t g_rec_n, g_rec_cap;

static void record_ray(const OScene *s, const float o[3], const float d[3], float tmin, float tmax, int any,
                       int hit, uint32_t inst, const InstHit *ih)
{
    if (!g_rec_rays || g_rec_n >= g_rec_cap) return;
    ORay *r = &g_rec_rays[g_rec_n];
    OHitRec *h = &g_rec_hits[g_rec_n++];
    memcpy(r->o, o, 12); memcpy(r->d, d, 12); r->tmin = tmin; r->tmax = tmax;
    h->voxel = any ? 0x80000000u : 0u;                        /* bit 31 = any-hit (shadow) ray */
    if (hit) { h->t = ih->t; h->instance = inst; h->leaf = s->inst[inst].leaf_order[ih->leaf_ptr]; h->voxel |= ih->voxel; }
    else { h->t = INFINITY; h->instance = O_MISS; h->leaf = O_MISS; h->voxel |= 0xFFu; }
}

/* entry time of the ray into a padded box within [tmin, tmax], or +inf (conservative like box_may_hit) */
static float box_entry_time(const float *b, const float o[3], const float inv[3], float tmin, float tmax)
{
    float t0 = tmin, t1 = tmax;
    for (int a = 0; a < 3; a++) {
        float ta = (b[a] - o[a]) * inv[a], tb = (b[3 + a] - o[a]) * inv[a];
        t0 = fmaxf(t0, fminf(ta, tb));
        t1 = fminf(t1, fmaxf(ta, tb));
    }
    return t0 <= t1 ? t0 : INFINITY;
}

/* Closest hit over all instances: lexicographic minimum of (t, instance index) — an order-independent minimum, so the instances may be
 * visited in any order.  With culling on they are visited FRONT TO BACK by box entry time and the loop ends at the first box that starts
 * behind the best hit (a plain speed-up of the CPU baseline: the boxes only cull, tests check cull on == off and order independence). */
#define O_SORT_MAX 256
static void trace_closest(const OScene *s, const float o[3], const float d[3], float tmin, float tmax,
                          SceneHit *best, OStats *st)
{
    best->hit = 0; best->instance = O_MISS;
    float cur = tmax;
    float inv[3];
    cull_inv(d, inv);
    if (s->cull && s->n <= O_SORT_MAX) {
        float te[O_SORT_MAX];
        uint32_t id[O_SORT_MAX], m = 0;
        for (uint32_t k = 0; k < s->n; k++) {
            float t = box_entry_time(s->boxes + 6 * k, o, inv, tmin, tmax);
            if (t == INFINITY) continue;
            uint32_t j = m++;
            while (j > 0 && te[j - 1] > t) { te[j] = te[j - 1]; id[j] = id[j - 1]; j--; }      /* stable: equal times stay in index order */
            te[j] = t; id[j] = k;
        }
        for (uint32_t j = 0; j < m; j++) {
            if (te[j] > cur) break;                                    /* every remaining box starts behind the best hit */
            InstHit ih;
            uint32_t k = id[j];
            if (trace_instance(&s->inst[k], o, d, tmin, cur, &ih, st)) {
                if (!best->hit || ih.t < best->ih.t || (ih.t == best->ih.t && k < best->instance)) { best->hit = 1; best->instance = k; best->ih = ih; cur = ih.t; }
            }
        }
    } else {
        for (uint32_t k = 0; k < s->n; k++) {
            InstHit ih;
            if (s->cull && !box_may_hit(s->boxes + 6 * k, o, inv, tmin, cur)) continue;
            if (trace_instance(&s->inst[k], o, d, tmin, cur, &ih, st)) {
                if (!best->hit || ih.t < best->ih.t) { best->hit = 1; best->instance = k; best->ih = ih; cur = ih.t; }
            }
        }
    }
    record_ray(s, o, d, tmin, tmax, 0, best->hit, best->instance, &best->ih);
}

/* gl_RayFlagsTerminateOnFirstHitEXT (ambient_occlusion.rgen:37-38): any accepted hit */
static int trace_any(const OScene *s, const float o[3], const float d[3], float tmin, float tmax, OStats *st)
{
    float inv[3];
    cull_inv(d, inv);
    if (s->cull && s->n <= O_SORT_MAX) {                               /* nearest boxes first: an occluder is found sooner; the answer (any hit?) is the same */
        float te[O_SORT_MAX];
        uint32_t id[O_SORT_MAX], m = 0;
        for (uint32_t k = 0; k < s->n; k++) {
            float t = box_entry_time(s->boxes + 6 * k, o, inv, tmin, tmax);
            if (t == INFINITY) continue;
            uint32_t j = m++;
            while (j > 0 && te[j - 1] > t) { te[j] = te[j - 1]; id[j] = id[j - 1]; j--; }
            te[j] = t; id[j] = k;
        }
        for (uint32_t j = 0; j < m; j++) {
            InstHit ih;
            if (trace_instance(&s->inst[id[j]], o, d, tmin, tmax, &ih, st)) { record_ray(s, o, d, tmin, tmax, 1, 1, id[j], &ih); return 1; }
        }
        record_ray(s, o, d, tmin, tmax, 1, 0, O_MISS, NULL);
        return 0;
    }
    for (uint32_t k = 0; k < s->n; k++) {
        InstHit ih;
        if (s->cull && !box_may_hit(s->boxes + 6 * k, o, inv, tmin, tmax)) continue;
        if (trace_instance(&s->inst[k], o, d, tmin, tmax, &ih, st)) { record_ray(s, o, d, tmin, tmax, 1, 1, k, &ih); return 1; }
    }
    record_ray(s, o, d, tmin, tmax, 1, 0, O_MISS, NULL);
    return 0;
}

/* ------------------------------------------------------------------ camera */

/* assets/shaders/headers/camera.glsl:1-16; jitter (0.5,0.5) reproduces the reference exactly */
void o_camera_ray(const OCamera *cam, uint32_t width, uint32_t height, uint32_t px, uint32_t py,
                  float jx, float jy, float o[3], float d[3])
{
    float ndc_x = ((float)px + jx) / (float)width;
    float ndc_y = ((float)py + jy) / (float)height;
    float cx = 2.0f * ndc_x - 1.0f;
    float cy = 2.0f * ndc_y - 1.0f;
    cy = cy * -1.0f;
    cx = cx * ((float)width / (float)height);
    cx = cx * cam->tan_half_fov;
    cy = cy * cam->tan_half_fov;
    for (int i = 0; i < 3; i++)
        d[i] = (cam->camera_view_col0[i] * cx + cam->camera_view_col1[i] * cy) + cam->camera_view_col2[i] * -1.0f;
    o[0] = cam->position_x; o[1] = cam->position_y; o[2] = cam->position_z;
}

/* ----------------------------------------------------------------- shading */

typedef struct Surface {
    float x[3];        /* world hit point + 0.01 * n */
    float n[3];        /* world normal */
    float albedo[3];   /* palette rgb / 255 (non-linear sRGB) */
    uint32_t leaf, voxel, palette_index;
} Surface;

/* assets/shaders/primary/hit.rchit:16-76 for the hit `h` of world ray (o, d) */
static void surface_of(const OScene *s, const SceneHit *h, const float o[3], const float d[3], Surface *sf, OStats *st)
{
    const OInstance *in = &s->inst[h->instance];
    const InstHit *ih = &h->ih;
    float n_obj[3], v[3], av[3];
    for (int a = 0; a < 3; a++) {
        float hp = ih->t * ih->d[a] + ih->o[a];                                      /* :20 */
        float off = (float)((ih->voxel >> (4 - 2 * a)) & 3u);                        /* :21 */
        float center = ((float)ih->leaf_min[a] + off) + 0.5f;                        /* :26 */
        v[a] = hp - center;
        av[a] = fabsf(v[a]);
    }
    float mx = fmaxf(av[0], fmaxf(av[1], av[2]));                                    /* normal.glsl:39-43 */
    for (int a = 0; a < 3; a++) n_obj[a] = signf(v[a]) * (av[a] >= mx ? 1.0f : 0.0f);
    const float *W = in->obj2world;
    for (int a = 0; a < 3; a++)                                                      /* :30 */
        sf->n[a] = (W[4 * a] * n_obj[0] + W[4 * a + 1] * n_obj[1]) + W[4 * a + 2] * n_obj[2];
    const uint8_t *leaf = o_pool_get(&in->tree->pool[0], ih->leaf_ptr);
    uint64_t mask; uint32_t material_ptr;
    memcpy(&mask, leaf, 8); memcpy(&material_ptr, leaf + 16, 4);
    uint32_t off = (uint32_t)__builtin_popcountll(mask & (((uint64_t)1 << ih->voxel) - 1u));   /* :60-61 */
    uint8_t pi = in->materials[material_ptr + off];                                  /* :71 */
    const uint8_t *col = in->palette + 4 * (size_t)pi;                               /* :72 */
    for (int a = 0; a < 3; a++) sf->albedo[a] = (float)col[a] / 255.0f;              /* :74 */
    for (int a = 0; a < 3; a++)                                                      /* ambient_occlusion.rgen:23 */
        sf->x[a] = (ih->t * d[a] + o[a]) + sf->n[a] * 0.01f;
    sf->leaf = in->leaf_order[ih->leaf_ptr];
    sf->voxel = ih->voxel;
    sf->palette_index = pi;
    st->hits_shaded++;
}

typedef struct Frame {
    const OScene *s; const OCamera *cam; const OSky *sky; const OTraceConfig *cfg;
    float *radiance; uint32_t *hit_instance, *hit_leaf; uint8_t *hit_voxel; float *hit_t;
    float sun_dir[3];        /* normalize(sunlight_config.direction.xyz) */
    float sun_cone;          /* 1 - cos(solar radius), nee.rmiss:18-19 */
    const uint32_t *rows;    /* o_render_rows: the rows to render (NULL: cfg->y0 .. cfg->y1) */
    uint32_t n_rows;
    const OGBuffer *gb;
} Frame;

/* float -> UNORM: round to nearest (even) of clamp(x, 0, 1) * (2^bits - 1) */
static uint32_t unorm(float x, float maxv) { return (uint32_t)rintf(fminf(fmaxf(x, 0.0f), 1.0f) * maxv); }
/* A2R10G10B10_UNORM_PACK32: A 31..30, R 29..20, G 19..10, B 9..0 */
static uint32_t pack_a2r10g10b10(float r, float g, float b, float a)
{
    return (unorm(a, 3.0f) << 30) | (unorm(r, 1023.0f) << 20) | (unorm(g, 1023.0f) << 10) | unorm(b, 1023.0f);
}

/* the image stores of hit.rchit:57-94 for the primary hit `h` of world ray (o, d) */
static void gbuffer_hit(const Frame *F, uint32_t pi, const SceneHit *h, const Surface *sf, const float o[3], const float d[3])
{
    const OGBuffer *gb = F->gb;
    const OInstance *in = &F->s->inst[h->instance];
    if (gb->albedo) gb->albedo[pi] = pack_a2r10g10b10(sf->albedo[0], sf->albedo[1], sf->albedo[2], 1.0f);       /* :71-76 */
    if (gb->depth) gb->depth[pi] = h->ih.t;                                                                    /* :81 */
    if (gb->normal) {
        /* NRD_FrontEnd_PackNormalAndRoughness(normalWorld, 1.0, float(palette_index)), headers/nrd.glsl:1-52 */
        const float *n = sf->n;
        float s = (fabsf(n[0]) + fabsf(n[1])) + fabsf(n[2]);
        float vx = n[0] / s, vy = n[1] / s, vz = n[2] / s;
        float wx = (1.0f - fabsf(vy)) * ((vx < 0.0f ? 0.0f : 1.0f) * 2.0f - 1.0f);
        float wy = (1.0f - fabsf(vx)) * ((vy < 0.0f ? 0.0f : 1.0f) * 2.0f - 1.0f);
        float ex = vz >= 0.0f ? vx : wx, ey = vz >= 0.0f ? vy : wy;
        gb->normal[pi] = pack_a2r10g10b10(ex * 0.5f + 0.5f, ey * 0.5f + 0.5f, 1.0f,
                                          fminf(fmaxf((float)sf->palette_index / 3.0f, 0.0f), 1.0f));         /* :83 */
    }
    if (gb->voxel_id) gb->voxel_id[pi] = (sf->voxel << 24) | (h->instance & 0xFFFFu) | (sf->palette_index << 16);   /* :87 */
    if (gb->motion) {                                                                                          /* :90-94 */
        float hw[3], hm[3];
        for (int a = 0; a < 3; a++) hw[a] = h->ih.t * d[a] + o[a];
        for (int a = 0; a < 3; a++) {
            const float *M = in->w2o + 4 * a;
            hm[a] = ((M[0] * hw[0] + M[1] * hw[1]) + M[2] * hw[2]) + M[3];
        }
        for (int a = 0; a < 3; a++) {
            const float *P = in->prev_obj2world + 4 * a;
            float last = ((P[0] * hm[0] + P[1] * hm[1]) + P[2] * hm[2]) + P[3];
            gb->motion[4 * (size_t)pi + a] = last - hw[a];
        }
        gb->motion[4 * (size_t)pi + 3] = 0.0f;
    }
}

/* miss.rmiss:13-16 */
static void gbuffer_miss(const Frame *F, uint32_t pi)
{
    const OGBuffer *gb = F->gb;
    if (gb->albedo) gb->albedo[pi] = 0xFFFFFFFFu;
    if (gb->depth) gb->depth[pi] = INFINITY;
    if (gb->motion) for (int a = 0; a < 4; a++) gb->motion[4 * (size_t)pi + a] = 0.0f;
}

static void shade_pixel(const Frame *F, uint32_t px, uint32_t py, OStats *st)
{
    const OTraceConfig *cfg = F->cfg;
    uint32_t pixel_index = py * cfg->width + px;
    uint32_t key = o_pcg(cfg->seed + 0x9E3779B9u * cfg->frame_index);
    uint32_t s0 = o_pcg(pixel_index ^ key);
    float acc[3] = { 0.0f, 0.0f, 0.0f };
    for (uint32_t smp = 0; smp < cfg->spp; smp++) {
        uint32_t s1 = o_pcg(s0 + smp);
        float jx = 0.5f, jy = 0.5f;
        if (smp > 0) { jx = u01(o_pcg(s1 + 0u)); jy = u01(o_pcg(s1 + 1u)); }
        float o[3], d[3];
        o_camera_ray(F->cam, cfg->width, cfg->height, px, py, jx, jy, o, d);
        SceneHit h;
        trace_closest(F->s, o, d, F->cam->near, F->cam->far, &h, st);     /* primary.rgen:9-21 */
        st->primary_rays++;
        float L[3] = { 0.0f, 0.0f, 0.0f };        /* accumulated in linear sRGB primaries */
        if (!h.hit) {
            /* primary/miss.rmiss:8-14 — (sky + sun) / 3.14, albedo = 1 */
            float dn[3], sk[3], su[3], term[3];
            normalize3(d, dn);
            o_sky_radiance(F->sky, dn, sk);
            o_sun_radiance(F->sky, dn, su);
            for (int a = 0; a < 3; a++) term[a] = (sk[a] + su[a]) / 3.14f;
            o_acescg2srgb(term, L);
            if (smp == 0 && F->hit_instance) {
                F->hit_instance[pixel_index] = O_MISS; F->hit_leaf[pixel_index] = O_MISS;
                F->hit_voxel[pixel_index] = 0xFF; F->hit_t[pixel_index] = INFINITY;      /* miss.rmiss:15 depth = 1/0 */
            }
            if (smp == 0 && F->gb) gbuffer_miss(F, pixel_index);
        } else {
            Surface sf;
            surface_of(F->s, &h, o, d, &sf, st);
            if (smp == 0 && F->hit_instance) {
                F->hit_instance[pixel_index] = h.instance; F->hit_leaf[pixel_index] = sf.leaf;
                F->hit_voxel[pixel_index] = (uint8_t)sf.voxel; F->hit_t[pixel_index] = h.ih.t;
            }
            if (smp == 0 && F->gb) gbuffer_hit(F, pixel_index, &h, &sf, o, d);
            float T[3];                             /* throughput, starts at the primary albedo (tone_map.comp:195-203) */
            for (int a = 0; a < 3; a++) T[a] = (cfg->flags & O_TRACE_DEMODULATE) ? 1.0f : o_srgb_to_linear(sf.albedo[a]);
            for (uint32_t v = 0;; v++) {
                /* sun next-event estimation: ambient_occlusion.rgen:33-50, nee.rmiss:11-22 */
                if (dot3(F->sky->direction, sf.n) > 0.0f) {
                    st->shadow_rays++;
                    if (!trace_any(F->s, sf.x, F->sun_dir, 0.1f, 10000.0f, st)) {
                        float dn[3], su[3], term[3], ts[3];
                        normalize3(F->sun_dir, dn);
                        o_sun_radiance(F->sky, dn, su);
                        float ndl = dot3(sf.n, F->sun_dir);
                        for (int a = 0; a < 3; a++) term[a] = (su[a] * F->sun_cone) * ndl;
                        o_acescg2srgb(term, ts);
                        for (int a = 0; a < 3; a++) L[a] += T[a] * ts[a];
                    }
                }
                if (v == cfg->max_bounces) break;
                /* cosine-weighted bounce: final_gather.rgen:32-51 (blue noise -> pcg, deviation 3) */
                float u1 = u01(o_pcg(s1 + 2u + 2u * v)), u2 = u01(o_pcg(s1 + 3u + 2u * v));
                float r = sqrtf(u1), cz = sqrtf(1.0f - u1), c, sn;
                o_cossin_2pi(u2, &c, &sn);
                float local[3] = { r * c, r * sn, cz }, rot[3], w[3];
                o_rotate_by_normal(sf.n, local, rot);
                normalize3(rot, w);
                SceneHit bh;
                trace_closest(F->s, sf.x, w, 0.1f, F->cam->far, &bh, st);
                st->bounce_rays++;
                if (!bh.hit) {
                    /* final_gather.rmiss:13-17 */
                    float dn[3], sk[3], ts[3];
                    normalize3(w, dn);
                    o_sky_radiance(F->sky, dn, sk);
                    o_acescg2srgb(sk, ts);
                    for (int a = 0; a < 3; a++) L[a] += T[a] * ts[a];
                    break;
                }
                /* final_gather.rchit:76-80 — outgoing = albedo_lin (x) incoming, in sRGB primaries */
                float xo[3] = { sf.x[0], sf.x[1], sf.x[2] };
                surface_of(F->s, &bh, xo, w, &sf, st);
                for (int a = 0; a < 3; a++) T[a] = T[a] * o_srgb_to_linear(sf.albedo[a]);
            }
        }
        float Lc[3];
        o_srgb2acescg(L, Lc);
        for (int a = 0; a < 3; a++) acc[a] += Lc[a];
    }
    float *out = F->radiance + 4 * (size_t)pixel_index;
    for (int a = 0; a < 3; a++) out[a] = acc[a] / (float)cfg->spp;
    out[3] = 1.0f;
}

/* ----------------------------------------------------------------- threads */

typedef struct Job {
    const Frame *F;
    const OScene *s; const ORay *rays; OHitRec *hits; uint64_t n;
    uint32_t tid, nthreads;
    OStats st;
} Job;

static void *render_worker(void *arg)
{
    Job *j = (Job *)arg;
    const OTraceConfig *cfg = j->F->cfg;
    OStats st;                      /* thread-local: the jobs array is contiguous, counting in place would share cache lines */
    memset(&st, 0, sizeof st);
    /* 32-pixel spans of the row range, dealt round-robin: even a single row keeps every thread busy */
    const uint32_t n_rows = j->F->rows ? j->F->n_rows : cfg->y1 - cfg->y0;
    const uint64_t total = (uint64_t)n_rows * cfg->width, span = 32;
    for (uint64_t base = (uint64_t)j->tid * span; base < total; base += (uint64_t)j->nthreads * span) {
        uint64_t end = base + span < total ? base + span : total;
        for (uint64_t i = base; i < end; i++) {
            const uint32_t r = (uint32_t)(i / cfg->width);
            shade_pixel(j->F, (uint32_t)(i % cfg->width), j->F->rows ? j->F->rows[r] : cfg->y0 + r, &st);
        }
    }
    j->st = st;
    return NULL;
}

static void *rays_worker(void *arg)
{
    Job *j = (Job *)arg;
    OStats st;
    memset(&st, 0, sizeof st);
    const uint64_t chunk = 4096;
    for (uint64_t base = (uint64_t)j->tid * chunk; base < j->n; base += (uint64_t)j->nthreads * chunk) {
        uint64_t end = base + chunk < j->n ? base + chunk : j->n;
        for (uint64_t i = base; i < end; i++) {
            const ORay *r = &j->rays[i];
            SceneHit h;
            trace_closest(j->s, r->o, r->d, r->tmin, r->tmax, &h, &st);
            st.primary_rays++;
            OHitRec *o = &j->hits[i];
            if (h.hit) {
                o->t = h.ih.t; o->instance = h.instance;
                o->leaf = j->s->inst[h.instance].leaf_order[h.ih.leaf_ptr]; o->voxel = h.ih.voxel;
            } else {
                o->t = INFINITY; o->instance = O_MISS; o->leaf = O_MISS; o->voxel = O_MISS;
            }
        }
    }
    j->st = st;
    return NULL;
}

static void run_jobs(Job *jobs, uint32_t n_threads, void *(*fn)(void *), OStats *stats)
{
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    for (uint32_t i = 1; i < n_threads; i++) pthread_create(&th[i], NULL, fn, &jobs[i]);
    fn(&jobs[0]);
    for (uint32_t i = 1; i < n_threads; i++) pthread_join(th[i], NULL);
    free(th);
    if (stats) {
        memset(stats, 0, sizeof *stats);
        for (uint32_t i = 0; i < n_threads; i++) {
            stats->primary_rays += jobs[i].st.primary_rays; stats->bounce_rays += jobs[i].st.bounce_rays;
            stats->shadow_rays += jobs[i].st.shadow_rays; stats->mask_words += jobs[i].st.mask_words;
            stats->child_ptrs += jobs[i].st.child_ptrs; stats->hits_shaded += jobs[i].st.hits_shaded;
        }
    }
}

void o_render(const OScene *s, const OCamera *cam, const OSky *sky, const OTraceConfig *cfg,
              float *radiance, uint32_t *hit_instance, uint32_t *hit_leaf, uint8_t *hit_voxel, float *hit_t,
              OStats *stats, uint32_t n_threads)
{
    o_render_rows(s, cam, sky, cfg, NULL, 0, radiance, hit_instance, hit_leaf, hit_voxel, hit_t, stats, n_threads);
}

/* As o_render, for an explicit list of rows (rows == NULL: cfg->y0 .. cfg->y1): bench.py's bounded CPU
 * sample is spread over the frame and wants ONE call (one set of threads) for all of its rows. */
void o_render_rows(const OScene *s, const OCamera *cam, const OSky *sky, const OTraceConfig *cfg,
                   const uint32_t *rows, uint32_t n_rows,
                   float *radiance, uint32_t *hit_instance, uint32_t *hit_leaf, uint8_t *hit_voxel, float *hit_t,
                   OStats *stats, uint32_t n_threads)
{
    o_render_g(s, cam, sky, cfg, rows, n_rows, radiance, hit_instance, hit_leaf, hit_voxel, hit_t, NULL, stats, n_threads);
}

void o_render_g(const OScene *s, const OCamera *cam, const OSky *sky, const OTraceConfig *cfg,
                const uint32_t *rows, uint32_t n_rows,
                float *radiance, uint32_t *hit_instance, uint32_t *hit_leaf, uint8_t *hit_voxel, float *hit_t,
                const OGBuffer *gbuffer, OStats *stats, uint32_t n_threads)
{
    Frame F = { s, cam, sky, cfg, radiance, hit_instance, hit_leaf, hit_voxel, hit_t, { 0, 0, 0 }, 0.0f, rows, n_rows, gbuffer };
    normalize3(sky->direction, F.sun_dir);                              /* ambient_occlusion.rgen:45 */
    F.sun_cone = 1.0f - cosf(sky->sunlight_intensity[3]);
    if (n_threads < 1) n_threads = 1;
    Job *jobs = (Job *)calloc(n_threads, sizeof(Job));
    for (uint32_t i = 0; i < n_threads; i++) { jobs[i].F = &F; jobs[i].tid = i; jobs[i].nthreads = n_threads; }
    run_jobs(jobs, n_threads, render_worker, stats);
    free(jobs);
}

/* Debug aid: shade ONE pixel (all samples) and record every ray its paths trace, with the oracle's
 * answer (voxel bit 31 set = any-hit shadow ray).  Returns the number of rays recorded. */
uint32_t o_debug_pixel(const OScene *s, const OCamera *cam, const OSky *sky, const OTraceConfig *cfg,
                       uint32_t px, uint32_t py, ORay *rays, OHitRec *hits, uint32_t cap, float rgba[4])
{
    float *rad = (float *)calloc((size_t)cfg->width * cfg->height, 4 * sizeof(float));
    Frame F = { s, cam, sky, cfg, rad, NULL, NULL, NULL, NULL, { 0, 0, 0 }, 0.0f, NULL, 0, NULL };
    normalize3(sky->direction, F.sun_dir);
    F.sun_cone = 1.0f - cosf(sky->sunlight_intensity[3]);
    OStats st;
    memset(&st, 0, sizeof st);
    g_rec_rays = rays; g_rec_hits = hits; g_rec_n = 0; g_rec_cap = cap;
    shade_pixel(&F, px, py, &st);
    g_rec_rays = NULL; g_rec_hits = NULL;
    memcpy(rgba, rad + 4 * ((size_t)py * cfg->width + px), 4 * sizeof(float));
    free(rad);
    return g_rec_n;
}

void o_trace_rays(const OScene *s, const ORay *rays, uint64_t n, OHitRec *hits, OStats *stats, uint32_t n_threads)
{
    if (n_threads < 1) n_threads = 1;
    Job *jobs = (Job *)calloc(n_threads, sizeof(Job));
    for (uint32_t i = 0; i < n_threads; i++) {
        jobs[i].s = s; jobs[i].rays = rays; jobs[i].hits = hits; jobs[i].n = n;
        jobs[i].tid = i; jobs[i].nthreads = n_threads;
    }
    run_jobs(jobs, n_threads, rays_worker, stats);
    free(jobs);
}

#include "oracle_gi.c"    /* the GI passes use this file's static traversal and shading functions */
