/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_vox.h).
 * CPU restatement of crates/vox/src/{loader.rs, collector.rs, geometry.rs} and of the subset of
 * dot_vox 5.1.1 they call.  Citations are reference file:line.
 */
#include "oracle_vox.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* ------------------------------------------------------------------ parsing */

typedef struct { const uint8_t *p; uint64_t len, pos; int err; } Rd;

static int32_t rd_i32(Rd *r)
{
    if (r->pos + 4 > r->len) { r->err = 1; return 0; }
    int32_t v; memcpy(&v, r->p + r->pos, 4); r->pos += 4; return v;
}

/* DICT value lookup: returns 1 and copies the value (NUL-terminated, truncated) when `key` is present */
static int rd_dict_find(Rd *r, const char *key, char *val, size_t cap)
{
    int found = 0;
    int32_t n = rd_i32(r);
    for (int32_t k = 0; k < n && !r->err; k++) {
        int32_t kl = rd_i32(r);
        if (kl < 0 || r->pos + (uint64_t)kl > r->len) { r->err = 1; break; }
        const uint8_t *kp = r->p + r->pos; r->pos += (uint64_t)kl;
        int32_t vl = rd_i32(r);
        if (vl < 0 || r->pos + (uint64_t)vl > r->len) { r->err = 1; break; }
        const uint8_t *vp = r->p + r->pos; r->pos += (uint64_t)vl;
        if (key && (size_t)kl == strlen(key) && memcmp(kp, key, (size_t)kl) == 0) {
            size_t c = (size_t)vl < cap - 1 ? (size_t)vl : cap - 1;
            memcpy(val, vp, c); val[c] = 0; found = 1;
        }
    }
    return found;
}

static void scan_dict_frame(Rd *r, OVoxNode *nd)
{
    /* one frame DICT: `_t` = "x y z", `_r` = decimal byte (dot_vox Frame::position/orientation) */
    uint64_t save = r->pos;
    char buf[96];
    if (rd_dict_find(r, "_t", buf, sizeof buf)) {
        long a = 0, b = 0, c = 0;
        char *e = buf;
        a = strtol(e, &e, 10); b = strtol(e, &e, 10); c = strtol(e, &e, 10);
        nd->translation[0] = (int32_t)a; nd->translation[1] = (int32_t)b; nd->translation[2] = (int32_t)c;
    }
    uint64_t end = r->pos;
    r->pos = save;
    if (rd_dict_find(r, "_r", buf, sizeof buf)) nd->rotation = (uint8_t)strtol(buf, NULL, 10);
    r->pos = end;
}

OVoxScene *o_vox_parse(const uint8_t *bytes, uint64_t len)
{
    Rd r = { bytes, len, 0, 0 };
    if (len < 8 || memcmp(bytes, "VOX ", 4) != 0) return NULL;
    r.pos = 8;                                   /* magic + version */
    if (r.pos + 12 > len || memcmp(bytes + r.pos, "MAIN", 4) != 0) return NULL;
    r.pos += 4;
    int32_t main_content = rd_i32(&r);
    (void)rd_i32(&r);
    r.pos += (uint64_t)main_content;

    OVoxScene *s = (OVoxScene *)calloc(1, sizeof *s);
    uint32_t cap_models = 0, cap_nodes = 0;
    int have_size = 0; uint32_t cur_size[3] = { 0, 0, 0 };
    int have_rgba = 0;

    while (r.pos + 12 <= len && !r.err) {
        char id[5] = { 0 };
        memcpy(id, bytes + r.pos, 4); r.pos += 4;
        int32_t content = rd_i32(&r);
        int32_t children = rd_i32(&r);
        if (content < 0 || children < 0 || r.pos + (uint64_t)content > len) { r.err = 1; break; }
        uint64_t end = r.pos + (uint64_t)content;
        Rd c = { bytes, end, r.pos, 0 };
        if (!strcmp(id, "SIZE")) {
            cur_size[0] = (uint32_t)rd_i32(&c); cur_size[1] = (uint32_t)rd_i32(&c); cur_size[2] = (uint32_t)rd_i32(&c);
            have_size = 1;
        } else if (!strcmp(id, "XYZI") && have_size) {
            int32_t n = rd_i32(&c);
            if (n < 0 || c.pos + 4ull * (uint64_t)n > end) { r.err = 1; break; }
            if (s->n_models == cap_models) {
                cap_models = cap_models ? cap_models * 2 : 16;
                s->models = (OVoxModel *)realloc(s->models, sizeof(OVoxModel) * cap_models);
            }
            OVoxModel *m = &s->models[s->n_models++];
            memcpy(m->size, cur_size, sizeof cur_size);
            m->n_voxels = (uint32_t)n;
            m->xyzi = (uint8_t *)malloc(4ull * (uint64_t)n + 4);
            memcpy(m->xyzi, bytes + c.pos, 4ull * (uint64_t)n);
            /* dot_vox: Voxel.i = raw index - 1 (saturating) */
            for (int32_t k = 0; k < n; k++) {
                uint8_t *i = &m->xyzi[4 * (size_t)k + 3];
                *i = *i ? (uint8_t)(*i - 1) : 0;
            }
            have_size = 0;
        } else if (!strcmp(id, "RGBA")) {
            if (content >= 1024) { memcpy(s->palette, bytes + c.pos, 1024); have_rgba = 1; }
        } else if (!strcmp(id, "nTRN") || !strcmp(id, "nGRP") || !strcmp(id, "nSHP")) {
            int32_t node_id = rd_i32(&c);
            if (node_id < 0 || (uint32_t)node_id != s->n_nodes) { r.err = 1; break; }  /* ids are file-ordered */
            if (s->n_nodes == cap_nodes) {
                cap_nodes = cap_nodes ? cap_nodes * 2 : 64;
                s->nodes = (OVoxNode *)realloc(s->nodes, sizeof(OVoxNode) * cap_nodes);
            }
            OVoxNode *nd = &s->nodes[s->n_nodes++];
            memset(nd, 0, sizeof *nd);
            nd->rotation = 4;                    /* Rotation::IDENTITY = 0b0000100 */
            char tmp[8];
            rd_dict_find(&c, NULL, tmp, sizeof tmp);          /* node attributes, unused */
            if (id[1] == 'T') {
                nd->kind = O_NODE_TRANSFORM;
                nd->child = rd_i32(&c);
                (void)rd_i32(&c);                /* reserved */
                (void)rd_i32(&c);                /* layer id */
                int32_t frames = rd_i32(&c);
                if (frames != 1) { r.err = 1; break; }        /* loader.rs:104-106 unimplemented!() */
                scan_dict_frame(&c, nd);
            } else if (id[1] == 'G') {
                nd->kind = O_NODE_GROUP;
                int32_t n = rd_i32(&c);
                if (n < 0 || c.pos + 4ull * (uint64_t)n > end) { r.err = 1; break; }
                nd->n_children = (uint32_t)n;
                nd->children = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n ? n : 1));
                for (int32_t k = 0; k < n; k++) nd->children[k] = rd_i32(&c);
            } else {
                nd->kind = O_NODE_SHAPE;
                int32_t n = rd_i32(&c);
                if (n != 1) { r.err = 1; break; }             /* loader.rs:149-151 unimplemented!() */
                nd->model_id = rd_i32(&c);
                rd_dict_find(&c, NULL, tmp, sizeof tmp);
            }
            if (c.err) { r.err = 1; break; }
        }
        r.pos = end;                              /* children of non-MAIN chunks are always 0 bytes */
    }
    if (r.err) { o_vox_scene_free(s); return NULL; }
    if (!have_rgba) {
        /* dot_vox falls back to MagicaVoxel's default palette (published format description: colour 1..215 = 6x6x6 cube over
         * ff cc 99 66 33 00, blue fastest, black omitted; 216..255 = red / green / blue / grey ramps); entry k = colour k + 1 */
        static const uint8_t ramp[10] = { 0xee, 0xdd, 0xbb, 0xaa, 0x88, 0x77, 0x55, 0x44, 0x22, 0x11 };
        memset(s->palette, 0, 1024);
        for (int i = 1; i < 256; i++) {
            uint8_t *c = s->palette[i - 1];
            if (i <= 215) {
                int k = i - 1;
                c[2] = (uint8_t)(0xff - 0x33 * (k % 6)); c[1] = (uint8_t)(0xff - 0x33 * ((k / 6) % 6)); c[0] = (uint8_t)(0xff - 0x33 * (k / 36));
            } else {
                int k = i - 216, v = ramp[k % 10];
                if (k < 10) c[0] = (uint8_t)v; else if (k < 20) c[1] = (uint8_t)v; else if (k < 30) c[2] = (uint8_t)v; else c[0] = c[1] = c[2] = (uint8_t)v;
            }
            c[3] = 255;
        }
    }
    return s;
}

void o_vox_scene_free(OVoxScene *s)
{
    if (!s) return;
    for (uint32_t i = 0; i < s->n_models; i++) free(s->models[i].xyzi);
    for (uint32_t i = 0; i < s->n_nodes; i++) free(s->nodes[i].children);
    free(s->models); free(s->nodes); free(s);
}

/* --------------------------------------------------------------- transforms */

static void affine_identity(float m[12])
{
    memset(m, 0, sizeof(float) * 12);
    m[0] = m[5] = m[10] = 1.0f;
}

/* out = a * b for row-major 3x4 affines (glam Affine3A multiplication, bevy GlobalTransform) */
static void affine_mul(const float a[12], const float b[12], float out[12])
{
    float r[12];
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++)
            r[4 * i + j] = a[4 * i] * b[j] + a[4 * i + 1] * b[4 + j] + a[4 * i + 2] * b[8 + j];
        r[4 * i + 3] = a[4 * i] * b[3] + a[4 * i + 1] * b[7] + a[4 * i + 2] * b[11] + a[4 * i + 3];
    }
    memcpy(out, r, sizeof r);
}

/*
 * loader.rs:178-204.  dot_vox's Rotation is a signed permutation matrix M (row r has its
 * non-zero, sign_r, in column idx_r).  `to_quat_scale()` factors it as M = R * (s*I) with
 * s = det(M); the loader re-expresses the quaternion in Bevy axes with (x, z, -y, w), i.e.
 * R' = C R C^-1 for C:(x,y,z)->(x,z,-y), and swizzles the scale.  The linear part of the
 * resulting Transform is therefore C M C^-1 with entries in {0,+1,-1} (the reference's f32
 * quaternion->matrix round trip leaves ~6e-8 residue on 90-degree turns; we keep exact values).
 */
void o_vox_to_transform(const int32_t translation[3], uint8_t rotation, const uint32_t size[3], float out[12])
{
    int i0 = rotation & 3, i1 = (rotation >> 2) & 3;
    if (i0 == i1 || i0 == 3 || i1 == 3) { i0 = 0; i1 = 1; rotation = 4; }   /* invalid -> identity */
    int i2 = 3 - i0 - i1;
    float M[3][3] = { { 0 } };
    M[0][i0] = (rotation & 16) ? -1.0f : 1.0f;
    M[1][i1] = (rotation & 32) ? -1.0f : 1.0f;
    M[2][i2] = (rotation & 64) ? -1.0f : 1.0f;
    float det = M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1])
              - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0])
              + M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
    float s = det < 0.0f ? -1.0f : 1.0f;
    static const float C[3][3] = { { 1, 0, 0 }, { 0, 0, 1 }, { 0, -1, 0 } };
    float CM[3][3], L[3][3];                                  /* L = C M C^T */
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
        CM[i][j] = C[i][0] * M[0][j] + C[i][1] * M[1][j] + C[i][2] * M[2][j];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
        L[i][j] = CM[i][0] * C[j][0] + CM[i][1] * C[j][1] + CM[i][2] * C[j][2];
    /* R' = s * L (proper rotation) */
    float tr[3] = { (float)translation[0], (float)translation[2], -(float)translation[1] };   /* .xzy, z *= -1 */
    float off[3] = { size[0] % 2 == 0 ? 0.0f : 0.5f, size[2] % 2 == 0 ? 0.0f : 0.5f, size[1] % 2 == 0 ? 0.0f : -0.5f };
    float half[3] = { (float)size[0] / 2.0f, (float)size[2] / 2.0f, (float)size[1] / 2.0f };  /* size.xzy / 2 */
    for (int i = 0; i < 3; i++) {
        float center = s * (L[i][0] * half[0] + L[i][1] * half[1] + L[i][2] * half[2]);   /* quat * v */
        float offset = s * (L[i][0] * off[0] + L[i][1] * off[1] + L[i][2] * off[2]);
        out[4 * i + 0] = L[i][0]; out[4 * i + 1] = L[i][1]; out[4 * i + 2] = L[i][2];
        out[4 * i + 3] = tr[i] - center * s + offset;          /* translation - center*scale + offset */
    }
}

typedef struct { const OVoxScene *s; OVoxInstance *out; uint32_t cap, n; } Walk;

/* loader.rs:86-176 */
static void walk(Walk *w, int32_t node, const float parent[12], const int32_t translation[3], uint8_t rotation)
{
    const OVoxScene *s = w->s;
    if (node < 0 || (uint32_t)node >= s->n_nodes) return;
    const OVoxNode *nd = &s->nodes[node];
    if (nd->kind == O_NODE_TRANSFORM) {
        int32_t t[3] = { translation[0] + nd->translation[0], translation[1] + nd->translation[1],
                         translation[2] + nd->translation[2] };
        walk(w, nd->child, parent, t, nd->rotation);           /* rotation is replaced, not composed (:120) */
    } else if (nd->kind == O_NODE_GROUP) {
        static const uint32_t zero[3] = { 0, 0, 0 };
        static const int32_t zt[3] = { 0, 0, 0 };
        float local[12], global[12];
        o_vox_to_transform(translation, rotation, zero, local);
        affine_mul(parent, local, global);
        for (uint32_t k = 0; k < nd->n_children; k++) walk(w, nd->children[k], global, zt, 4);
    } else {
        if (nd->model_id < 0 || (uint32_t)nd->model_id >= s->n_models) return;
        const OVoxModel *m = &s->models[nd->model_id];
        if (m->n_voxels == 0) return;                          /* :154-156 */
        float local[12];
        o_vox_to_transform(translation, rotation, m->size, local);
        if (w->n < w->cap) {
            w->out[w->n].model_id = (uint32_t)nd->model_id;
            affine_mul(parent, local, w->out[w->n].obj2world);
        }
        w->n++;
    }
}

uint32_t o_vox_instances(const OVoxScene *s, OVoxInstance *out, uint32_t cap)
{
    Walk w = { s, out, cap, 0 };
    float ident[12];
    affine_identity(ident);
    if (s->n_nodes == 0) {
        /* loader.rs:67-83 — no scene graph: the single model with the default transform */
        if (s->n_models >= 1 && s->models[0].n_voxels) {
            if (cap) { out[0].model_id = 0; memcpy(out[0].obj2world, ident, sizeof ident); }
            w.n = 1;
        }
        return w.n;
    }
    static const int32_t zt[3] = { 0, 0, 0 };
    walk(&w, 0, ident, zt, 4);
    return w.n;
}

/* --------------------------------------------------------------- load_model */

static float linear2srgb(float c)
{
    /* geometry.rs:96-102 */
    if (c <= 0.0031308f) return 12.92f * c;
    return 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
}

OVoxModelData *o_vox_load_model(const OVoxScene *s, uint32_t model_id)
{
    if (model_id >= s->n_models) return NULL;
    const OVoxModel *model = &s->models[model_id];
    if (model->size[0] > 256 || model->size[1] > 256 || model->size[2] > 256) return NULL;   /* loader.rs:365 */
    static const uint32_t H[3] = { 4, 2, 2 };                 /* crates/vox/src/lib.rs:19 */
    OVoxModelData *out = (OVoxModelData *)calloc(1, sizeof *out);
    out->tree = o_tree_new(3, H);
    out->size[0] = (uint8_t)model->size[0]; out->size[1] = (uint8_t)model->size[2]; out->size[2] = (uint8_t)model->size[1];

    /* collector.rs:9-22 */
    uint8_t *grid = (uint8_t *)calloc(256u * 256u * 256u, 1);
    uint32_t *block_counts = (uint32_t *)calloc(64u * 64u * 64u, sizeof(uint32_t));
    uint64_t count = 0;

    for (uint32_t k = 0; k < model->n_voxels; k++) {
        const uint8_t *v = &model->xyzi[4 * (size_t)k];
        /* loader.rs:247-253 swizzle */
        uint8_t x = v[0], y = v[2], z = (uint8_t)(model->size[1] - (uint32_t)v[1] - 1u), i = v[3];
        o_tree_set_value(out->tree, x, y, z, 1);               /* :259 */
        /* collector.rs:23-34 */
        count += 1;
        size_t block_index = (size_t)(x >> 2) + (size_t)(y >> 2) * 64 + (size_t)(z >> 2) * 64 * 64;
        block_counts[block_index] += 1;
        uint32_t index = (uint32_t)(z & 3) | ((uint32_t)(y & 3) << 2) | ((uint32_t)(x & 3) << 4);
        grid[block_index * 64 + index] = (uint8_t)(i + 1);
    }
    /* collector.rs:76-82 exclusive prefix sum */
    uint32_t sum = 0;
    for (size_t b = 0; b < 64u * 64u * 64u; b++) { uint32_t v = block_counts[b]; block_counts[b] = sum; sum += v; }

    /* loader.rs:265-272 — material_ptr of every leaf */
    uint64_t n_leaves = o_tree_num_leaves(out->tree);
    uint32_t *pos = (uint32_t *)malloc(sizeof(uint32_t) * 3 * (n_leaves ? n_leaves : 1));
    uint32_t *ptr = (uint32_t *)malloc(sizeof(uint32_t) * (n_leaves ? n_leaves : 1));
    o_tree_iter_leaf(out->tree, pos, ptr, n_leaves);
    for (uint64_t l = 0; l < n_leaves; l++) {
        size_t block_index = (size_t)(pos[3 * l] >> 2) + (size_t)(pos[3 * l + 1] >> 2) * 64 + (size_t)(pos[3 * l + 2] >> 2) * 64 * 64;
        uint8_t *leaf = o_pool_get(&out->tree->pool[0], ptr[l]);
        memcpy(leaf + 16, &block_counts[block_index], 4);
    }
    /* collector.rs:47-60 — scan the grid; note a voxel listed twice is emitted once but counted twice,
     * exactly as in the reference (count is only a size hint there) */
    out->materials = (uint8_t *)malloc(count ? count : 1);
    uint64_t nm = 0;
    for (size_t c = 0; c < 256u * 256u * 256u; c++) if (grid[c]) out->materials[nm++] = (uint8_t)(grid[c] - 1);
    out->n_materials = nm;

    /* geometry.rs:68-128 */
    out->n_leaves = (uint32_t)n_leaves;
    out->nodes = (OGpuVoxNode *)calloc(n_leaves ? n_leaves : 1, sizeof(OGpuVoxNode));
    out->aabbs = (float *)calloc(n_leaves ? n_leaves : 1, 6 * sizeof(float));
    for (uint64_t l = 0; l < n_leaves; l++) {
        const uint8_t *leaf = o_pool_get(&out->tree->pool[0], ptr[l]);
        uint64_t mask; uint32_t material_ptr;
        memcpy(&mask, leaf, 8); memcpy(&material_ptr, leaf + 16, 4);
        float *bb = &out->aabbs[6 * l];
        for (int a = 0; a < 3; a++) { bb[a] = (float)pos[3 * l + a]; bb[3 + a] = 4.0f + (float)pos[3 * l + a]; }
        uint32_t col[4] = { 0, 0, 0, 0 };
        uint32_t nv = (uint32_t)__builtin_popcountll(mask);
        for (uint32_t i = 0; i < nv; i++) {
            uint8_t pi = out->materials[material_ptr + i];
            for (int c = 0; c < 4; c++) col[c] += s->palette[pi][c];
        }
        float denom = (float)nv * 255.0f;
        float cx = linear2srgb((float)col[0] / denom), cy = linear2srgb((float)col[1] / denom);
        float cz = linear2srgb((float)col[2] / denom), cw = (float)col[3] / denom;
        uint32_t r = (uint32_t)(cx * 1023.0f), g = (uint32_t)(cy * 1023.0f), b = (uint32_t)(cz * 1023.0f), a = (uint32_t)(cw * 3.0f);
        OGpuVoxNode *n = &out->nodes[l];
        n->x = (uint16_t)pos[3 * l]; n->y = (uint16_t)pos[3 * l + 1]; n->z = (uint16_t)pos[3 * l + 2]; n->w = 0;
        n->mask = mask; n->material_ptr = material_ptr;
        n->avg_albedo = (r << 22) | (g << 12) | (b << 2) | a;
    }
    free(grid); free(block_counts); free(pos); free(ptr);
    return out;
}

void o_vox_model_free(OVoxModelData *m)
{
    if (!m) return;
    o_tree_destroy(m->tree);
    free(m->materials); free(m->nodes); free(m->aabbs); free(m);
}
