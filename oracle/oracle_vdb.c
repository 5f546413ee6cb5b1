/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_vdb.h).
 * CPU restatement of dust-engine/dust crates/vdb.  Citations are reference file:line.
 */
#include "oracle_vdb.h"
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ BitMask */

/* crates/vdb/src/bitmask.rs:44-51 — word = index/64, bit = index%64, LSB first */
int o_bitmask_get(const uint64_t *data, uint32_t index)
{
    uint32_t i = index / 64u;
    uint32_t j = index - i * 64u;
    return (int)((data[i] >> j) & 1u);
}

/* crates/vdb/src/bitmask.rs:53-62 */
void o_bitmask_set(uint64_t *data, uint32_t index, int val)
{
    uint32_t i = index / 64u;
    uint32_t j = index - i * 64u;
    if (val) data[i] |= (uint64_t)1 << j;
    else     data[i] &= ~((uint64_t)1 << j);
}

/* crates/vdb/src/bitmask.rs:76-78 */
uint32_t o_bitmask_count_ones(const uint64_t *data, uint32_t n_words)
{
    uint32_t c = 0;
    for (uint32_t i = 0; i < n_words; i++) c += (uint32_t)__builtin_popcountll(data[i]);
    return c;
}

/* crates/vdb/src/bitmask.rs:64-70,106-124 — ascending word, then ascending bit */
uint32_t o_bitmask_iter_set_bits(const uint64_t *data, uint32_t n_words, uint32_t *out, uint32_t cap)
{
    uint32_t n = 0;
    for (uint32_t i = 0; i < n_words; i++) {
        uint64_t state = data[i];
        while (state) {
            uint64_t t = state & (~state + 1u);          /* lowest set bit */
            uint32_t r = (uint32_t)__builtin_ctzll(state);
            if (n < cap) out[n] = i * 64u + r;
            n++;
            state ^= t;
        }
    }
    return n;
}

/* --------------------------------------------------------------------- Pool */

/* crates/vdb/src/pool.rs:44-53 */
void o_pool_init(OPool *p, uint32_t item_bytes, uint32_t chunk_log2)
{
    memset(p, 0, sizeof *p);
    p->item_bytes = item_bytes;
    p->head = UINT32_MAX;
    p->chunk_log2 = chunk_log2;
}

void o_pool_destroy(OPool *p)
{
    for (uint32_t i = 0; i < p->n_chunks; i++) free(p->chunks[i]);
    free(p->chunks);
    memset(p, 0, sizeof *p);
}

/* crates/vdb/src/pool.rs:109-116 */
uint8_t *o_pool_get(const OPool *p, uint32_t ptr)
{
    uint32_t chunk = ptr >> p->chunk_log2;
    uint32_t item = ptr & ((1u << p->chunk_log2) - 1u);
    return p->chunks[chunk] + (size_t)item * p->item_bytes;
}

/* crates/vdb/src/pool.rs:57-86 (alloc = alloc_uninitialized + write T::default(), all-zero for
 * every node type: leaf.rs:14 #[derive(Default)], internal.rs:37-43) */
uint32_t o_pool_alloc(OPool *p)
{
    uint32_t ptr;
    p->count += 1;
    if (p->head == UINT32_MAX) {
        uint32_t top = p->top;
        uint32_t chunk = top >> p->chunk_log2;
        if (chunk >= p->n_chunks) {
            if (p->n_chunks == p->cap_chunks) {
                p->cap_chunks = p->cap_chunks ? p->cap_chunks * 2 : 8;
                p->chunks = (uint8_t **)realloc(p->chunks, sizeof(uint8_t *) * p->cap_chunks);
            }
            p->chunks[p->n_chunks++] = (uint8_t *)calloc((size_t)1 << p->chunk_log2, p->item_bytes);
        }
        p->top += 1;
        ptr = top;
    } else {
        uint32_t next;
        memcpy(&next, o_pool_get(p, p->head), 4);
        ptr = p->head;
        p->head = next;
    }
    memset(o_pool_get(p, ptr), 0, p->item_bytes);
    return ptr;
}

/* crates/vdb/src/pool.rs:87-102 — first 4 bytes hold the next free index, the rest is zeroed */
void o_pool_free(OPool *p, uint32_t index)
{
    uint8_t *loc = o_pool_get(p, index);
    p->count -= 1;
    memcpy(loc, &p->head, 4);
    memset(loc + 4, 0, p->item_bytes - 4);
    p->head = index;
}

uint32_t o_pool_num_chunks(const OPool *p) { return p->n_chunks; }

/* --------------------------------------------------------------------- Tree */

static uint32_t round_up8(uint32_t v) { return (v + 7u) & ~7u; }

/* Node layouts: leaf.rs:13-25 {occupancy, active, material_ptr:u32}; internal.rs:22-32
 * {child_mask, child_ptrs:[u32; 8^f]} — both #[repr(C)], align 8. */
OTree *o_tree_new(uint32_t n_levels, const uint32_t *hierarchy)
{
    if (n_levels < 1 || n_levels > O_MAX_LEVELS) return NULL;
    OTree *t = (OTree *)calloc(1, sizeof *t);
    t->n_levels = n_levels;
    uint32_t ext = 0;
    for (uint32_t lv = 0; lv < n_levels; lv++) {
        uint32_t f = hierarchy[n_levels - 1 - lv];
        uint32_t grid = 1u << (3 * f);
        if (grid < 64) { free(t); return NULL; }   /* BitMask<SIZE> has SIZE/64 words (bitmask.rs:8) */
        t->fanout_log2[lv] = f;
        ext += f;
        t->extent_log2[lv] = ext;
        t->mask_words[lv] = grid / 64u;
        if (lv == 0) t->node_bytes[lv] = round_up8(2u * 8u * t->mask_words[lv] + 4u);
        else         t->node_bytes[lv] = round_up8(8u * t->mask_words[lv] + 4u * grid);
    }
    /* tree.rs:30-48 — one pool per non-root level, chunks of 1<<10 items */
    for (uint32_t lv = 0; lv + 1 < n_levels; lv++) o_pool_init(&t->pool[lv], t->node_bytes[lv], 10);
    t->root = (uint8_t *)calloc(1, t->node_bytes[n_levels - 1]);
    return t;
}

void o_tree_destroy(OTree *t)
{
    if (!t) return;
    for (uint32_t lv = 0; lv + 1 < t->n_levels; lv++) o_pool_destroy(&t->pool[lv]);
    free(t->root);
    free(t);
}

OTree *o_tree_view(uint32_t n_levels, const uint32_t *hierarchy, const void *root,
                   const void *const *const *pool_chunks, const uint32_t *n_chunks, const uint32_t *tops)
{
    OTree *t = o_tree_new(n_levels, hierarchy);
    if (!t) return NULL;
    free(t->root);
    t->root = (uint8_t *)root;
    for (uint32_t lv = 0; lv + 1 < n_levels; lv++) {
        OPool *p = &t->pool[lv];
        p->n_chunks = p->cap_chunks = n_chunks[lv];
        p->chunks = (uint8_t **)malloc(sizeof(uint8_t *) * (n_chunks[lv] ? n_chunks[lv] : 1));
        for (uint32_t c = 0; c < n_chunks[lv]; c++) p->chunks[c] = (uint8_t *)pool_chunks[lv][c];
        p->top = tops[lv];
        p->count = tops[lv];
    }
    return t;
}

void o_tree_view_free(OTree *t)
{
    if (!t) return;
    for (uint32_t lv = 0; lv + 1 < t->n_levels; lv++) free(t->pool[lv].chunks);
    free(t);
}

static uint32_t *node_ptrs(const OTree *t, uint32_t lv, uint8_t *node)
{
    return (uint32_t *)(node + 8u * t->mask_words[lv]);
}

/* leaf.rs:92-108 */
static void leaf_set(const OTree *t, uint8_t *leaf, uint32_t x, uint32_t y, uint32_t z, int value)
{
    uint32_t f = t->fanout_log2[0];
    uint32_t index = (x << (2 * f)) | (y << f) | z;
    uint64_t *occ = (uint64_t *)leaf;
    uint64_t *act = occ + t->mask_words[0];
    if (value >= 0) {
        o_bitmask_set(occ, index, 1);
        o_bitmask_set(act, index, value);
    } else {
        o_bitmask_set(occ, index, 0);
    }
}

/* leaf.rs:80-90 */
static int leaf_get(const OTree *t, const uint8_t *leaf, uint32_t x, uint32_t y, uint32_t z)
{
    uint32_t f = t->fanout_log2[0];
    uint32_t index = (x << (2 * f)) | (y << f) | z;
    const uint64_t *occ = (const uint64_t *)leaf;
    const uint64_t *act = occ + t->mask_words[0];
    if (!o_bitmask_get(occ, index)) return -1;
    return o_bitmask_get(act, index);
}

/* internal.rs:97-130 (set) — allocate the child on first touch, then recurse with masked coords */
static int node_set(OTree *t, uint32_t lv, uint8_t *node, uint32_t x, uint32_t y, uint32_t z, int value)
{
    if (lv == 0) { leaf_set(t, node, x, y, z, value); return 0; }
    uint32_t f = t->fanout_log2[lv];
    uint32_t c = t->extent_log2[lv - 1];
    uint32_t index = ((x >> c) << (2 * f)) | ((y >> c) << f) | (z >> c);
    uint64_t *mask = (uint64_t *)node;
    uint32_t *ptrs = node_ptrs(t, lv, node);
    if (value < 0) return -1;                 /* internal.rs:121-124 `todo!()` */
    if (!o_bitmask_get(mask, index)) {
        o_bitmask_set(mask, index, 1);
        uint32_t p = o_pool_alloc(&t->pool[lv - 1]);
        /* the pool may have grown, but `node` lives in another pool (or is the root) */
        ptrs[index] = p;
    }
    uint32_t m = (1u << c) - 1u;
    return node_set(t, lv - 1, o_pool_get(&t->pool[lv - 1], ptrs[index]), x & m, y & m, z & m, value);
}

/* internal.rs:77-95 (get) */
static int node_get(const OTree *t, uint32_t lv, const uint8_t *node, uint32_t x, uint32_t y, uint32_t z,
                    uint32_t *cached_path)
{
    if (lv == 0) return leaf_get(t, node, x, y, z);
    uint32_t f = t->fanout_log2[lv];
    uint32_t c = t->extent_log2[lv - 1];
    uint32_t index = ((x >> c) << (2 * f)) | ((y >> c) << f) | (z >> c);
    const uint64_t *mask = (const uint64_t *)node;
    const uint32_t *ptrs = node_ptrs(t, lv, (uint8_t *)node);
    if (!o_bitmask_get(mask, index)) return -1;
    uint32_t m = (1u << c) - 1u;
    uint32_t child = ptrs[index];
    if (cached_path) cached_path[lv - 1] = child;   /* internal.rs:139 / leaf.rs:116-118 */
    return node_get(t, lv - 1, o_pool_get(&t->pool[lv - 1], child), x & m, y & m, z & m, cached_path);
}

/* tree.rs:83-85 */
int o_tree_set_value(OTree *t, uint32_t x, uint32_t y, uint32_t z, int value)
{
    return node_set(t, t->n_levels - 1, t->root, x, y, z, value);
}

/* tree.rs:78-80 */
int o_tree_get_value(const OTree *t, uint32_t x, uint32_t y, uint32_t z)
{
    return node_get(t, t->n_levels - 1, t->root, x, y, z, NULL);
}

typedef struct {
    const OTree *t;
    uint32_t *xyz; uint32_t *ptr; uint64_t cap, n;
    int leaves;
} IterCtx;

/* internal.rs:255-290 (voxels), :306-332 (leaves); leaf.rs:213-224.  Ascending child index at
 * every level; child offset = (index >> 2f, (index >> f) & m, index & m) * CHILD::EXTENT. */
static void node_iter(IterCtx *c, uint32_t lv, const uint8_t *node, uint32_t ptr,
                      uint32_t ox, uint32_t oy, uint32_t oz)
{
    const OTree *t = c->t;
    uint32_t f = t->fanout_log2[lv];
    uint32_t m = (1u << f) - 1u;
    if (lv == 0 && c->leaves) {
        if (c->n < c->cap) {
            if (c->xyz) { c->xyz[3 * c->n] = ox; c->xyz[3 * c->n + 1] = oy; c->xyz[3 * c->n + 2] = oz; }
            if (c->ptr) c->ptr[c->n] = ptr;
        }
        c->n++;
        return;
    }
    const uint64_t *mask = (const uint64_t *)node;
    for (uint32_t w = 0; w < t->mask_words[lv]; w++) {
        uint64_t state = mask[w];
        while (state) {
            uint32_t index = w * 64u + (uint32_t)__builtin_ctzll(state);
            state &= state - 1u;
            uint32_t x = index >> (2 * f), y = (index >> f) & m, z = index & m;
            if (lv == 0) {
                if (c->n < c->cap && c->xyz) {
                    c->xyz[3 * c->n] = ox + x; c->xyz[3 * c->n + 1] = oy + y; c->xyz[3 * c->n + 2] = oz + z;
                }
                c->n++;
            } else {
                uint32_t e = t->extent_log2[lv - 1];
                uint32_t child = node_ptrs(t, lv, (uint8_t *)node)[index];
                node_iter(c, lv - 1, o_pool_get(&t->pool[lv - 1], child), child,
                          ox + (x << e), oy + (y << e), oz + (z << e));
            }
        }
    }
}

/* tree.rs:102-104 */
uint64_t o_tree_iter(const OTree *t, uint32_t *xyz_out, uint64_t cap)
{
    IterCtx c = { t, xyz_out, NULL, cap, 0, 0 };
    node_iter(&c, t->n_levels - 1, t->root, 0, 0, 0, 0);
    return c.n;
}

/* tree.rs:106-113 */
uint64_t o_tree_iter_leaf(const OTree *t, uint32_t *xyz_out, uint32_t *ptr_out, uint64_t cap)
{
    IterCtx c = { t, xyz_out, ptr_out, cap, 0, 1 };
    node_iter(&c, t->n_levels - 1, t->root, 0, 0, 0, 0);
    return c.n;
}

uint64_t o_tree_num_leaves(const OTree *t) { return o_tree_iter_leaf(t, NULL, NULL, 0); }

const uint8_t *o_tree_root(const OTree *t) { return t->root; }
uint32_t o_tree_node_bytes(const OTree *t, uint32_t level) { return t->node_bytes[level]; }
uint32_t o_tree_pool_count(const OTree *t, uint32_t level) { return t->pool[level].count; }
uint32_t o_tree_pool_top(const OTree *t, uint32_t level) { return t->pool[level].top; }
uint32_t o_tree_pool_num_chunks(const OTree *t, uint32_t level) { return t->pool[level].n_chunks; }
const uint8_t *o_tree_pool_chunk(const OTree *t, uint32_t level, uint32_t chunk)
{
    return t->pool[level].chunks[chunk];
}

/* ----------------------------------------------------------------- Accessor */

/* tree.rs:154-167 — OR over all levels of 1 << (extent_log2 - 1) */
uint32_t o_meta_mask(const OTree *t)
{
    uint32_t mask = 0;
    for (uint32_t lv = 0; lv < t->n_levels; lv++) mask |= 1u << (t->extent_log2[lv] - 1u);
    return mask;
}

static uint32_t lzcnt32(uint32_t v) { return v ? (uint32_t)__builtin_clz(v) : 32u; }
static uint32_t min_u32(uint32_t a, uint32_t b) { return a < b ? a : b; }

/* accessor.rs:15-30 */
uint32_t o_lowest_common_ancestor_level(const uint32_t a[3], const uint32_t b[3], uint32_t mask, uint32_t root_level)
{
    uint32_t pc = 32;
    for (int k = 0; k < 3; k++) {
        uint32_t diff = a[k] ^ b[k];
        uint32_t last_set_bit = 1u << (31u - min_u32(lzcnt32(diff), 31u));
        uint32_t result = mask & ~(last_set_bit - 1u);
        pc = min_u32(pc, (uint32_t)__builtin_popcount(result));
    }
    return root_level + 1u - pc;
}

/* accessor.rs:133-139 */
void o_accessor_init(OAccessor *a, const OTree *t)
{
    a->tree = t;
    memset(a->ptrs, 0, sizeof a->ptrs);
    a->last[0] = a->last[1] = a->last[2] = UINT32_MAX;
}

/* accessor.rs:37-57 — re-enter the descent at the lowest common ancestor of the previous query */
int o_accessor_get(OAccessor *a, uint32_t x, uint32_t y, uint32_t z)
{
    const OTree *t = a->tree;
    uint32_t root_level = t->n_levels - 1;
    uint32_t coords[3] = { x, y, z };
    uint32_t lca = o_lowest_common_ancestor_level(a->last, coords, o_meta_mask(t), root_level);
    a->last[0] = x; a->last[1] = y; a->last[2] = z;
    if (lca >= root_level) return node_get(t, root_level, t->root, x, y, z, a->ptrs);
    uint32_t m = (1u << t->extent_log2[lca]) - 1u;
    uint32_t ptr = a->ptrs[lca];
    return node_get(t, lca, o_pool_get(&t->pool[lca], ptr), x & m, y & m, z & m, a->ptrs);
}
