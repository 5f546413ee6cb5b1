/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under dust_b200/ may include, link or call this.
 *
 * CPU restatement (plain C) of the reference's `crates/vdb` sparse voxel tree: BitMask, Pool,
 * LeafNode, InternalNode, Tree, Accessor.  Every function cites the reference file:line it
 * follows (paths relative to the reference checkout, dust-engine/dust @ 7ea772b).
 *
 * Parity status: the DATA STRUCTURE is pinned by the reference's own doc-tests / unit tests
 * (tests/test_oracle_vdb.py ports them verbatim).  Byte layouts are the reference's #[repr(C)]
 * layouts on a 64-bit host, with the const-generic hierarchy replaced by run-time fan-outs.
 */
#ifndef ORACLE_VDB_H
#define ORACLE_VDB_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define O_MAX_LEVELS 4

/* crates/vdb/src/pool.rs:3-17 */
typedef struct OPool {
    uint32_t item_bytes;      /* layout.pad_to_align().size() */
    uint32_t head;            /* free-list head, u32::MAX when empty */
    uint32_t top;             /* first never-allocated slot */
    uint32_t chunk_log2;      /* items per chunk = 1 << chunk_log2 */
    uint32_t count;
    uint32_t n_chunks, cap_chunks;
    uint8_t **chunks;
} OPool;

/* crates/vdb/src/tree.rs:7-13.  LEVEL 0 = leaf, LEVEL n_levels-1 = root (owned inline). */
typedef struct OTree {
    uint32_t n_levels;
    uint32_t fanout_log2[O_MAX_LEVELS];   /* indexed by LEVEL (leaf first) */
    uint32_t extent_log2[O_MAX_LEVELS];   /* indexed by LEVEL */
    uint32_t mask_words[O_MAX_LEVELS];    /* u64 words of the node's bitmask */
    uint32_t node_bytes[O_MAX_LEVELS];    /* #[repr(C)] size of a node at LEVEL */
    uint8_t *root;                        /* node_bytes[n_levels-1] bytes */
    OPool pool[O_MAX_LEVELS];             /* pool[LEVEL] for LEVEL < n_levels-1 */
} OTree;

/* ---- BitMask (crates/vdb/src/bitmask.rs) ---- */
int      o_bitmask_get(const uint64_t *data, uint32_t index);
void     o_bitmask_set(uint64_t *data, uint32_t index, int val);
uint32_t o_bitmask_count_ones(const uint64_t *data, uint32_t n_words);
/* writes the set-bit indices in iteration order; returns how many */
uint32_t o_bitmask_iter_set_bits(const uint64_t *data, uint32_t n_words, uint32_t *out, uint32_t cap);

/* ---- Pool (crates/vdb/src/pool.rs) ---- */
void     o_pool_init(OPool *p, uint32_t item_bytes, uint32_t chunk_log2);
void     o_pool_destroy(OPool *p);
uint32_t o_pool_alloc(OPool *p);
void     o_pool_free(OPool *p, uint32_t index);
uint8_t *o_pool_get(const OPool *p, uint32_t ptr);
uint32_t o_pool_num_chunks(const OPool *p);

/* ---- Tree ---- */
/* `hierarchy` is written root-first exactly like hierarchy!(4,2,2). */
OTree   *o_tree_new(uint32_t n_levels, const uint32_t *hierarchy);
void     o_tree_destroy(OTree *t);
/* Borrowing view over a tree that already lies in memory in the reference layout (root node +
 * per-LEVEL pool chunk pointers, `tops[lv]` slots in use).  Nothing is copied; destroy with
 * o_tree_view_free.  Lets the oracle trace exactly the bytes the product was handed. */
OTree   *o_tree_view(uint32_t n_levels, const uint32_t *hierarchy, const void *root,
                     const void *const *const *pool_chunks, const uint32_t *n_chunks, const uint32_t *tops);
void     o_tree_view_free(OTree *t);
/* value: -1 = None (clear; panics in the reference above leaf level), 0 = Some(false), 1 = Some(true) */
int      o_tree_set_value(OTree *t, uint32_t x, uint32_t y, uint32_t z, int value);
/* returns -1 = None, 0 = Some(false), 1 = Some(true) */
int      o_tree_get_value(const OTree *t, uint32_t x, uint32_t y, uint32_t z);
/* Tree::iter: voxel coordinates in iteration order (xyz triples). returns count (may exceed cap) */
uint64_t o_tree_iter(const OTree *t, uint32_t *xyz_out, uint64_t cap);
/* Tree::iter_leaf: leaf origins (xyz triples) + pool pointers of the leaves, iteration order */
uint64_t o_tree_iter_leaf(const OTree *t, uint32_t *xyz_out, uint32_t *ptr_out, uint64_t cap);
uint64_t o_tree_num_leaves(const OTree *t);

/* raw access used by the tracer / tests */
const uint8_t *o_tree_root(const OTree *t);
uint32_t o_tree_node_bytes(const OTree *t, uint32_t level);
uint32_t o_tree_pool_count(const OTree *t, uint32_t level);
uint32_t o_tree_pool_top(const OTree *t, uint32_t level);
uint32_t o_tree_pool_num_chunks(const OTree *t, uint32_t level);
const uint8_t *o_tree_pool_chunk(const OTree *t, uint32_t level, uint32_t chunk);

/* ---- Accessor (crates/vdb/src/accessor.rs) ---- */
uint32_t o_meta_mask(const OTree *t);   /* same value on x, y and z */
uint32_t o_lowest_common_ancestor_level(const uint32_t a[3], const uint32_t b[3], uint32_t mask, uint32_t root_level);
typedef struct OAccessor {
    const OTree *tree;
    uint32_t ptrs[O_MAX_LEVELS];
    uint32_t last[3];
} OAccessor;
void o_accessor_init(OAccessor *a, const OTree *t);
int  o_accessor_get(OAccessor *a, uint32_t x, uint32_t y, uint32_t z);

#ifdef __cplusplus
}
#endif
#endif
