// Device-side (HBM) layout of an uploaded tree / scene.
//
// The host contract is the reference's own node layout (include/dust_b200.h, DustTreeDesc:
// crates/vdb InternalNode = mask + 8^f child pointers = 16 896 B for the 16^3 root, 264 B per
// 4^3 mid node).  Those nodes are sparse tables meant for point get/set; a ray walk only needs
// "is bit i set, and where is child i".  dust_tree_upload therefore re-lays the tree out once:
//
//   * children of every node are stored CONTIGUOUSLY in ascending child-index order, i.e. in
//     Tree::iter_leaf order (crates/vdb/src/node/internal.rs:306-332).  A child's slot is then
//     base + popcount(mask & below(i)): no pointer table, and a leaf's slot IS the reference's
//     gl_PrimitiveID (crates/vox/src/geometry.rs:68), so hit ids need no translation.
//   * EVERY node word, at every level, is one uint4 = {mask lo, mask hi, base, 0}: one 128-bit
//     load answers "is the bit set" and "where is the child".
//       16^3 node : 64 words (word w = bits 64w..64w+63 of child_mask, base = slot of the first
//                   child of that word)                                = 1024 B  (was 16 896 B)
//       4^3 node  : 1 word, base = slot of its first leaf               = 16 B    (was 264 B)
//       leaf      : 1 word, mask = occupancy, base = material_ptr, w = GPUVoxNode.avg_albedo (geometry.rs:86-113)   = 16 B    (was 24 B)
//
//   * per 16^3 node one extra 64-bit BLOCK mask (bit per 4x4x4 block of children): the walk asks it first
//     and leaves an empty block as one cell of 4x the edge (64 voxels under a 256^3 root).
//
// Everything a ray touches above the leaves of a 256^3 model is ~100 KB and stays in L1/L2.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#if !defined(DUST_NO_DENSE_MIDS) && !defined(DUST_TMA_TOP)
#define DUST_DENSE_MIDS 1            // walk the root level through DInstance::dense_root (below)
#endif

namespace dust {

struct alignas(16) DNode16 {
    uint4 word[64];        // child_mask of an InternalNode<_, {4,4,4}>, bit = x<<8 | y<<4 | z
};
static_assert(sizeof(DNode16) == 1024, "DNode16 layout");

struct alignas(16) DInstance {
    float w2o[12];                 // world -> object, row-major 3x4                       (3 x 128 bit)
    // occupied bounds in voxels (multiples of 4: whole leaves), as floats: the walk is clipped to them and only ever wants them as
    // floats (plane times, the clamp of the entry point) — packed integers cost 21 unpack / convert instructions per instance entered
    float bounds_lo[3];
    uint32_t n_levels;             // 3: (4,2,2)   4: (4,4,2,2)                            (128 bit with bounds_lo)
    float bounds_hi[3];
    uint32_t _pad1;                //                                                      (128 bit with bounds_hi)
    const DNode16 *n16;            // [0] = root; for 4 levels [1..] = the 16^3 nodes below it
    const uint4 *mids;             // 4^3 internal nodes                                   (128 bit with n16)
    const uint4 *leaves;
    const uint8_t *materials;
    const float4 *palette_lin;     // 256 x SRGBToLinear(rgb/255), w unused
    const uint2 *blocks;           // per 16^3 node: 64-bit block mask, bit bx<<4|by<<2|bz = "some child in that 4x4x4 block of children"
    const uchar4 *palette_raw;     // 256 x RGBA8 as uploaded (G-buffer albedo)
    float o2w[9];                  // linear part of object -> world (normals)
    float o2w_t[3];                // its translation (block centres of the far-field gather, final_gather.rchit:42)
    // dense_root: the root node's 4096 cells, dense (a derived table of 64 KB per tree, built on the device from the blob by vb_dense_root; NOT
    // part of the blob contract of dust_tree_layout / dust_tree_download).  One load answers "is there a child" AND delivers what
    // the walk wants of it, where the compact layout needs the root word, a bit test, a rank and a second, dependent load:
    //   3 levels: the child 4^3 node's word {mask lo, mask hi, first leaf slot, 0}
    //   4 levels: {block mask lo, block mask hi of the child 16^3 node, its index, 0}
    // All zero = no child (a node that exists has a non-empty mask).
    uint32_t _pad2[2];
    const uint4 *dense_root;       //                                                      (128 bit with root_blocks)
    uint2 root_blocks;             // block mask of the root node (blocks[0]): in registers from walk_enter on, no load through `blocks`
};
static_assert(sizeof(DInstance) == 208, "DInstance layout");

// Binary BVH over instance world boxes (replaces the TLAS).  A node holds the padded boxes of BOTH
// children, so one node fetch (4 x 128 bit) decides both.  child >= 0: inner node index;
// child < 0: instance ~child.
struct alignas(16) DBvhNode {
    float lo0[3]; int32_t child0;
    float hi0[3]; int32_t child1;
    float lo1[3]; int32_t _p0;
    float hi1[3]; int32_t _p1;
};
static_assert(sizeof(DBvhNode) == 64, "DBvhNode layout");

// WIDE BVH (the layout the kernels walk unless built with -DDUST_BINARY_BVH): the binary SAH tree collapsed to FOUR children per
// node, child boxes QUANTISED to 16 bits per plane on ONE power-of-two grid per scene (DSceneView::grid_org / grid_step; boxes are
// rounded outwards and widened by one step, which also covers the rounding of the decode below).  One node = 4 x 128 bit, the same
// 64 bytes as a binary node but for twice the children, and all four loads are independent:
//   word 0 / 1 / 2   x / y / z of children 0..3: lo | hi << 16      (an unused slot has lo = 0xFFFF, hi = 0: never hit)
//   word 3           child references 0..3 (>= 0 wide node index, < 0 instance ~ref)
// plane(q) = org + q * step.  A ray computes S = step / d and A = (org - o) / d - 2^23 S ONCE (the grid is the scene's), and then
// t(q) = fma(2^23 + q, S, A), where the operand 2^23 + q is built by ONE byte permute (PRMT) that also picks lo or hi by the sign
// of d: 6 PRMT + 6 FFMA + 4 min / max per child box, where the float boxes took 6 FADD + 6 FMUL + 10 min / max.
#define DUST_BVH_WIDTH 4

struct DSceneView {
    const DInstance *inst;
    const DBvhNode *bvh;           // binary: n_inst - 1 nodes of 64 bytes; wide (default): 4 x uint4 per node (none for 0 / 1 instances)
    const float4 *prev_o2w;        // 3 rows per instance: last frame's object -> world (G-buffer motion)
    const float4 *inst_boxes;      // 2 per instance: padded world box lo.xyz, hi.xyz (the boxes the BVH is built over; per-tile candidate lists test them directly)
    float root_lo[3], root_hi[3];  // padded world box of the whole scene
    float grid_org[3], grid_step[3];   // the wide BVH's quantisation grid (steps are powers of two)
    int32_t root_ref;              // >= 0 inner node, < 0 instance ~root_ref
    uint32_t n_inst;
    uint32_t id_bits;              // low bits of a candidate key that hold the instance index
};

struct DCounters {                 // DUST_TRACE_STATS
    unsigned long long primary_rays, bounce_rays, shadow_rays, mask_words, child_ptrs, hits_shaded;
    unsigned long long dbg[16];    // scheduling diagnostics (trace_core.cuh Counters::dbg)
    unsigned long long mask_words_strict;   // W by SURVEY.md 8d's letter: node words tested + mid nodes entered + leaves entered (no block masks, no re-counts)
};

}  // namespace dust
