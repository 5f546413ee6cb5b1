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
//   * 16^3 node  : 64 mask words + 64 per-word child bases            = 768 B  (was 16 896 B)
//   * 4^3  node  : {mask lo, mask hi, child base, -}                   = 16 B   (was 264 B)
//   * leaf       : {occupancy lo, occupancy hi, material_ptr, -}       = 16 B   (one 128-bit load)
//
// Everything a ray touches above the leaves of a 256^3 model is < 100 KB and stays in L1/L2.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace dust {

struct alignas(16) DNode16 {
    uint64_t mask[64];     // child_mask of an InternalNode<_, {4,4,4}>, bit = x<<8 | y<<4 | z
    uint32_t child[64];    // child[w] = slot of the first child whose bit lives in mask[w]
};
static_assert(sizeof(DNode16) == 768, "DNode16 layout");

// 4^3 internal node and leaf are both uint4:  .x/.y = 64-bit mask, .z = child base / material_ptr

struct alignas(16) DInstance {
    float w2o[12];                 // world -> object, row-major 3x4
    float o2w[9];                  // linear part of object -> world (normals)
    uint32_t n_levels;             // 3: (4,2,2)   4: (4,4,2,2)
    const DNode16 *n16;            // [0] = root; for 4 levels [1..] = the 16^3 nodes below it
    const uint4 *mids;             // 4^3 internal nodes
    const uint4 *leaves;
    const uint8_t *materials;
    const float4 *palette_lin;     // 256 x SRGBToLinear(rgb/255), w unused
    int32_t bmin[3], bmax[3];      // occupied bounds [bmin, bmax) in object voxels (multiples of 4): the walk is clipped to it
    uint32_t _pad[2];
};
static_assert(sizeof(DInstance) == 160, "DInstance layout");

// Binary BVH over instance world boxes (replaces the TLAS).  left < 0 => leaf, instance = ~left.
struct alignas(16) DBvhNode {
    float lo[3]; int32_t left;
    float hi[3]; int32_t right;
};

struct DSceneView {
    const DInstance *inst;
    const DBvhNode *bvh;
    uint32_t n_inst;
    uint32_t n_bvh;
    uint32_t max_levels;           // max n_levels over instances (kernel specialisation)
};

struct DCounters {                 // DUST_TRACE_STATS
    unsigned long long primary_rays, bounce_rays, shadow_rays, mask_words, child_ptrs, hits_shaded;
};

}  // namespace dust
