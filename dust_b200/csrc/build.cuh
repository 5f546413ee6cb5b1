// sm_100a device code, part 6: building a model's tree ON THE DEVICE from its voxel list.
// Replaces, for one .vox model, the serial host chain VoxLoader::load_model (crates/vox/src/loader.rs:238-308:
// per-voxel Tree::set_value + ModelIndexCollector, collector.rs:2-87) + VoxGeometry::from_tree + the re-layout of
// dust_tree_upload, and produces the SAME device blob section by section (tests compare them bytewise).
//
//   vb_scatter    one thread per voxel: axis swizzle (loader.rs:248-253), block_counts[block] += 1 and
//                 grid[block * 64 + bit] = max((order + 1) << 9 | (i + 1))   (the last voxel of a cell wins, like
//                 the collector's overwrite; duplicates count twice in block_counts, like the collector)
//   vb_scan_*     exclusive scans (two coalesced launches each) over the 64^3 blocks (collector order: material_ptr and material
//                 write positions; tree order: leaf slots) and the 16^3 mid nodes
//   vb_leaves     one thread per 4^3 block in TREE order (root child index, then mid child index =
//                 Tree::iter_leaf order): occupancy mask from the grid, popcount, flag
//   vb_emit_*     leaf / mid / 16^3-node words, block mask, compacted material bytes, occupied bounds
//
// Integer work throughout; every array is dense over the 256^3 domain (64^3 = 262 144 blocks), so the kernels are
// streaming passes over 128 MB (grid) + a few MB: HBM-bound, ~100 us per model instead of seconds on the host.
#pragma once
#include "device_layout.h"
#include <stdint.h>

namespace dust {

constexpr uint32_t VB_BLOCKS = 64u * 64u * 64u;      // 4^3-voxel blocks (= leaf slots) of a 256^3 model
constexpr uint32_t VB_MIDS = 16u * 16u * 16u;        // 4^3-leaf mid nodes

struct VbCounts { uint32_t n_leaves, n_mids, n_materials, max_mat_end; int32_t bmin[3], bmax[3]; };   // max_mat_end = max over leaves of material_ptr + popcount

// collector block index (x>>2) + (y>>2) * 64 + (z>>2) * 4096  <->  tree-order key (mid index << 6 | leaf-in-mid index)
__device__ __forceinline__ uint32_t vb_block_of_key(uint32_t key)
{
    const uint32_t mid = key >> 6, li = key & 63u;
    const uint32_t bx = ((mid >> 8) << 2) | (li >> 4), by = (((mid >> 4) & 15u) << 2) | ((li >> 2) & 3u), bz = ((mid & 15u) << 2) | (li & 3u);
    return bx + by * 64u + bz * 4096u;
}

__global__ void __launch_bounds__(256) vb_scatter(const uchar4 *__restrict__ voxels, uint32_t n, uint32_t size_y,
                                                  unsigned long long *__restrict__ grid, uint32_t *__restrict__ block_counts)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= n) return;
    const uchar4 v = __ldg(&voxels[i]);
    const uint32_t x = v.x, y = v.z, z = (size_y - (uint32_t)v.y - 1u) & 255u;           // loader.rs:248-253
    const uint32_t block = (x >> 2) + (y >> 2) * 64u + (z >> 2) * 4096u;                 // collector.rs:25-27
    const uint32_t bit = (z & 3u) | ((y & 3u) << 2) | ((x & 3u) << 4);                   // collector.rs:31 == leaf.rs:81-83
    atomicAdd(&block_counts[block], 1u);
    atomicMax(&grid[(size_t)block * 64u + bit], ((unsigned long long)(i + 1u) << 9) | ((unsigned long long)v.w + 1ull));
}

// occupancy of every block, in tree order
__global__ void __launch_bounds__(256) vb_leaves(const unsigned long long *__restrict__ grid, unsigned long long *__restrict__ leaf_mask,
                                                 uint32_t *__restrict__ flag_tree, uint32_t *__restrict__ pop_block)
{
    const uint32_t key = blockIdx.x * 256u + threadIdx.x;                                // < VB_BLOCKS
    const uint32_t block = vb_block_of_key(key);
    const ulonglong2 *g = reinterpret_cast<const ulonglong2 *>(grid + (size_t)block * 64u);
    unsigned long long m = 0;
#pragma unroll 8
    for (uint32_t k = 0; k < 32; k++) {
        const ulonglong2 two = __ldg(&g[k]);
        m |= (unsigned long long)(two.x != 0ull) << (2 * k);
        m |= (unsigned long long)(two.y != 0ull) << (2 * k + 1);
    }
    leaf_mask[key] = m;
    flag_tree[key] = m != 0ull;
    pop_block[block] = (uint32_t)__popcll(m);
}

// ---- exclusive scan of n = 1024 * gridDim.x uint32 in place (n <= 1024 * 1024): two launches, coalesced.
//   vb_scan_partials   CTA b: sum of its 1024 elements -> partial[b]
//   vb_scan_apply      CTA b: offset = sum of partial[0 .. b) (re-derived by every CTA: <= 1024 values), then a block-wide
//                      exclusive scan of its own 1024 elements; the last CTA also writes the grand total
__device__ __forceinline__ uint32_t vb_block_scan(uint32_t v, uint32_t *warp_sums, uint32_t &block_total)   // 1024 threads; returns the exclusive prefix
{
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int ofs = 1; ofs < 32; ofs <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, ofs); if (lane >= (unsigned)ofs) inc += t; }
    if (lane == 31u) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sums[lane];
#pragma unroll
        for (int ofs = 1; ofs < 32; ofs <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, w, ofs); if (lane >= (unsigned)ofs) w += t; }
        warp_sums[lane] = w;                                                             // inclusive over warps
    }
    __syncthreads();
    block_total = warp_sums[31];
    const uint32_t before = warp ? warp_sums[warp - 1] : 0u;
    __syncthreads();
    return before + inc - v;
}

__global__ void __launch_bounds__(1024) vb_scan_partials(const uint32_t *__restrict__ a, uint32_t *__restrict__ partial)
{
    __shared__ uint32_t ws[32];
    uint32_t total;
    vb_block_scan(a[(size_t)blockIdx.x * 1024u + threadIdx.x], ws, total);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) vb_scan_apply(uint32_t *__restrict__ a, const uint32_t *__restrict__ partial, uint32_t *__restrict__ total_out)
{
    __shared__ uint32_t ws[32];
    uint32_t ptotal, total;
    const uint32_t pv = threadIdx.x < gridDim.x ? partial[threadIdx.x] : 0u;
    const uint32_t pex = vb_block_scan(pv, ws, ptotal);                                  // exclusive prefix of the CTA sums
    __shared__ uint32_t offset;
    if (threadIdx.x == blockIdx.x) offset = pex;
    __syncthreads();
    const size_t i = (size_t)blockIdx.x * 1024u + threadIdx.x;
    const uint32_t ex = vb_block_scan(a[i], ws, total);
    a[i] = offset + ex;
    if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total_out = ptotal;
}

// per mid node (tree order): child mask from the leaf flags
__global__ void __launch_bounds__(256) vb_mid_masks(const uint32_t *__restrict__ flag_tree, unsigned long long *__restrict__ mid_mask, uint32_t *__restrict__ mid_flag)
{
    const uint32_t mid = blockIdx.x * 256u + threadIdx.x;                                // < VB_MIDS
    unsigned long long m = 0;
    for (uint32_t li = 0; li < 64; li++) m |= (unsigned long long)(flag_tree[mid * 64u + li] != 0u) << li;
    mid_mask[mid] = m;
    mid_flag[mid] = m != 0ull;
}

// leaf words {occupancy lo, hi, material_ptr, 0} at their slot; materials compacted in collector order; bounds
__global__ void __launch_bounds__(256) vb_emit_leaves(const unsigned long long *__restrict__ grid, const unsigned long long *__restrict__ leaf_mask,
                                                      const uint32_t *__restrict__ leaf_slot, const uint32_t *__restrict__ mat_ptr,
                                                      const uint32_t *__restrict__ mat_pos, uint4 *__restrict__ leaves, uint8_t *__restrict__ materials,
                                                      VbCounts *__restrict__ counts)
{
    const uint32_t key = blockIdx.x * 256u + threadIdx.x;
    const unsigned long long m = leaf_mask[key];
    if (!m) return;
    const uint32_t block = vb_block_of_key(key);
    leaves[leaf_slot[key]] = make_uint4((uint32_t)m, (uint32_t)(m >> 32), mat_ptr[block], 0u);
    // material_ptr counts duplicate voxels (collector.rs:29), the compacted array does not: the caller rejects a model whose
    // material_ptr + popcount runs past it, exactly like dust_tree_upload does for a host-built tree
    atomicMax(&counts->max_mat_end, mat_ptr[block] + (uint32_t)__popcll(m));
    uint32_t pos = mat_pos[block];
    for (unsigned long long b = m; b; b &= b - 1)                                        // ascending bit order (collector.rs:47-60)
        materials[pos++] = (uint8_t)((grid[(size_t)block * 64u + (uint32_t)(__ffsll((long long)b) - 1)] & 511ull) - 1ull);
    const int32_t lo[3] = { (int32_t)(block & 63u) * 4, (int32_t)((block >> 6) & 63u) * 4, (int32_t)(block >> 12) * 4 };
    for (int a = 0; a < 3; a++) { atomicMin(&counts->bmin[a], lo[a]); atomicMax(&counts->bmax[a], lo[a] + 4); }
}

__global__ void __launch_bounds__(256) vb_emit_mids(const unsigned long long *__restrict__ mid_mask, const uint32_t *__restrict__ mid_slot,
                                                    const uint32_t *__restrict__ leaf_slot, uint4 *__restrict__ mids)
{
    const uint32_t mid = blockIdx.x * 256u + threadIdx.x;
    const unsigned long long m = mid_mask[mid];
    if (m) mids[mid_slot[mid]] = make_uint4((uint32_t)m, (uint32_t)(m >> 32), leaf_slot[mid * 64u], 0u);   // slot of its first leaf
}

// the root 16^3 node (64 words) and its block mask: one CTA of 64 threads
__global__ void __launch_bounds__(64) vb_emit_root(const uint32_t *__restrict__ mid_flag, const uint32_t *__restrict__ mid_slot,
                                                   DNode16 *__restrict__ n16, uint2 *__restrict__ blocks)
{
    __shared__ unsigned long long words[64];
    const uint32_t w = threadIdx.x;
    unsigned long long m = 0;
    for (uint32_t b = 0; b < 64; b++) m |= (unsigned long long)(mid_flag[w * 64u + b] != 0u) << b;
    words[w] = m;
    n16->word[w] = make_uint4((uint32_t)m, (uint32_t)(m >> 32), mid_slot[w * 64u], 0u);
    __syncthreads();
    if (w == 0) {
        unsigned long long bm = 0;                                                       // same rule as dust_tree_upload
        for (uint32_t k = 0; k < 64; k++)
            for (uint32_t bz = 0; bz < 4; bz++)
                if (words[k] & (0x000F000F000F000Full << (4 * bz))) bm |= 1ull << (((k >> 4) << 4) | ((k & 3) << 2) | bz);
        blocks[0] = make_uint2((uint32_t)bm, (uint32_t)(bm >> 32));
    }
}

// The dense root table of a tree (device_layout.h DInstance::dense_root): cell i of the root gets what the walk wants of its child, or zero.
__global__ void __launch_bounds__(256) vb_dense_root(const DNode16 *n16, const uint2 *blocks, const uint4 *mids, uint32_t n_levels, uint4 *dense)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;                 // 16 CTAs: 4096 cells
    const uint4 q = n16[0].word[i >> 6];
    const uint64_t m = ((uint64_t)q.y << 32) | q.x;
    const uint32_t b = i & 63u;
    // 3 levels: w = 1 marks a cell whose whole 4x4x4 block of cells is empty (the walk leaves such a block as one cell of 4x the edge
    // and would otherwise ask the root's block mask first)
    const uint2 rb = blocks[0];
    const uint32_t blk = ((i >> 10) << 4) | (((i >> 6) & 3u) << 2) | ((i >> 2) & 3u);
    const bool block_empty = !(((((uint64_t)rb.y << 32) | rb.x) >> blk) & 1ull);
    uint4 v = make_uint4(0u, 0u, 0u, n_levels == 3 && block_empty ? 1u : 0u);
    if ((m >> b) & 1ull) {
        const uint32_t child = q.z + (uint32_t)__popcll(m << (63u - b)) - 1u;
        if (n_levels == 3) { v = mids[child]; v.w = 0u; }
        else { const uint2 cb = blocks[child]; v = make_uint4(cb.x, cb.y, child, 0u); }
    }
    dense[i] = v;
}

}  // namespace dust
