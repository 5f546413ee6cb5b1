// Host-side mirror of the reference's `crates/vdb` (dust_vdb): the sparse voxel tree the Rust side
// builds and hands to dust_tree_upload.  Same public operations and, crucially, the same bytes in
// memory (#[repr(C)] node layouts, pooled nodes in chunks of 1024) so that a DustTreeDesc filled
// from this class is indistinguishable from one filled from a real `dust_vdb::Tree`.
//
// The reference fixes the hierarchy at compile time with const generics (`hierarchy!(4,2,2)`,
// crates/vdb/src/node/mod.rs:113-123); here it is a run-time list of fan-outs.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "../../../include/dust_b200.h"

namespace dust::vdb {

struct UVec3 { uint32_t x, y, z; };

// crates/vdb/src/bitmask.rs — LSB-first bits in u64 words
struct BitMaskRef {
    uint64_t *words;
    bool get(uint32_t i) const { return (words[i >> 6] >> (i & 63)) & 1u; }
    void set(uint32_t i, bool v) const
    {
        if (v) words[i >> 6] |= uint64_t(1) << (i & 63);
        else   words[i >> 6] &= ~(uint64_t(1) << (i & 63));
    }
};

// crates/vdb/src/pool.rs:3-17 — fixed-size items, chunked storage, LIFO free-list threaded through
// the first 4 bytes of freed items
class Pool {
public:
    Pool() = default;
    Pool(uint32_t item_bytes, uint32_t chunk_size_log2) : item_bytes_(item_bytes), chunk_log2_(chunk_size_log2) {}
    Pool(const Pool &) = delete;
    Pool &operator=(const Pool &) = delete;
    Pool(Pool &&o) noexcept { *this = std::move(o); }
    Pool &operator=(Pool &&o) noexcept
    {
        release();
        item_bytes_ = o.item_bytes_; chunk_log2_ = o.chunk_log2_; head_ = o.head_; top_ = o.top_; count_ = o.count_;
        chunks_ = std::move(o.chunks_);
        o.chunks_.clear();
        return *this;
    }
    ~Pool() { release(); }

    uint32_t alloc()   // pool.rs:57-86; every node type's Default is all-zero
    {
        uint32_t ptr;
        ++count_;
        if (head_ == UINT32_MAX) {
            ptr = top_++;
            if ((ptr >> chunk_log2_) >= chunks_.size()) {
                void *c = std::calloc(size_t(1) << chunk_log2_, item_bytes_);
                if (!c) throw std::bad_alloc();
                chunks_.push_back(static_cast<uint8_t *>(c));
            }
        } else {
            ptr = head_;
            std::memcpy(&head_, get(ptr), 4);
        }
        std::memset(get(ptr), 0, item_bytes_);
        return ptr;
    }
    void free(uint32_t ptr)   // pool.rs:87-102
    {
        --count_;
        uint8_t *p = get(ptr);
        std::memcpy(p, &head_, 4);
        std::memset(p + 4, 0, item_bytes_ - 4);
        head_ = ptr;
    }
    uint8_t *get(uint32_t ptr) const   // pool.rs:109-116
    {
        return chunks_[ptr >> chunk_log2_] + size_t(ptr & ((1u << chunk_log2_) - 1u)) * item_bytes_;
    }
    uint32_t count() const { return count_; }
    uint32_t top() const { return top_; }
    uint32_t item_bytes() const { return item_bytes_; }
    size_t num_chunks() const { return chunks_.size(); }
    const std::vector<uint8_t *> &chunks() const { return chunks_; }

private:
    void release()
    {
        for (uint8_t *c : chunks_) std::free(c);
        chunks_.clear();
    }
    uint32_t item_bytes_ = 0, chunk_log2_ = 10;
    uint32_t head_ = UINT32_MAX, top_ = 0, count_ = 0;
    std::vector<uint8_t *> chunks_;
};

// crates/vdb/src/tree.rs:7-13 + node/{leaf,internal}.rs.  LEVEL 0 = leaf, LEVEL n-1 = root.
class Tree {
public:
    static constexpr uint32_t kMaxLevels = 4;

    // `hierarchy` root-first, e.g. {4,2,2}
    explicit Tree(const std::vector<uint32_t> &hierarchy)
    {
        n_ = static_cast<uint32_t>(hierarchy.size());
        if (n_ < 1 || n_ > kMaxLevels) throw std::invalid_argument("1..4 levels");
        uint32_t ext = 0;
        for (uint32_t lv = 0; lv < n_; ++lv) {
            const uint32_t f = hierarchy[n_ - 1 - lv];
            if (3 * f < 6 || 3 * f > 12) throw std::invalid_argument("fan-out must be 2..4");   // BitMask needs >= 64 bits
            fan_[lv] = f;
            ext += f;
            ext_[lv] = ext;
            words_[lv] = (1u << (3 * f)) / 64u;
            const uint32_t raw = lv == 0 ? 16u * words_[lv] + 4u : 8u * words_[lv] + 4u * (1u << (3 * f));
            bytes_[lv] = (raw + 7u) & ~7u;
        }
        for (uint32_t lv = 0; lv + 1 < n_; ++lv) pools_[lv] = Pool(bytes_[lv], 10);   // tree.rs:36
        root_.assign(bytes_[n_ - 1], 0);
    }

    uint32_t levels() const { return n_; }
    uint32_t extent_log2() const { return ext_[n_ - 1]; }
    uint32_t fanout_log2(uint32_t level) const { return fan_[level]; }
    uint32_t node_bytes(uint32_t level) const { return bytes_[level]; }
    const Pool &pool(uint32_t level) const { return pools_[level]; }
    const uint8_t *root() const { return root_.data(); }

    // tree.rs:83-85.  value: -1 None, 0 Some(false), 1 Some(true).  Clearing below an internal node
    // is `todo!()` in the reference (internal.rs:121-124) and throws here.
    void set_value(UVec3 c, int value)
    {
        uint8_t *node = root_.data();
        for (uint32_t lv = n_ - 1; lv > 0; --lv) {
            if (value < 0) throw std::logic_error("clearing through an InternalNode is todo!() in the reference");
            const uint32_t f = fan_[lv], e = ext_[lv - 1];
            const uint32_t idx = ((c.x >> e) << (2 * f)) | ((c.y >> e) << f) | (c.z >> e);
            BitMaskRef mask{ reinterpret_cast<uint64_t *>(node) };
            uint32_t *ptrs = reinterpret_cast<uint32_t *>(node + 8u * words_[lv]);
            if (!mask.get(idx)) {
                mask.set(idx, true);
                ptrs[idx] = pools_[lv - 1].alloc();
            }
            const uint32_t m = (1u << e) - 1u;
            c = { c.x & m, c.y & m, c.z & m };
            node = pools_[lv - 1].get(ptrs[idx]);
        }
        const uint32_t f = fan_[0];
        const uint32_t idx = (c.x << (2 * f)) | (c.y << f) | c.z;   // leaf.rs:99-101
        BitMaskRef occ{ reinterpret_cast<uint64_t *>(node) }, act{ reinterpret_cast<uint64_t *>(node) + words_[0] };
        if (value >= 0) { occ.set(idx, true); act.set(idx, value != 0); }
        else occ.set(idx, false);
    }

    // tree.rs:78-80
    int get_value(UVec3 c) const { return get_from(n_ - 1, root_.data(), c, nullptr); }

    // OR a whole 64-voxel occupancy word into the leaf at `leaf_origin` (multiple of 4), marking the
    // voxels active — the bulk equivalent of 64 set_value(.., Some(true)) calls.
    void or_leaf(UVec3 leaf_origin, uint64_t occupancy)
    {
        if (!occupancy || fan_[0] != 2) return;
        uint8_t *node = root_.data();
        UVec3 c = leaf_origin;
        for (uint32_t lv = n_ - 1; lv > 0; --lv) {
            const uint32_t f = fan_[lv], e = ext_[lv - 1];
            const uint32_t idx = ((c.x >> e) << (2 * f)) | ((c.y >> e) << f) | (c.z >> e);
            BitMaskRef mask{ reinterpret_cast<uint64_t *>(node) };
            uint32_t *ptrs = reinterpret_cast<uint32_t *>(node + 8u * words_[lv]);
            if (!mask.get(idx)) { mask.set(idx, true); ptrs[idx] = pools_[lv - 1].alloc(); }
            const uint32_t m = (1u << e) - 1u;
            c = { c.x & m, c.y & m, c.z & m };
            node = pools_[lv - 1].get(ptrs[idx]);
        }
        reinterpret_cast<uint64_t *>(node)[0] |= occupancy;
        reinterpret_cast<uint64_t *>(node)[1] |= occupancy;
    }

    // Tree::iter (tree.rs:102-104): f(UVec3) for every occupied voxel, x-major / z-minor bit order
    template <class F> void for_each_voxel(F &&f) const { walk(n_ - 1, root_.data(), 0, { 0, 0, 0 }, false, f); }
    // Tree::iter_leaf (tree.rs:106-113): f(UVec3 origin, uint32 pool_ptr, uint8_t* leaf)
    template <class F> void for_each_leaf(F &&f) const { walk(n_ - 1, root_.data(), 0, { 0, 0, 0 }, true, f); }

    uint64_t count_leaves() const
    {
        uint64_t n = 0;
        for_each_leaf([&](UVec3, uint32_t, uint8_t *) { ++n; });
        return n;
    }

    // accessor.rs / tree.rs:154-167
    uint32_t meta_mask() const
    {
        uint32_t m = 0;
        for (uint32_t lv = 0; lv < n_; ++lv) m |= 1u << (ext_[lv] - 1u);
        return m;
    }
    static uint32_t lowest_common_ancestor_level(UVec3 a, UVec3 b, uint32_t mask, uint32_t root_level)
    {
        auto bits = [&](uint32_t p, uint32_t q) {
            const uint32_t diff = p ^ q;
            const uint32_t lz = diff ? static_cast<uint32_t>(__builtin_clz(diff)) : 32u;
            const uint32_t last = 1u << (31u - (lz < 31u ? lz : 31u));
            return static_cast<uint32_t>(__builtin_popcount(mask & ~(last - 1u)));
        };
        const uint32_t bx = bits(a.x, b.x), by = bits(a.y, b.y), bz = bits(a.z, b.z);
        const uint32_t m = bx < by ? (bx < bz ? bx : bz) : (by < bz ? by : bz);
        return root_level + 1u - m;
    }

    // accessor.rs:5-57
    class Accessor {
    public:
        explicit Accessor(const Tree &t) : t_(t) {}
        int get(UVec3 c)
        {
            const uint32_t root_level = t_.n_ - 1;
            const uint32_t lca = lowest_common_ancestor_level(last_, c, t_.meta_mask(), root_level);
            last_ = c;
            if (lca >= root_level) return t_.get_from(root_level, t_.root_.data(), c, ptrs_);
            const uint32_t m = (1u << t_.ext_[lca]) - 1u;
            return t_.get_from(lca, t_.pools_[lca].get(ptrs_[lca]), { c.x & m, c.y & m, c.z & m }, ptrs_);
        }
    private:
        const Tree &t_;
        uint32_t ptrs_[kMaxLevels] = { 0, 0, 0, 0 };
        UVec3 last_{ UINT32_MAX, UINT32_MAX, UINT32_MAX };
    };
    Accessor accessor() const { return Accessor(*this); }

    // What the Rust side would pass: &tree.root, tree.pool[i].chunks (tree.rs:11-12, pool.rs:14)
    void describe(DustTreeDesc &d, const uint8_t *materials, uint64_t n_materials, const uint8_t *palette) const
    {
        std::memset(&d, 0, sizeof d);
        d.n_levels = n_;
        for (uint32_t i = 0; i < n_; ++i) d.fanout_log2[i] = fan_[n_ - 1 - i];
        d.root = root_.data();
        d.root_bytes = root_.size();
        for (uint32_t lv = 0; lv + 1 < n_ && lv < 3; ++lv) {
            d.pool_chunks[lv] = reinterpret_cast<const void *const *>(pools_[lv].chunks().data());
            d.n_chunks[lv] = static_cast<uint32_t>(pools_[lv].num_chunks());
            d.item_bytes[lv] = pools_[lv].item_bytes();
            d.n_items[lv] = pools_[lv].top();
        }
        d.materials = materials;
        d.n_materials = n_materials;
        d.palette = palette;
    }

    uint8_t *leaf_ptr(uint32_t pool_ptr) const { return pools_[0].get(pool_ptr); }

private:
    int get_from(uint32_t lv, const uint8_t *node, UVec3 c, uint32_t *path) const
    {
        for (; lv > 0; --lv) {   // internal.rs:77-95
            const uint32_t f = fan_[lv], e = ext_[lv - 1];
            const uint32_t idx = ((c.x >> e) << (2 * f)) | ((c.y >> e) << f) | (c.z >> e);
            BitMaskRef mask{ reinterpret_cast<uint64_t *>(const_cast<uint8_t *>(node)) };
            if (!mask.get(idx)) return -1;
            const uint32_t child = reinterpret_cast<const uint32_t *>(node + 8u * words_[lv])[idx];
            if (path) path[lv - 1] = child;
            const uint32_t m = (1u << e) - 1u;
            c = { c.x & m, c.y & m, c.z & m };
            node = pools_[lv - 1].get(child);
        }
        const uint32_t f = fan_[0];
        const uint32_t idx = (c.x << (2 * f)) | (c.y << f) | c.z;   // leaf.rs:80-90
        const uint64_t *w = reinterpret_cast<const uint64_t *>(node);
        if (!((w[idx >> 6] >> (idx & 63)) & 1u)) return -1;
        return static_cast<int>((w[words_[0] + (idx >> 6)] >> (idx & 63)) & 1u);
    }

    template <class F>
    void walk(uint32_t lv, const uint8_t *node, uint32_t ptr, UVec3 o, bool leaves, F &f) const
    {
        if (lv == 0 && leaves) {
            if constexpr (std::is_invocable_v<F, UVec3, uint32_t, uint8_t *>) f(o, ptr, const_cast<uint8_t *>(node));
            return;
        }
        const uint32_t fo = fan_[lv], m = (1u << fo) - 1u;
        const uint64_t *mask = reinterpret_cast<const uint64_t *>(node);
        for (uint32_t w = 0; w < words_[lv]; ++w)
            for (uint64_t bits = mask[w]; bits; bits &= bits - 1) {
                const uint32_t idx = w * 64u + static_cast<uint32_t>(__builtin_ctzll(bits));
                const uint32_t x = idx >> (2 * fo), y = (idx >> fo) & m, z = idx & m;
                if (lv == 0) {
                    if constexpr (std::is_invocable_v<F, UVec3>) f(UVec3{ o.x + x, o.y + y, o.z + z });
                } else {
                    const uint32_t e = ext_[lv - 1];
                    const uint32_t child = reinterpret_cast<const uint32_t *>(node + 8u * words_[lv])[idx];
                    walk(lv - 1, pools_[lv - 1].get(child), child, UVec3{ o.x + (x << e), o.y + (y << e), o.z + (z << e) }, leaves, f);
                }
            }
    }

    uint32_t n_ = 0;
    uint32_t fan_[kMaxLevels] = {}, ext_[kMaxLevels] = {}, words_[kMaxLevels] = {}, bytes_[kMaxLevels] = {};
    std::vector<uint8_t> root_;
    Pool pools_[kMaxLevels];
};

}  // namespace dust::vdb
