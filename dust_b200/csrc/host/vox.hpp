// Host-side mirror of the reference's `crates/vox` ingest: MagicaVoxel .vox -> per-model
// hierarchy!(4,2,2) trees, compact material bytes, GPUVoxNode records and scene-graph instances
// (crates/vox/src/{loader.rs, collector.rs, geometry.rs}).  The .vox reader stands in for the
// third-party dot_vox 5.1.1 the reference calls (loader.rs:331-332): Voxel.i = stored colour
// index - 1, palette = the 256 RGBA entries as stored, `_r` = packed signed-permutation byte.
#pragma once
#include "vdb.hpp"

#include <array>
#include <cmath>
#include <memory>
#include <string>
#include <vector>

namespace dust::vox {

using vdb::Tree;
using vdb::UVec3;

struct Voxel { uint8_t x, y, z, i; };

struct Model {
    uint32_t size[3] = { 0, 0, 0 };          // .vox axes, z up
    std::vector<Voxel> voxels;
};

struct SceneNode {
    enum Kind { Transform, Group, Shape } kind = Transform;
    int32_t child = -1;                      // Transform
    int32_t translation[3] = { 0, 0, 0 };
    uint8_t rotation = 4;                    // dot_vox Rotation::IDENTITY
    std::vector<int32_t> children;           // Group
    int32_t model_id = -1;                   // Shape
};

struct Affine { float m[12]; };              // row-major 3x4

struct Instance { uint32_t model_id; Affine obj2world; };

// GPUVoxNode, crates/vox/src/geometry.rs:40-49
struct GpuVoxNode {
    uint16_t x, y, z, w;
    uint64_t mask;
    uint32_t material_ptr;
    uint32_t avg_albedo;
};
static_assert(sizeof(GpuVoxNode) == 24, "GPUVoxNode is 24 bytes (geometry.rs:150)");

class VoxFile {
public:
    std::vector<Model> models;
    std::vector<SceneNode> scenes;
    std::array<std::array<uint8_t, 4>, 256> palette{};

    static std::unique_ptr<VoxFile> parse(const uint8_t *data, size_t len, std::string &err)
    {
        Cursor c{ data, len, 0 };
        auto f = std::make_unique<VoxFile>();
        if (len < 20 || std::memcmp(data, "VOX ", 4) != 0 || std::memcmp(data + 8, "MAIN", 4) != 0) { err = "not a .vox file"; return nullptr; }
        c.pos = 12;
        const int32_t main_content = c.i32();
        c.i32();
        c.pos += static_cast<size_t>(main_content);
        bool have_size = false, have_palette = false;
        uint32_t size[3] = { 0, 0, 0 };
        while (c.pos + 12 <= len) {
            const std::string id(reinterpret_cast<const char *>(data + c.pos), 4);
            c.pos += 4;
            const int32_t content = c.i32();
            const int32_t children = c.i32();
            if (content < 0 || children < 0 || c.pos + static_cast<size_t>(content) > len) { err = "truncated chunk " + id; return nullptr; }
            Cursor k{ data, c.pos + static_cast<size_t>(content), c.pos };
            if (id == "SIZE") {
                for (uint32_t &s : size) s = static_cast<uint32_t>(k.i32());
                have_size = true;
            } else if (id == "XYZI" && have_size) {
                const int32_t n = k.i32();
                if (n < 0 || k.pos + 4u * static_cast<size_t>(n) > k.len) { err = "truncated XYZI"; return nullptr; }
                Model m;
                std::memcpy(m.size, size, sizeof size);
                m.voxels.resize(static_cast<size_t>(n));
                for (int32_t v = 0; v < n; ++v) {
                    const uint8_t *p = data + k.pos + 4u * static_cast<size_t>(v);
                    m.voxels[static_cast<size_t>(v)] = { p[0], p[1], p[2], static_cast<uint8_t>(p[3] ? p[3] - 1 : 0) };
                }
                f->models.push_back(std::move(m));
                have_size = false;
            } else if (id == "RGBA" && content >= 1024) {
                std::memcpy(f->palette.data(), data + k.pos, 1024);
                have_palette = true;
            } else if (id == "nTRN" || id == "nGRP" || id == "nSHP") {
                const int32_t node_id = k.i32();
                if (node_id != static_cast<int32_t>(f->scenes.size())) { err = "scene node ids must be file-ordered"; return nullptr; }
                SceneNode nd;
                k.skip_dict();
                if (id == "nTRN") {
                    nd.kind = SceneNode::Transform;
                    nd.child = k.i32();
                    k.i32(); k.i32();
                    if (k.i32() != 1) { err = "multiple frames in transform node (loader.rs:104-106)"; return nullptr; }
                    const int32_t pairs = k.i32();
                    for (int32_t q = 0; q < pairs && !k.bad; ++q) {
                        const std::string key = k.str(), val = k.str();
                        if (key == "_t") std::sscanf(val.c_str(), "%d %d %d", &nd.translation[0], &nd.translation[1], &nd.translation[2]);
                        else if (key == "_r") nd.rotation = static_cast<uint8_t>(std::atoi(val.c_str()));
                    }
                } else if (id == "nGRP") {
                    nd.kind = SceneNode::Group;
                    const int32_t n = k.i32();
                    for (int32_t q = 0; q < n && !k.bad; ++q) nd.children.push_back(k.i32());
                } else {
                    nd.kind = SceneNode::Shape;
                    if (k.i32() != 1) { err = "multiple models in shape node (loader.rs:149-151)"; return nullptr; }
                    nd.model_id = k.i32();
                }
                if (k.bad) { err = "truncated scene node"; return nullptr; }
                f->scenes.push_back(std::move(nd));
            }
            c.pos += static_cast<size_t>(content);
        }
        if (!have_palette) default_palette(f->palette);
        return f;
    }

    // A file without an RGBA chunk uses MagicaVoxel's default palette (dot_vox substitutes its DEFAULT_PALETTE table).  The table of
    // the published format description is regular: colour index i = 1..215 walks a 6x6x6 cube with levels ff cc 99 66 33 00 (blue
    // fastest, then green, then red; black omitted), 216..255 are four ramps (red, green, blue, grey) over ee dd bb aa 88 77 55 44 22 11.
    // Stored like an RGBA chunk: entry k holds colour index k + 1.  Restated from the format description — dot_vox's own table is
    // not in the reference tree, so this is unverifiable here (DESIGN.md "deviations").
    static void default_palette(std::array<std::array<uint8_t, 4>, 256> &pal)
    {
        static const uint8_t ramp[10] = { 0xee, 0xdd, 0xbb, 0xaa, 0x88, 0x77, 0x55, 0x44, 0x22, 0x11 };
        for (int i = 1; i < 256; ++i) {
            uint8_t r = 0, g = 0, b = 0;
            if (i <= 215) {
                const int k = i - 1;
                b = uint8_t(0xff - 0x33 * (k % 6)); g = uint8_t(0xff - 0x33 * ((k / 6) % 6)); r = uint8_t(0xff - 0x33 * (k / 36));
            } else {
                const int k = i - 216, v = ramp[k % 10];
                if (k < 10) r = uint8_t(v); else if (k < 20) g = uint8_t(v); else if (k < 30) b = uint8_t(v); else r = g = b = uint8_t(v);
            }
            pal[static_cast<size_t>(i - 1)] = { r, g, b, 255 };
        }
        pal[255] = { 0, 0, 0, 0 };
    }

    // loader.rs:178-204 — see DESIGN.md "Scene ingest" for the signed-permutation derivation
    static Affine to_transform(const int32_t t[3], uint8_t rot, const uint32_t size[3])
    {
        int c0 = rot & 3, c1 = (rot >> 2) & 3;
        if (c0 == c1 || c0 == 3 || c1 == 3) { c0 = 0; c1 = 1; rot = 4; }
        const int c2 = 3 - c0 - c1;
        float M[3][3] = {};
        M[0][c0] = rot & 16 ? -1.f : 1.f;
        M[1][c1] = rot & 32 ? -1.f : 1.f;
        M[2][c2] = rot & 64 ? -1.f : 1.f;
        // parity of the permutation times the row signs
        const int perm_sign = ((c0 == 0 && c1 == 1) || (c0 == 1 && c1 == 2) || (c0 == 2 && c1 == 0)) ? 1 : -1;
        const float s = static_cast<float>(perm_sign) * M[0][c0] * M[1][c1] * M[2][c2];
        // L = C M C^T with C: (x,y,z) -> (x,z,-y)
        const int src[3] = { 0, 2, 1 };
        const float sg[3] = { 1.f, 1.f, -1.f };
        float L[3][3];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) L[i][j] = sg[i] * M[src[i]][src[j]] * sg[j];
        const float tr[3] = { float(t[0]), float(t[2]), -float(t[1]) };
        const float off[3] = { size[0] % 2 ? 0.5f : 0.f, size[2] % 2 ? 0.5f : 0.f, size[1] % 2 ? -0.5f : 0.f };
        const float half[3] = { float(size[0]) / 2.f, float(size[2]) / 2.f, float(size[1]) / 2.f };
        Affine a;
        for (int i = 0; i < 3; ++i) {
            const float center = s * (L[i][0] * half[0] + L[i][1] * half[1] + L[i][2] * half[2]);
            const float offset = s * (L[i][0] * off[0] + L[i][1] * off[1] + L[i][2] * off[2]);
            a.m[4 * i] = L[i][0]; a.m[4 * i + 1] = L[i][1]; a.m[4 * i + 2] = L[i][2];
            a.m[4 * i + 3] = tr[i] - center * s + offset;
        }
        return a;
    }

    // SceneGraphTraverser, loader.rs:59-176
    std::vector<Instance> instances() const
    {
        std::vector<Instance> out;
        Affine ident{ { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 } };
        if (scenes.empty()) {
            if (!models.empty() && !models[0].voxels.empty()) out.push_back({ 0, ident });
            return out;
        }
        const int32_t zero[3] = { 0, 0, 0 };
        visit(0, ident, zero, 4, out);
        return out;
    }

private:
    struct Cursor {
        const uint8_t *p; size_t len, pos; bool bad = false;
        int32_t i32()
        {
            if (pos + 4 > len) { bad = true; return 0; }
            int32_t v; std::memcpy(&v, p + pos, 4); pos += 4; return v;
        }
        std::string str()
        {
            const int32_t n = i32();
            if (bad || n < 0 || pos + static_cast<size_t>(n) > len) { bad = true; return {}; }
            std::string s(reinterpret_cast<const char *>(p + pos), static_cast<size_t>(n));
            pos += static_cast<size_t>(n);
            return s;
        }
        void skip_dict() { const int32_t n = i32(); for (int32_t q = 0; q < n && !bad; ++q) { str(); str(); } }
    };

    static Affine mul(const Affine &a, const Affine &b)
    {
        Affine r;
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) r.m[4 * i + j] = a.m[4 * i] * b.m[j] + a.m[4 * i + 1] * b.m[4 + j] + a.m[4 * i + 2] * b.m[8 + j];
            r.m[4 * i + 3] = a.m[4 * i] * b.m[3] + a.m[4 * i + 1] * b.m[7] + a.m[4 * i + 2] * b.m[11] + a.m[4 * i + 3];
        }
        return r;
    }

    void visit(int32_t node, const Affine &parent, const int32_t t[3], uint8_t rot, std::vector<Instance> &out) const
    {
        if (node < 0 || static_cast<size_t>(node) >= scenes.size()) return;
        const SceneNode &nd = scenes[static_cast<size_t>(node)];
        switch (nd.kind) {
        case SceneNode::Transform: {
            const int32_t acc[3] = { t[0] + nd.translation[0], t[1] + nd.translation[1], t[2] + nd.translation[2] };
            visit(nd.child, parent, acc, nd.rotation, out);
            break;
        }
        case SceneNode::Group: {
            const uint32_t no_size[3] = { 0, 0, 0 };
            const int32_t zero[3] = { 0, 0, 0 };
            const Affine g = mul(parent, to_transform(t, rot, no_size));
            for (int32_t ch : nd.children) visit(ch, g, zero, 4, out);
            break;
        }
        case SceneNode::Shape: {
            if (nd.model_id < 0 || static_cast<size_t>(nd.model_id) >= models.size()) return;
            const Model &m = models[static_cast<size_t>(nd.model_id)];
            if (m.voxels.empty()) return;
            out.push_back({ static_cast<uint32_t>(nd.model_id), mul(parent, to_transform(t, rot, m.size)) });
            break;
        }
        }
    }
};

// One model after VoxLoader::load_model (loader.rs:238-308) + VoxGeometry::from_tree (geometry.rs:55-128)
struct VoxGeometry {
    Tree tree{ std::vector<uint32_t>{ 4, 2, 2 } };     // crates/vox/src/lib.rs:19
    uint8_t size[3] = { 0, 0, 0 };                     // [x, z, y] (loader.rs:292)
    std::vector<uint8_t> materials;                    // PaletteMaterial data (loader.rs:274-288)
    std::vector<GpuVoxNode> nodes;                     // iter_leaf order == gl_PrimitiveID
    std::vector<float> aabbs;                          // 6 floats per leaf (geometry.rs:71-82)
};

inline std::unique_ptr<VoxGeometry> load_model(const VoxFile &file, uint32_t model_id, std::string &err)
{
    if (model_id >= file.models.size()) { err = "model id out of range"; return nullptr; }
    const Model &model = file.models[model_id];
    if (model.size[0] > 256 || model.size[1] > 256 || model.size[2] > 256) { err = "model larger than 256^3 (loader.rs:365)"; return nullptr; }
    auto g = std::make_unique<VoxGeometry>();
    g->size[0] = static_cast<uint8_t>(model.size[0]); g->size[1] = static_cast<uint8_t>(model.size[2]); g->size[2] = static_cast<uint8_t>(model.size[1]);

    // ModelIndexCollector (collector.rs:2-34): 1-based palette index per voxel, block-major
    std::vector<uint8_t> grid(size_t(256) * 256 * 256, 0);
    std::vector<uint32_t> block_counts(size_t(64) * 64 * 64, 0);
    for (const Voxel &v : model.voxels) {
        const uint32_t x = v.x, y = v.z, z = (model.size[1] - uint32_t(v.y) - 1u) & 255u;   // loader.rs:248-253
        g->tree.set_value({ x, y, z }, 1);
        const size_t block = size_t(x >> 2) + size_t(y >> 2) * 64 + size_t(z >> 2) * 4096;
        block_counts[block] += 1;
        grid[block * 64 + ((z & 3) | ((y & 3) << 2) | ((x & 3) << 4))] = static_cast<uint8_t>(v.i + 1);
    }
    // exclusive running sum (collector.rs:76-82)
    uint32_t run = 0;
    for (uint32_t &c : block_counts) { const uint32_t n = c; c = run; run += n; }
    // leaf.material_ptr (loader.rs:265-272)
    g->tree.for_each_leaf([&](UVec3 o, uint32_t, uint8_t *leaf) {
        const uint32_t ptr = block_counts[size_t(o.x >> 2) + size_t(o.y >> 2) * 64 + size_t(o.z >> 2) * 4096];
        std::memcpy(leaf + 16, &ptr, 4);
    });
    // compact scan (collector.rs:47-60)
    g->materials.reserve(model.voxels.size());
    for (uint8_t v : grid) if (v) g->materials.push_back(static_cast<uint8_t>(v - 1));

    // from_tree (geometry.rs:68-128)
    auto linear2srgb = [](float c) { return c <= 0.0031308f ? 12.92f * c : 1.055f * std::pow(c, 1.0f / 2.4f) - 0.055f; };
    g->tree.for_each_leaf([&](UVec3 o, uint32_t, uint8_t *leaf) {
        uint64_t mask; uint32_t mat;
        std::memcpy(&mask, leaf, 8); std::memcpy(&mat, leaf + 16, 4);
        for (int a = 0; a < 2; ++a) {
            g->aabbs.push_back(float(o.x) + 4.f * a); g->aabbs.push_back(float(o.y) + 4.f * a); g->aabbs.push_back(float(o.z) + 4.f * a);
        }
        uint32_t sum[4] = { 0, 0, 0, 0 };
        const uint32_t n = static_cast<uint32_t>(__builtin_popcountll(mask));
        for (uint32_t i = 0; i < n; ++i) {
            const auto &col = file.palette[g->materials[mat + i]];
            for (int c = 0; c < 4; ++c) sum[c] += col[static_cast<size_t>(c)];
        }
        const float denom = float(n) * 255.0f;
        const uint32_t r = static_cast<uint32_t>(linear2srgb(float(sum[0]) / denom) * 1023.0f);
        const uint32_t gg = static_cast<uint32_t>(linear2srgb(float(sum[1]) / denom) * 1023.0f);
        const uint32_t b = static_cast<uint32_t>(linear2srgb(float(sum[2]) / denom) * 1023.0f);
        const uint32_t a = static_cast<uint32_t>((float(sum[3]) / denom) * 3.0f);
        g->nodes.push_back({ uint16_t(o.x), uint16_t(o.y), uint16_t(o.z), 0, mask, mat, (r << 22) | (gg << 12) | (b << 2) | a });
    });
    return g;
}

// The device-side alternative to load_model + dust_tree_upload: only the parse stays on the host, the tree of the
// model is built by dust_tree_build_voxels from its voxel list (same device blob, tests/test_device_build.py).
inline DustTree *build_model_on_device(DustCtx *ctx, const VoxFile &file, uint32_t model_id, std::string &err)
{
    if (model_id >= file.models.size()) { err = "model id out of range"; return nullptr; }
    const Model &model = file.models[model_id];
    static_assert(sizeof(Voxel) == sizeof(DustVoxel), "Voxel == DustVoxel");
    DustTree *tree = nullptr;
    const int32_t rc = dust_tree_build_voxels(ctx, reinterpret_cast<const DustVoxel *>(model.voxels.data()), model.voxels.size(), model.size,
                                              file.palette[0].data(), DUST_MEM_HOST, &tree);
    if (rc != DUST_OK) { err = dust_last_error_string(ctx); return nullptr; }
    return tree;
}

}  // namespace dust::vox
