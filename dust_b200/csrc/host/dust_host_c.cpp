// extern "C" façade over the host-side mirror (vdb.hpp / vox.hpp / render.hpp) so that the Python
// tests and bench.py can drive it through ctypes.  This is harness surface, NOT the drop-in
// boundary (that is include/dust_b200.h).  Nothing here computes a ray: traces go through
// libdust_b200.so.
#include "render.hpp"
#include "vox.hpp"

#include <cstdio>
#include <memory>
#include <string>

using dust::vdb::Tree;
using dust::vdb::UVec3;

namespace {
thread_local std::string g_err;

struct DhVox {
    std::unique_ptr<dust::vox::VoxFile> file;
    std::vector<dust::vox::Instance> instances;
};
struct DhModel {
    std::unique_ptr<dust::vox::VoxGeometry> geom;
    std::array<std::array<uint8_t, 4>, 256> palette;
};

// pcg hash (assets/shaders/headers/spatial_hash.glsl:105-111) as the generator's PRNG
inline uint32_t pcg(uint32_t v)
{
    uint32_t state = v * 747796405u + 2891336453u;
    uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}
}  // namespace

extern "C" {

const char *dh_last_error(void) { return g_err.c_str(); }

// ---- vdb ------------------------------------------------------------------------------------------
void *dh_tree_new(uint32_t n_levels, const uint32_t *hierarchy)
{
    try { return new Tree(std::vector<uint32_t>(hierarchy, hierarchy + n_levels)); }
    catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}
void dh_tree_free(void *t) { delete static_cast<Tree *>(t); }
int32_t dh_tree_set(void *t, uint32_t x, uint32_t y, uint32_t z, int32_t value)
{
    try { static_cast<Tree *>(t)->set_value({ x, y, z }, value); return 0; }
    catch (const std::exception &e) { g_err = e.what(); return -1; }
}
int32_t dh_tree_get(const void *t, uint32_t x, uint32_t y, uint32_t z) { return static_cast<const Tree *>(t)->get_value({ x, y, z }); }
void dh_tree_set_many(void *t, const uint32_t *xyz, uint64_t n)
{
    Tree *tr = static_cast<Tree *>(t);
    for (uint64_t i = 0; i < n; ++i) tr->set_value({ xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2] }, 1);
}
void dh_tree_or_leaves(void *t, const uint32_t *leaf_xyz, const uint64_t *masks, uint64_t n)
{
    Tree *tr = static_cast<Tree *>(t);
    for (uint64_t i = 0; i < n; ++i) tr->or_leaf({ leaf_xyz[3 * i], leaf_xyz[3 * i + 1], leaf_xyz[3 * i + 2] }, masks[i]);
}
uint64_t dh_tree_iter(const void *t, uint32_t *xyz, uint64_t cap)
{
    uint64_t n = 0;
    static_cast<const Tree *>(t)->for_each_voxel([&](UVec3 v) {
        if (n < cap) { xyz[3 * n] = v.x; xyz[3 * n + 1] = v.y; xyz[3 * n + 2] = v.z; }
        ++n;
    });
    return n;
}
uint64_t dh_tree_iter_leaf(const void *t, uint32_t *xyz, uint32_t *ptrs, uint64_t cap)
{
    uint64_t n = 0;
    static_cast<const Tree *>(t)->for_each_leaf([&](UVec3 o, uint32_t ptr, uint8_t *) {
        if (n < cap) {
            if (xyz) { xyz[3 * n] = o.x; xyz[3 * n + 1] = o.y; xyz[3 * n + 2] = o.z; }
            if (ptrs) ptrs[n] = ptr;
        }
        ++n;
    });
    return n;
}
uint32_t dh_tree_node_bytes(const void *t, uint32_t level) { return static_cast<const Tree *>(t)->node_bytes(level); }
uint32_t dh_tree_pool_count(const void *t, uint32_t level) { return static_cast<const Tree *>(t)->pool(level).count(); }
uint32_t dh_tree_pool_top(const void *t, uint32_t level) { return static_cast<const Tree *>(t)->pool(level).top(); }
uint32_t dh_tree_pool_chunks(const void *t, uint32_t level) { return static_cast<uint32_t>(static_cast<const Tree *>(t)->pool(level).num_chunks()); }
const uint8_t *dh_tree_pool_chunk(const void *t, uint32_t level, uint32_t chunk) { return static_cast<const Tree *>(t)->pool(level).chunks()[chunk]; }
const uint8_t *dh_tree_root(const void *t) { return static_cast<const Tree *>(t)->root(); }
uint32_t dh_tree_meta_mask(const void *t) { return static_cast<const Tree *>(t)->meta_mask(); }
uint32_t dh_lca_level(const uint32_t *a, const uint32_t *b, uint32_t mask, uint32_t root_level)
{
    return Tree::lowest_common_ancestor_level({ a[0], a[1], a[2] }, { b[0], b[1], b[2] }, mask, root_level);
}
// runs `n` accessor queries in order; out[i] = -1/0/1
void dh_tree_accessor_get_many(const void *t, const uint32_t *xyz, uint64_t n, int32_t *out)
{
    auto acc = static_cast<const Tree *>(t)->accessor();
    for (uint64_t i = 0; i < n; ++i) out[i] = acc.get({ xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2] });
}
void dh_tree_describe(const void *t, const uint8_t *materials, uint64_t n_materials, const uint8_t *palette, DustTreeDesc *out)
{
    static_cast<const Tree *>(t)->describe(*out, materials, n_materials, palette);
}

// Pool doc-test surface (crates/vdb/src/pool.rs:22-42)
void *dh_pool_new(uint32_t item_bytes, uint32_t chunk_log2) { return new dust::vdb::Pool(item_bytes, chunk_log2); }
void dh_pool_free_pool(void *p) { delete static_cast<dust::vdb::Pool *>(p); }
uint32_t dh_pool_alloc(void *p) { return static_cast<dust::vdb::Pool *>(p)->alloc(); }
void dh_pool_free(void *p, uint32_t i) { static_cast<dust::vdb::Pool *>(p)->free(i); }
uint32_t dh_pool_num_chunks(const void *p) { return static_cast<uint32_t>(static_cast<const dust::vdb::Pool *>(p)->num_chunks()); }

// ---- vox ------------------------------------------------------------------------------------------
void *dh_vox_parse(const uint8_t *bytes, uint64_t len)
{
    auto v = std::make_unique<DhVox>();
    v->file = dust::vox::VoxFile::parse(bytes, static_cast<size_t>(len), g_err);
    if (!v->file) return nullptr;
    v->instances = v->file->instances();
    return v.release();
}
void dh_vox_free(void *v) { delete static_cast<DhVox *>(v); }
uint32_t dh_vox_num_models(const void *v) { return static_cast<uint32_t>(static_cast<const DhVox *>(v)->file->models.size()); }
uint32_t dh_vox_num_instances(const void *v) { return static_cast<uint32_t>(static_cast<const DhVox *>(v)->instances.size()); }
void dh_vox_instance(const void *v, uint32_t i, uint32_t *model_id, float *obj2world)
{
    const auto &in = static_cast<const DhVox *>(v)->instances[i];
    *model_id = in.model_id;
    std::memcpy(obj2world, in.obj2world.m, sizeof in.obj2world.m);
}
const uint8_t *dh_vox_palette(const void *v) { return static_cast<const DhVox *>(v)->file->palette[0].data(); }
void dh_vox_model_info(const void *v, uint32_t model, uint32_t *size3, uint64_t *n_voxels)
{
    const auto &m = static_cast<const DhVox *>(v)->file->models[model];
    std::memcpy(size3, m.size, sizeof m.size);
    *n_voxels = m.voxels.size();
}
// the model's dot_vox voxels as parsed ({x, y, z, i} bytes): input of dust_tree_build_voxels
const uint8_t *dh_vox_model_voxels(const void *v, uint32_t model)
{
    const auto &m = static_cast<const DhVox *>(v)->file->models[model];
    static_assert(sizeof(dust::vox::Voxel) == 4, "Voxel is 4 bytes");
    return m.voxels.empty() ? nullptr : reinterpret_cast<const uint8_t *>(m.voxels.data());
}
void dh_to_transform(const int32_t *t, uint8_t rot, const uint32_t *size, float *out)
{
    const auto a = dust::vox::VoxFile::to_transform(t, rot, size);
    std::memcpy(out, a.m, sizeof a.m);
}
void *dh_vox_load_model(const void *v, uint32_t model)
{
    const DhVox *vx = static_cast<const DhVox *>(v);
    auto m = std::make_unique<DhModel>();
    m->geom = dust::vox::load_model(*vx->file, model, g_err);
    if (!m->geom) return nullptr;
    m->palette = vx->file->palette;
    return m.release();
}
void dh_model_free(void *m) { delete static_cast<DhModel *>(m); }
void *dh_model_tree(void *m) { return &static_cast<DhModel *>(m)->geom->tree; }
uint64_t dh_model_num_materials(const void *m) { return static_cast<const DhModel *>(m)->geom->materials.size(); }
const uint8_t *dh_model_materials(const void *m) { return static_cast<const DhModel *>(m)->geom->materials.data(); }
uint64_t dh_model_num_leaves(const void *m) { return static_cast<const DhModel *>(m)->geom->nodes.size(); }
const void *dh_model_nodes(const void *m) { return static_cast<const DhModel *>(m)->geom->nodes.data(); }
const float *dh_model_aabbs(const void *m) { return static_cast<const DhModel *>(m)->geom->aabbs.data(); }
void dh_model_size(const void *m, uint8_t *size3) { std::memcpy(size3, static_cast<const DhModel *>(m)->geom->size, 3); }
void dh_model_describe(const void *m, DustTreeDesc *out)
{
    const DhModel *mm = static_cast<const DhModel *>(m);
    mm->geom->tree.describe(*out, mm->geom->materials.data(), mm->geom->materials.size(), mm->palette[0].data());
}

// ---- synthetic 4096^3 domain (BASELINE.json config 5) ----------------------------
// hierarchy!(4,4,2,2); leaf slots are occupied by a union of random spheres (radius 8..64 voxels,
// seeded) until `fill` of the `domain_leaves`^3 leaf slots are taken; every occupied leaf gets a
// seeded non-zero 64-bit mask.  `domain_leaves` <= 1024 lets tests build a smaller domain corner.
// Material bytes: one per voxel in iter_leaf order (index = pcg of the running counter).
struct DhSynth {
    Tree tree{ std::vector<uint32_t>{ 4, 4, 2, 2 } };
    std::vector<uint8_t> materials;
    uint8_t palette[256 * 4];
    uint64_t n_leaves = 0;
};
void *dh_synth_new(uint32_t seed, double fill, uint32_t domain_leaves)
{
    if (domain_leaves == 0 || domain_leaves > 1024) { g_err = "domain_leaves must be 1..1024"; return nullptr; }
    auto s = std::make_unique<DhSynth>();
    const uint64_t slots = uint64_t(domain_leaves) * domain_leaves * domain_leaves;
    const uint64_t target = static_cast<uint64_t>(fill * static_cast<double>(slots));
    std::vector<uint64_t> taken((slots + 63) / 64, 0);
    uint32_t rng = pcg(seed ^ 0xD057u);
    auto next = [&]() { rng = pcg(rng + 0x9E3779B9u); return rng; };
    uint64_t count = 0;
    uint32_t guard = 0;
    while (count < target && guard++ < (1u << 24)) {
        const int32_t cx = int32_t(next() % domain_leaves), cy = int32_t(next() % domain_leaves), cz = int32_t(next() % domain_leaves);
        const int32_t r = 2 + int32_t(next() % 15);          // 2..16 leaves = 8..64 voxels
        for (int32_t x = cx - r; x <= cx + r && count < target; ++x)
            for (int32_t y = cy - r; y <= cy + r; ++y)
                for (int32_t z = cz - r; z <= cz + r; ++z) {
                    if (x < 0 || y < 0 || z < 0 || x >= int32_t(domain_leaves) || y >= int32_t(domain_leaves) || z >= int32_t(domain_leaves)) continue;
                    const int32_t dx = x - cx, dy = y - cy, dz = z - cz;
                    if (dx * dx + dy * dy + dz * dz > r * r) continue;
                    const uint64_t slot = (uint64_t(x) * domain_leaves + uint64_t(y)) * domain_leaves + uint64_t(z);
                    if ((taken[slot >> 6] >> (slot & 63)) & 1u) continue;
                    taken[slot >> 6] |= uint64_t(1) << (slot & 63);
                    uint64_t mask = (uint64_t(pcg(uint32_t(slot) ^ seed)) << 32) | pcg(uint32_t(slot >> 32) + uint32_t(slot) * 31u + seed);
                    if (!mask) mask = 1;
                    s->tree.or_leaf({ uint32_t(x) * 4u, uint32_t(y) * 4u, uint32_t(z) * 4u }, mask);
                    ++count;
                }
    }
    // material_ptr = running voxel count in iter_leaf order
    uint64_t voxels = 0;
    s->tree.for_each_leaf([&](UVec3, uint32_t, uint8_t *leaf) {
        uint64_t mask;
        std::memcpy(&mask, leaf, 8);
        const uint32_t ptr = static_cast<uint32_t>(voxels);
        std::memcpy(leaf + 16, &ptr, 4);
        voxels += static_cast<uint64_t>(__builtin_popcountll(mask));
        ++s->n_leaves;
    });
    if (voxels > 0xFFFFFFFFull) { g_err = "material_ptr would overflow u32"; return nullptr; }
    s->materials.resize(voxels);
    for (uint64_t i = 0; i < voxels; ++i) s->materials[i] = static_cast<uint8_t>(pcg(uint32_t(i) + seed) % 255u);
    for (uint32_t i = 0; i < 256; ++i) {
        const uint32_t h = pcg(i * 7919u + seed);
        s->palette[4 * i] = uint8_t(64 + (h & 127)); s->palette[4 * i + 1] = uint8_t(64 + ((h >> 8) & 127));
        s->palette[4 * i + 2] = uint8_t(64 + ((h >> 16) & 127)); s->palette[4 * i + 3] = 255;
    }
    return s.release();
}
void dh_synth_free(void *s) { delete static_cast<DhSynth *>(s); }
void *dh_synth_tree(void *s) { return &static_cast<DhSynth *>(s)->tree; }
uint64_t dh_synth_num_leaves(const void *s) { return static_cast<const DhSynth *>(s)->n_leaves; }
const uint8_t *dh_synth_materials(const void *s) { return static_cast<const DhSynth *>(s)->materials.data(); }
uint64_t dh_synth_num_materials(const void *s) { return static_cast<const DhSynth *>(s)->materials.size(); }
const uint8_t *dh_synth_palette(const void *s) { return static_cast<const DhSynth *>(s)->palette; }
void dh_synth_describe(const void *s, DustTreeDesc *out)
{
    const DhSynth *ss = static_cast<const DhSynth *>(s);
    ss->tree.describe(*out, ss->materials.data(), ss->materials.size(), ss->palette);
}

// ---- render ---------------------------------------------------------------------------------------
void dh_sunlight_bake(float turbidity, const float *albedo, const float *direction, DustSky *out)
{
    dust::render::Sunlight s;
    s.turbidity = turbidity;
    s.albedo = { albedo[0], albedo[1], albedo[2] };
    s.direction = { direction[0], direction[1], direction[2] };
    *out = s.bake();
}
void dh_camera_look_at(const float *eye, const float *target, const float *up, float fov, float near_, float far_,
                       uint32_t width, uint32_t height, DustCamera *out)
{
    dust::render::PinholeProjection p;
    p.fov = fov; p.near = near_; p.far = far_;
    const auto xf = dust::render::CameraTransform::looking_at({ eye[0], eye[1], eye[2] }, { target[0], target[1], target[2] }, { up[0], up[1], up[2] });
    *out = dust::render::camera_settings(p, xf, width, height);
}
// examples/castle.rs:260-265 — sun direction from the scroll state (`current`)
void dh_castle_sun_direction(float current, float *out3)
{
    const float ang = current * 0.001f - 1.57079632679489661923f;
    // Mat3A::from_axis_angle(X, ang) * (0,0,1)
    const float s1 = std::sin(ang), c1 = std::cos(ang);
    const float v[3] = { 0.0f, -s1, c1 };
    // Mat3A::from_axis_angle(Y, 0.2) * v
    const float s2 = std::sin(0.2f), c2 = std::cos(0.2f);
    out3[0] = c2 * v[0] + s2 * v[2];
    out3[1] = v[1];
    out3[2] = -s2 * v[0] + c2 * v[2];
}
// CudaPipeline::render — the mirror of StandardPipeline::render (standard.rs:228-240)
int32_t dh_pipeline_render(DustCtx *ctx, const DustScene *scene, const DustFrame *gbuffer, uint32_t width, uint32_t height,
                           float turbidity, const float *albedo, const float *sun_direction,
                           const float *eye, const float *target, const float *up, float fov, float near_, float far_,
                           uint32_t spp, uint32_t max_bounces, uint32_t rand, uint32_t frame_index, uint32_t flags,
                           uint32_t shard_index, uint32_t shard_count)
{
    dust::render::CudaPipeline pipe(ctx);
    dust::render::RenderParams rp;
    rp.sunlight.turbidity = turbidity;
    rp.sunlight.albedo = { albedo[0], albedo[1], albedo[2] };
    rp.sunlight.direction = { sun_direction[0], sun_direction[1], sun_direction[2] };
    rp.spp = spp; rp.max_bounces = max_bounces; rp.rand = rand; rp.frame_index = frame_index; rp.flags = flags;
    rp.shard_index = shard_index; rp.shard_count = shard_count;
    dust::render::PinholeProjection p;
    p.fov = fov; p.near = near_; p.far = far_;
    const auto xf = dust::render::CameraTransform::looking_at({ eye[0], eye[1], eye[2] }, { target[0], target[1], target[2] }, { up[0], up[1], up[2] });
    pipe.render(*gbuffer, width, height, scene, rp, p, xf);
    return pipe.last_status();
}

}  // extern "C"
