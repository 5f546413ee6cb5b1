// C ABI of libdust_b200.so (include/dust_b200.h): context, one-time tree re-layout + upload,
// instance BVH, kernel launches.  No CPU compute path exists in this library: every trace call
// launches sm_100a kernels or fails.
#include "abi_internal.h"
#include "build.cuh"

#include <cstdlib>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

using namespace dust;

int32_t dust_fail(DustCtx *ctx, int32_t code, const std::string &msg)
{
    if (ctx) ctx->err = msg;
    return code;
}

int32_t dust_ensure_scratch(DustCtx *ctx, size_t bytes)
{
    if (bytes <= ctx->scratch_bytes) return DUST_OK;
    if (ctx->scratch) { CU_TRY(ctx, cudaFree(ctx->scratch)); ctx->scratch = nullptr; ctx->scratch_bytes = 0; }
    CU_TRY(ctx, cudaMalloc(&ctx->scratch, bytes));
    ctx->scratch_bytes = bytes;
    return DUST_OK;
}

namespace {

inline int32_t fail(DustCtx *ctx, int32_t code, const std::string &msg) { return dust_fail(ctx, code, msg); }
inline int32_t ensure_scratch(DustCtx *ctx, size_t bytes) { return dust_ensure_scratch(ctx, bytes); }

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }   // 128-bit loads / stores (include/dust_b200.h "Alignment")

// ---- reading the reference's host layout (crates/vdb) ------------------------------------------

struct HostTree {
    const DustTreeDesc *d;
    // LEVEL-indexed (0 = leaf)
    uint32_t fanout[4];
    bool ok(std::string &why) const
    {
        const uint32_t nl = d->n_levels;
        if (nl != 3 && nl != 4) { why = "only hierarchy!(4,2,2) and hierarchy!(4,4,2,2) are supported"; return false; }
        static const uint32_t h3[3] = { 4, 2, 2 }, h4[4] = { 4, 4, 2, 2 };
        for (uint32_t i = 0; i < nl; i++)
            if (d->fanout_log2[i] != (nl == 3 ? h3[i] : h4[i])) { why = "unsupported fan-out"; return false; }
        if (!d->root || d->root_bytes != 16896) { why = "root must be an InternalNode<_,{4,4,4}> (16896 bytes)"; return false; }
        const uint32_t want[3] = { 24, 264, 16896 };
        for (uint32_t lv = 0; lv + 1 < nl; lv++) {
            if (d->item_bytes[lv] != want[lv]) { why = "pool item size does not match the #[repr(C)] node layout"; return false; }
            if (d->n_items[lv] && (!d->pool_chunks[lv] || (uint64_t)d->n_chunks[lv] * 1024u < d->n_items[lv])) { why = "pool chunks missing"; return false; }
        }
        if (!d->palette) { why = "palette is NULL"; return false; }
        return true;
    }
    const uint8_t *item(uint32_t level, uint32_t ptr) const
    {
        // Pool::get, crates/vdb/src/pool.rs:109-116 (chunk_size_log2 = 10, tree.rs:36)
        if (ptr >= d->n_items[level]) return nullptr;
        return static_cast<const uint8_t *>(d->pool_chunks[level][ptr >> 10]) + (size_t)(ptr & 1023u) * d->item_bytes[level];
    }
};

// InternalNode<_, {4,4,4}>: 64 mask words then 4096 u32 (internal.rs:22-32)
inline const uint64_t *n16_mask(const uint8_t *node) { return reinterpret_cast<const uint64_t *>(node); }
inline const uint32_t *n16_ptrs(const uint8_t *node) { return reinterpret_cast<const uint32_t *>(node + 512); }

struct Relayout {
    std::vector<DNode16> n16;
    std::vector<uint4> mids, leaves;
    int32_t bmin[3] = { INT32_MAX, INT32_MAX, INT32_MAX }, bmax[3] = { INT32_MIN, INT32_MIN, INT32_MIN };
    bool bad_ptr = false;
};

// Emit the children of one 16^3 node whose children are 4^3 mid nodes (LEVEL 1), in ascending bit
// order; leaves of each mid node are appended in ascending bit order => Tree::iter_leaf order.
void emit_n16_over_mids(const HostTree &ht, const uint8_t *node, int32_t ox, int32_t oy, int32_t oz, DNode16 &out, Relayout &r)
{
    const uint64_t *mask = n16_mask(node);
    const uint32_t *ptrs = n16_ptrs(node);
    for (uint32_t w = 0; w < 64; w++) {
        out.word[w] = make_uint4((uint32_t)mask[w], (uint32_t)(mask[w] >> 32), (uint32_t)r.mids.size(), 0);
        for (uint64_t m = mask[w]; m; m &= m - 1) {
            const uint32_t idx = w * 64u + (uint32_t)__builtin_ctzll(m);
            const uint8_t *mid = ht.item(1, ptrs[idx]);
            if (!mid) { r.bad_ptr = true; continue; }
            const int32_t mx = ox + (int32_t)(idx >> 8) * 16, my = oy + (int32_t)((idx >> 4) & 15) * 16, mz = oz + (int32_t)(idx & 15) * 16;
            uint64_t mm;
            memcpy(&mm, mid, 8);
            const uint32_t *mptrs = reinterpret_cast<const uint32_t *>(mid + 8);
            r.mids.push_back(make_uint4((uint32_t)mm, (uint32_t)(mm >> 32), (uint32_t)r.leaves.size(), 0));
            for (uint64_t b = mm; b; b &= b - 1) {
                const uint32_t li = (uint32_t)__builtin_ctzll(b);
                const uint8_t *leaf = ht.item(0, mptrs[li]);
                if (!leaf) { r.bad_ptr = true; continue; }
                uint64_t occ; uint32_t mat;
                memcpy(&occ, leaf, 8); memcpy(&mat, leaf + 16, 4);      // leaf.rs:13-25
                r.leaves.push_back(make_uint4((uint32_t)occ, (uint32_t)(occ >> 32), mat, 0));
                const int32_t l[3] = { mx + (int32_t)(li >> 4) * 4, my + (int32_t)((li >> 2) & 3) * 4, mz + (int32_t)(li & 3) * 4 };
                for (int a = 0; a < 3; a++) { r.bmin[a] = std::min(r.bmin[a], l[a]); r.bmax[a] = std::max(r.bmax[a], l[a] + 4); }
            }
        }
    }
}

void relayout(const HostTree &ht, Relayout &r)
{
    const uint8_t *root = static_cast<const uint8_t *>(ht.d->root);
    if (ht.d->n_levels == 3) {
        r.n16.resize(1);
        emit_n16_over_mids(ht, root, 0, 0, 0, r.n16[0], r);
        return;
    }
    // (4,4,2,2): root 16^3 over 16^3 nodes (LEVEL 2) over mids over leaves
    const uint64_t *mask = n16_mask(root);
    const uint32_t *ptrs = n16_ptrs(root);
    uint32_t n_children = 0;
    for (uint32_t w = 0; w < 64; w++) n_children += (uint32_t)__builtin_popcountll(mask[w]);
    r.n16.resize(1 + (size_t)n_children);
    uint32_t next = 1;
    for (uint32_t w = 0; w < 64; w++) {
        r.n16[0].word[w] = make_uint4((uint32_t)mask[w], (uint32_t)(mask[w] >> 32), next, 0);
        for (uint64_t m = mask[w]; m; m &= m - 1) {
            const uint32_t idx = w * 64u + (uint32_t)__builtin_ctzll(m);
            const uint8_t *child = ht.item(2, ptrs[idx]);
            DNode16 tmp;
            memset(&tmp, 0, sizeof tmp);
            if (!child) r.bad_ptr = true;
            else emit_n16_over_mids(ht, child, (int32_t)(idx >> 8) * 256, (int32_t)((idx >> 4) & 15) * 256, (int32_t)(idx & 15) * 256, tmp, r);
            r.n16[next++] = tmp;
        }
    }
}

// SRGBToLinear, assets/shaders/headers/color.glsl:1-5, applied to palette[i].rgb / 255 (hit.rchit:72-74)
inline float srgb_to_linear(float c)
{
    return c < 0.04045f ? c / 12.92f : powf(fabsf(c + 0.055f) / 1.055f, 2.4f);
}

// GPUVoxNode.avg_albedo (crates/vox/src/geometry.rs:86-113): mean palette colour of a leaf's voxels, sRGB-encoded, packed
// R10G10B10A2 — the `Block.avg_albedo` the far-field gather multiplies cached radiance with (final_gather.rchit:68-80).
// Stored in the spare fourth word of the leaf's device word.  Host arithmetic (libm powf, like the Rust `powf`) for both build paths.
inline uint32_t leaf_avg_albedo(uint64_t mask, const uint8_t *materials, const uint8_t *palette)
{
    uint32_t col[4] = { 0, 0, 0, 0 };
    const uint32_t nv = (uint32_t)__builtin_popcountll(mask);
    if (!nv) return 0;
    for (uint32_t i = 0; i < nv; i++) {
        const uint8_t *c = palette + 4 * (size_t)materials[i];
        for (int k = 0; k < 4; k++) col[k] += c[k];
    }
    const float denom = (float)nv * 255.0f;
    float e[3];
    for (int k = 0; k < 3; k++) {
        const float c = (float)col[k] / denom;
        e[k] = c <= 0.0031308f ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
    }
    const float cw = (float)col[3] / denom;
    return ((uint32_t)(e[0] * 1023.0f) << 22) | ((uint32_t)(e[1] * 1023.0f) << 12) | ((uint32_t)(e[2] * 1023.0f) << 2) | (uint32_t)(cw * 3.0f);
}

// ---- instance transforms and BVH ----------------------------------------------------------------

// gl_WorldToObjectEXT: inverse of the row-major 3x4 affine, adjugate / determinant in double,
// rounded once to f32 (DESIGN.md "Instance transforms")
bool world2obj(const float m[12], float w[12])
{
    for (int k = 0; k < 12; k++) if (!std::isfinite(m[k])) return false;
    const double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    const double A = e * i - f * h, B = c * h - b * i, C = b * f - c * e;
    const double D = f * g - d * i, E = a * i - c * g, F = c * d - a * f;
    const double G = d * h - e * g, H = b * g - a * h, I = a * e - b * d;
    const double det = a * A + b * D + c * G;
    // singular (or nearly so, relative to the matrix scale): no world -> object map exists
    const double scale = std::max({ std::fabs(a), std::fabs(b), std::fabs(c), std::fabs(d), std::fabs(e), std::fabs(f), std::fabs(g), std::fabs(h), std::fabs(i) });
    if (!(std::fabs(det) > 1e-12 * scale * scale * scale)) return false;
    const double inv[9] = { A / det, B / det, C / det, D / det, E / det, F / det, G / det, H / det, I / det };
    const double t[3] = { m[3], m[7], m[11] };
    for (int r = 0; r < 3; r++) {
        w[4 * r + 0] = (float)inv[3 * r + 0];
        w[4 * r + 1] = (float)inv[3 * r + 1];
        w[4 * r + 2] = (float)inv[3 * r + 2];
        w[4 * r + 3] = (float)(-(inv[3 * r + 0] * t[0] + inv[3 * r + 1] * t[1] + inv[3 * r + 2] * t[2]));
    }
    for (int k = 0; k < 12; k++) if (!std::isfinite(w[k])) return false;
    return true;
}

struct Box { float lo[3], hi[3]; };

Box box_union(const Box &a, const Box &b)
{
    Box r;
    for (int k = 0; k < 3; k++) { r.lo[k] = std::min(a.lo[k], b.lo[k]); r.hi[k] = std::max(a.hi[k], b.hi[k]); }
    return r;
}

inline float box_area(const Box &b)
{
    const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return 2.0f * (dx * dy + dy * dz + dz * dx);
}

// Surface-area-heuristic sweep build (exact sweep over the three sorted axes; instance counts are small and
// the build runs once per scene).  Returns the reference of the subtree (>= 0: index into `nodes`,
// < 0: ~instance) and its box; inner nodes store the boxes of both children.  The traversal's results do
// not depend on the tree (boxes only cull), only the number of nodes a ray expands does.
int build_bvh(std::vector<DBvhNode> &nodes, std::vector<uint32_t> &ids, const std::vector<Box> &boxes, uint32_t begin, uint32_t end,
              Box &out_box, int depth, int &max_depth)
{
    max_depth = std::max(max_depth, depth);
    if (end - begin == 1) {
        out_box = boxes[ids[begin]];
        return ~(int)ids[begin];
    }
    const uint32_t n = end - begin;
    uint32_t best_split = begin + n / 2;
    int best_axis = -1;
    float best_cost = INFINITY;
    std::vector<float> right_area(n);
    // SAH may peel off one box at a time; it is used only while the worse child (n - 1 boxes) could still be
    // finished by median splits inside the traversal's depth budget
    int levels_left = 0;
    while ((1u << levels_left) < n - 1) levels_left++;
    const bool sah_ok = depth + 1 + levels_left <= DUST_MAX_BVH_DEPTH;
    for (int axis = 0; axis < 3 && sah_ok; axis++) {
        std::sort(ids.begin() + begin, ids.begin() + end, [&](uint32_t x, uint32_t y) {
            const float cx = boxes[x].lo[axis] + boxes[x].hi[axis], cy = boxes[y].lo[axis] + boxes[y].hi[axis];
            return cx < cy || (cx == cy && x < y);
        });
        Box acc = boxes[ids[end - 1]];
        for (uint32_t k = n - 1; k >= 1; k--) { acc = box_union(acc, boxes[ids[begin + k]]); right_area[k] = box_area(acc); }
        acc = boxes[ids[begin]];
        for (uint32_t k = 1; k < n; k++) {                                         // split: [begin, begin + k) | [begin + k, end)
            acc = box_union(acc, boxes[ids[begin + k - 1]]);
            const float cost = box_area(acc) * (float)k + right_area[k] * (float)(n - k);
            if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = begin + k; }
        }
    }
    if (best_axis < 0) {                                                            // median split over the widest centre extent
        float clo[3] = { INFINITY, INFINITY, INFINITY }, chi[3] = { -INFINITY, -INFINITY, -INFINITY };
        for (uint32_t k = begin; k < end; k++)
            for (int a = 0; a < 3; a++) {
                const float c = boxes[ids[k]].lo[a] + boxes[ids[k]].hi[a];
                clo[a] = std::min(clo[a], c); chi[a] = std::max(chi[a], c);
            }
        best_axis = 0;
        if (chi[1] - clo[1] > chi[best_axis] - clo[best_axis]) best_axis = 1;
        if (chi[2] - clo[2] > chi[best_axis] - clo[best_axis]) best_axis = 2;
        best_split = begin + n / 2;
    }
    const int axis = best_axis;
    std::sort(ids.begin() + begin, ids.begin() + end, [&](uint32_t x, uint32_t y) {
        const float cx = boxes[x].lo[axis] + boxes[x].hi[axis], cy = boxes[y].lo[axis] + boxes[y].hi[axis];
        return cx < cy || (cx == cy && x < y);
    });
    const uint32_t mid = best_split;
    const int me = (int)nodes.size();
    nodes.emplace_back();
    Box bl, br;
    const int l = build_bvh(nodes, ids, boxes, begin, mid, bl, depth + 1, max_depth);
    const int r = build_bvh(nodes, ids, boxes, mid, end, br, depth + 1, max_depth);
    DBvhNode &nd = nodes[me];
    for (int a = 0; a < 3; a++) { nd.lo0[a] = bl.lo[a]; nd.hi0[a] = bl.hi[a]; nd.lo1[a] = br.lo[a]; nd.hi1[a] = br.hi[a]; }
    nd.child0 = l; nd.child1 = r; nd._p0 = 0; nd._p1 = 0;
    out_box = box_union(bl, br);
    return me;
}

// The wide BVH's quantisation grid (device_layout.h "WIDE BVH"): per axis a power-of-two step that spreads the scene box over at most
// 65 000 of the 65 536 positions and is at least 4 ulp of the coordinates (so the float origin's own rounding stays below 1/8 step).
struct BvhGrid { float org[3]; double step[3]; };

BvhGrid bvh_grid(const Box &root)
{
    BvhGrid g;
    for (int a = 0; a < 3; a++) {
        const double maxabs = std::max({ std::fabs((double)root.lo[a]), std::fabs((double)root.hi[a]), 1e-30 });
        const double pad = 4e-7 * maxabs;
        const double lo = (double)root.lo[a] - pad, hi = (double)root.hi[a] + pad;
        int e = (int)std::ceil(std::log2(std::max(hi - lo, 1e-300) / 65000.0));
        e = std::min(std::max({ e, std::ilogb(maxabs) - 21, -100 }), 100);
        g.step[a] = std::ldexp(1.0, e);
        float o = (float)(lo - 2.0 * g.step[a]);
        while ((double)o > lo - g.step[a]) o = std::nextafterf(o, -INFINITY);
        g.org[a] = o;
    }
    return g;
}

// Collapse the binary SAH tree into the wide, quantised layout the kernels walk.  A wide node takes the shallowest inner child first
// (the larger box among those), so every path absorbs two binary levels per wide node where it has them and the wide tree is at most
// half as deep as the binary one.  Child boxes are quantised in double, outwards, and widened by one grid step: the decode on the
// device (trace_core.cuh cull_plane) is off by at most half a step plus a few 2^-24 relative roundings of (org - o), which the
// 4e-7 * |coordinate| term covers for ray origins within about the scene's own distance of it.
int build_wide(std::vector<uint4> &out, const std::vector<DBvhNode> &nodes, const BvhGrid &g, int ref, int depth, int &max_depth, bool &ok)
{
    struct Child { int ref; Box box; int depth; };
    max_depth = std::max(max_depth, depth);
    std::vector<Child> ch;
    auto push_kids = [&](int b, int d) {
        const DBvhNode &n = nodes[(size_t)b];
        Child c0{ n.child0, Box{ { n.lo0[0], n.lo0[1], n.lo0[2] }, { n.hi0[0], n.hi0[1], n.hi0[2] } }, d };
        Child c1{ n.child1, Box{ { n.lo1[0], n.lo1[1], n.lo1[2] }, { n.hi1[0], n.hi1[1], n.hi1[2] } }, d };
        ch.push_back(c0); ch.push_back(c1);
    };
    push_kids(ref, 1);
    while (ch.size() < (size_t)DUST_BVH_WIDTH) {
        int best = -1;
        for (int k = 0; k < (int)ch.size(); k++) {
            if (ch[k].ref < 0) continue;
            if (best < 0 || ch[k].depth < ch[best].depth || (ch[k].depth == ch[best].depth && box_area(ch[k].box) > box_area(ch[best].box))) best = k;
        }
        if (best < 0) break;
        const Child c = ch[best];
        ch.erase(ch.begin() + best);
        push_kids(c.ref, c.depth + 1);
    }
    const size_t me = out.size() / 4;
    out.resize(out.size() + 4, make_uint4(0, 0, 0, 0));
    uint32_t q[3][DUST_BVH_WIDTH], refs[DUST_BVH_WIDTH];
    for (int k = 0; k < DUST_BVH_WIDTH; k++) { q[0][k] = q[1][k] = q[2][k] = 0x0000FFFFu; refs[k] = 0; }   // unused slot: lo = 0xFFFF, hi = 0
    for (size_t k = 0; k < ch.size(); k++)
        for (int a = 0; a < 3; a++) {
            const double maxabs = std::max(std::fabs((double)ch[k].box.lo[a]), std::fabs((double)ch[k].box.hi[a]));
            const double pad = 4e-7 * maxabs;
            const double ql = std::floor(((double)ch[k].box.lo[a] - pad - (double)g.org[a]) / g.step[a]) - 1.0;
            const double qh = std::ceil(((double)ch[k].box.hi[a] + pad - (double)g.org[a]) / g.step[a]) + 1.0;
            if (!(ql >= 0.0) || !(qh <= 65535.0)) { ok = false; return 0; }       // cannot happen by construction (checked, not assumed)
            q[a][k] = (uint32_t)ql | ((uint32_t)qh << 16);
        }
    for (size_t k = 0; k < ch.size(); k++) {
        int r = ch[k].ref;
        if (r >= 0) { r = build_wide(out, nodes, g, r, depth + 1, max_depth, ok); if (!ok) return 0; }
        refs[k] = (uint32_t)r;
    }
    uint4 *node = out.data() + me * 4;                                           // (after the recursion: `out` may have moved)
    for (int a = 0; a < 3; a++) node[a] = make_uint4(q[a][0], q[a][1], q[a][2], q[a][3]);
    node[3] = make_uint4(refs[0], refs[1], refs[2], refs[3]);
    return (int)me;
}

// sun_radiance (assets/shaders/headers/sky.glsl:80-113) on the host, for the uniform-only NEE term
void sun_radiance_host(const DustSky &sky, float ar2, const float dir[3], float out[3])
{
    out[0] = out[1] = out[2] = 0.0f;
    const float cos_gamma = (dir[0] * sky.direction[0] + dir[1] * sky.direction[1]) + dir[2] * sky.direction[2];
    if (cos_gamma < 0.0f || dir[1] < 0.0f) return;
    const float singamma = 1.0f - (cos_gamma * cos_gamma);
    const float sc2 = 1.0f - ar2 * singamma * singamma;
    if (sc2 <= 0.0f) return;
    const float sc = sqrtf(sc2);
    const DustSkyChannel *ch = &sky.r;
    float xyz[3];
    for (int i = 0; i < 3; i++) {
        float dk = ch[i].ld_coefficient0;
        dk += ch[i].ld_coefficient1 * sc;
        float cur = sc;
        for (int k = 0; k < 4; k++) { cur *= sc; dk += ch[i].ld_coefficient2[k] * cur; }
        xyz[i] = sky.sunlight_intensity[i] * dk;
    }
    // XYZ2ACEScg, headers/color.glsl:24-31 (GLSL mat3 columns)
    static const float m[9] = { 1.6410228f, -0.66366285f, 0.011721907f, -0.32480323f, 1.6153315f, -0.0082844375f, -0.23642465f, 0.016756356f, 0.9883947f };
    for (int r = 0; r < 3; r++) out[r] = (m[r] * xyz[0] + m[3 + r] * xyz[1]) + m[6 + r] * xyz[2];
}

}  // namespace

// uniform-only terms hoisted out of the shaders (ambient_occlusion.rgen:45, nee.rmiss:18-19, sky.glsl:88-89): normalize(sky.direction),
// 1 / sin^2(solar radius), and sun_radiance(normalize(sun_dir)) * (1 - cos(solar radius)), which does not depend on the hit
void dust_sun_terms(const DustSky *sky, float sun_dir[3], float *sun_ar2, float sun_nee[3])
{
    const float *d = sky->direction;
    const float l = sqrtf((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
    sun_dir[0] = d[0] / l; sun_dir[1] = d[1] / l; sun_dir[2] = d[2] / l;
    const float sun_cone = 1.0f - cosf(sky->sunlight_intensity[3]);
    const float sn = sinf(sky->sunlight_intensity[3]);
    *sun_ar2 = 1.0f / (sn * sn);
    const float sl = sqrtf((sun_dir[0] * sun_dir[0] + sun_dir[1] * sun_dir[1]) + sun_dir[2] * sun_dir[2]);
    const float sdn[3] = { sun_dir[0] / sl, sun_dir[1] / sl, sun_dir[2] / sl };
    float su[3];
    sun_radiance_host(*sky, *sun_ar2, sdn, su);
    for (int a = 0; a < 3; a++) sun_nee[a] = su[a] * sun_cone;
}

namespace {

constexpr size_t kMaxLaunchRecs = 1024;                              // a 64-bounce wavefront frame issues ~7 launches per depth

// called after every kernel launch; with DUST_TRACE_TIMING it also drops an event behind the launch (launches of one call are
// back to back on one stream, so consecutive events bracket one kernel) and, in a STATS pass, snapshots the counters
void count_launch(DustCtx *ctx, const char *name = nullptr) { dust_count_launch(ctx, name); }

}  // namespace

void dust_count_launch(DustCtx *ctx, const char *name)
{
    ctx->stats.kernel_launches++;
    if (!ctx->timing_on || !name) return;
    if (ctx->launch_recs.size() >= kMaxLaunchRecs) { ctx->launch_dropped++; return; }
    cudaEvent_t ev = nullptr;
    if (ctx->launch_recs.size() < ctx->ev_pool.size()) ev = ctx->ev_pool[ctx->launch_recs.size()];
    else { if (cudaEventCreate(&ev) != cudaSuccess) return; ctx->ev_pool.push_back(ev); }
    if (ctx->snaps_valid) cudaMemcpyAsync(&ctx->h_snaps[ctx->launch_recs.size()], ctx->d_counters, sizeof(DCounters), cudaMemcpyDeviceToHost, ctx->stream);
    cudaEventRecord(ev, ctx->stream);
    ctx->launch_recs.push_back({ name, ev });
}

namespace {

int32_t fetch_counters(DustCtx *ctx)
{
    DCounters h;
    CU_TRY(ctx, cudaMemcpyAsync(&h, ctx->d_counters, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.primary_rays = h.primary_rays; ctx->stats.bounce_rays = h.bounce_rays; ctx->stats.shadow_rays = h.shadow_rays;
    ctx->stats.mask_words = h.mask_words; ctx->stats.child_ptrs = h.child_ptrs; ctx->stats.hits_shaded = h.hits_shaded;
    ctx->stats.mask_words_strict = h.mask_words_strict;
    if (std::getenv("DUST_DEBUG_STATS")) {
        fprintf(stderr, "[dust dbg] rays p/b/s %llu %llu %llu |", h.primary_rays, h.bounce_rays, h.shadow_rays);
        for (int i = 0; i < 16; i++) fprintf(stderr, " %llu", h.dbg[i]);
        fprintf(stderr, "\n");
    }
    return DUST_OK;
}

// ---- wavefront pipeline (wavefront.cuh) ---------------------------------------------------------

constexpr uint32_t kWfMaxDepths = 65;                                  // max_bounces <= 64
// per depth: 4 queue counters, 2 overflow counters, 2 x 64-bit work counters
constexpr size_t kWfCounterWords = (size_t)kWfMaxDepths * (4 + 2 + 4);

int32_t trace_wavefront(DustCtx *ctx, const FrameParams &fp, uint32_t tiles_per_shard, int nl, bool stats)
{
    const DustTraceConfig &cfg = fp.cfg;
    if (cfg.max_bounces + 1 > kWfMaxDepths) return fail(ctx, DUST_ERR_UNSUPPORTED, "wavefront pipeline: more than 64 bounces");
    const uint32_t n_slots = tiles_per_shard * 1024u;
    uint64_t max_paths = 16u << 20;
    if (const char *e = std::getenv("DUST_WF_MAX_PATHS")) max_paths = std::max<uint64_t>(1024, strtoull(e, nullptr, 10));   // tests: force several batches
    const uint32_t batch_spp = (uint32_t)std::min<uint64_t>(cfg.spp, std::max<uint64_t>(1, max_paths / n_slots));
    const uint64_t n_paths = (uint64_t)n_slots * batch_spp;
    if (n_paths >= (1ull << 31)) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "frame too large for one wavefront batch");

    // carve the scratch: [ps0 ps1 | 3 queues (a, b, path) | hits | acc | vis | cand | overflow]
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_ps0 = take(n_paths * 16), o_ps1 = take(n_paths * 16);
    size_t o_q[3][3];
    for (int q = 0; q < 3; q++) { o_q[q][0] = take(n_paths * 16); o_q[q][1] = take(n_paths * 16); o_q[q][2] = take(n_paths * 4); }
    const size_t o_hits = take(n_paths * 16), o_acc = take((size_t)n_slots * 16), o_vis = take(n_paths);
    const size_t o_cand = take(n_paths * 4 * WF_CANDS), o_ovf = take(n_paths * 4);
    if (off > ctx->wf_scratch_bytes) {
        if (ctx->wf_scratch) { CU_TRY(ctx, cudaStreamSynchronize(ctx->stream)); CU_TRY(ctx, cudaFree(ctx->wf_scratch)); ctx->wf_scratch = nullptr; ctx->wf_scratch_bytes = 0; }
        CU_TRY(ctx, cudaMalloc(&ctx->wf_scratch, off));
        ctx->wf_scratch_bytes = off;
    }
    if (!ctx->d_wf_counters) CU_TRY(ctx, cudaMalloc(&ctx->d_wf_counters, kWfCounterWords * sizeof(uint32_t)));
    uint8_t *base = static_cast<uint8_t *>(ctx->wf_scratch);

    WfParams p;
    memset(&p, 0, sizeof p);
    p.f = fp;
    p.n_slots = n_slots;
    p.ps0 = reinterpret_cast<float4 *>(base + o_ps0); p.ps1 = reinterpret_cast<float4 *>(base + o_ps1);
    WfQueue *qs[3] = { &p.qc, &p.qs, &p.qn };
    for (int q = 0; q < 3; q++) {
        qs[q]->a = reinterpret_cast<float4 *>(base + o_q[q][0]); qs[q]->b = reinterpret_cast<float4 *>(base + o_q[q][1]);
        qs[q]->path = reinterpret_cast<uint32_t *>(base + o_q[q][2]);
        qs[q]->stride = 1;
        qs[q]->kind = 0; qs[q]->utmin = 0.1f; qs[q]->ub = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
#ifndef DUST_TMA_QUEUE
    // compact records: the path word rides in the origin's fourth lane; tmin (0.1: ambient_occlusion.rgen:44, final_gather.rgen:47) is a uniform of
    // the queue, and so are the direction and tmax of the sun shadow rays.  (The TMA staging experiment keeps the separate arrays it copies.)
    p.qc.kind = 2; p.qn.kind = 2;
    p.qs.kind = 1; p.qs.ub = make_float4(fp.sun_dir[0], fp.sun_dir[1], fp.sun_dir[2], 10000.0f);
#endif
    const WfQueue q_even = p.qc, q_odd = p.qn;                         // closest queues of even / odd depths
    p.hits = reinterpret_cast<uint4 *>(base + o_hits); p.acc = reinterpret_cast<float4 *>(base + o_acc);
    p.vis = base + o_vis;
    p.cand = reinterpret_cast<uint4 *>(base + o_cand); p.overflow = reinterpret_cast<uint32_t *>(base + o_ovf);
    p.counts = ctx->d_wf_counters;
    uint32_t *ovf_counts = p.counts + 4 * kWfMaxDepths;                                          // [2 * depth + (shadow ? 1 : 0)]
    unsigned long long *work = reinterpret_cast<unsigned long long *>(ovf_counts + 2 * kWfMaxDepths);   // 8-byte aligned: all counts above are even
    // queues of incoherent rays (bounce rays and the shadow rays behind them) are pre-culled and traced by persistent lanes
    uint32_t pt_depth = 1;
    if (const char *e = std::getenv("DUST_WF_PT_DEPTH")) pt_depth = (uint32_t)atoi(e);           // tuning / A-B only
    int pt_blocks = 0;
    const bool fuse0 = std::getenv("DUST_WF_FUSE0") != nullptr;                                  // A/B only: depth 0 as ONE kernel (measured 2.8 % slower)

    auto trace = [&](const WfQueue &q, const uint32_t *count, const uint32_t *list, uint32_t grid) {
        if (nl == 3) { if (stats) wf_trace<3, true><<<grid, DUST_THREADS, 0, ctx->stream>>>(p, q, count, list); else wf_trace<3, false><<<grid, DUST_THREADS, 0, ctx->stream>>>(p, q, count, list); }
        else if (nl == 4) { if (stats) wf_trace<4, true><<<grid, DUST_THREADS, 0, ctx->stream>>>(p, q, count, list); else wf_trace<4, false><<<grid, DUST_THREADS, 0, ctx->stream>>>(p, q, count, list); }
        else { if (stats) wf_trace<0, true><<<grid, DUST_THREADS, 0, ctx->stream>>>(p, q, count, list); else wf_trace<0, false><<<grid, DUST_THREADS, 0, ctx->stream>>>(p, q, count, list); }
        count_launch(ctx, list ? "wf_trace (overflow)" : "wf_trace");
    };
    cudaError_t pt_err = cudaSuccess;
    auto trace_pt = [&](const WfQueue &q, const uint32_t *count, uint32_t slot, uint32_t grid) {
        if (stats) wf_cull<true><<<grid, DUST_THREADS, 0, ctx->stream>>>(p, q, count, ovf_counts + slot);
        else wf_cull<false><<<grid, DUST_THREADS, 0, ctx->stream>>>(p, q, count, ovf_counts + slot);
        count_launch(ctx, "wf_cull");
        auto launch = [&](auto kernel) {
            if (!pt_blocks) pt_err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pt_blocks, kernel, DUST_THREADS, 0);
            const uint32_t pgrid = std::min<uint32_t>(grid, (uint32_t)(ctx->sms * std::max(pt_blocks, 1)));
            kernel<<<pgrid, DUST_THREADS, 0, ctx->stream>>>(p, q, count, work + slot, ovf_counts + slot);
        };
#define DUST_PT_BY_KIND(NL_, STATS_)                                                                                         \
        do {                                                                                                                 \
            if (q.kind == 2u) launch(wf_trace_pt<NL_, STATS_, true, false, 2>);                                              \
            else if (q.kind == 1u) launch(wf_trace_pt<NL_, STATS_, true, false, 1>);                                         \
            else launch(wf_trace_pt<NL_, STATS_, true, false, 0>);                                                           \
        } while (0)
        if (nl == 3) { if (stats) DUST_PT_BY_KIND(3, true); else DUST_PT_BY_KIND(3, false); }
        else if (nl == 4) { if (stats) DUST_PT_BY_KIND(4, true); else DUST_PT_BY_KIND(4, false); }
        else { if (stats) DUST_PT_BY_KIND(0, true); else DUST_PT_BY_KIND(0, false); }
#undef DUST_PT_BY_KIND
        count_launch(ctx, "wf_trace_pt");
        // rays that wf_cull could not pre-cull: one CTA per ray, persistent grid (the list length lives on the device)
        const uint32_t sgrid = std::min<uint32_t>(grid, (uint32_t)ctx->sms * 8u);
        if (nl == 3) { if (stats) wf_trace_slow<3, true><<<sgrid, DUST_THREADS, 0, ctx->stream>>>(p, q, ovf_counts + slot, p.overflow); else wf_trace_slow<3, false><<<sgrid, DUST_THREADS, 0, ctx->stream>>>(p, q, ovf_counts + slot, p.overflow); }
        else if (nl == 4) { if (stats) wf_trace_slow<4, true><<<sgrid, DUST_THREADS, 0, ctx->stream>>>(p, q, ovf_counts + slot, p.overflow); else wf_trace_slow<4, false><<<sgrid, DUST_THREADS, 0, ctx->stream>>>(p, q, ovf_counts + slot, p.overflow); }
        else { if (stats) wf_trace_slow<0, true><<<sgrid, DUST_THREADS, 0, ctx->stream>>>(p, q, ovf_counts + slot, p.overflow); else wf_trace_slow<0, false><<<sgrid, DUST_THREADS, 0, ctx->stream>>>(p, q, ovf_counts + slot, p.overflow); }
        count_launch(ctx, "wf_trace_slow");
    };

    for (uint32_t first = 0; first < cfg.spp; first += batch_spp) {
        p.batch_first = first;
        p.batch_spp = std::min(batch_spp, cfg.spp - first);
        p.n_paths = n_slots * p.batch_spp;
        p.last_batch = first + p.batch_spp >= cfg.spp;
        const uint32_t grid128 = (p.n_paths + DUST_THREADS - 1) / DUST_THREADS, grid256 = (p.n_paths + 255) / 256;
        CU_TRY(ctx, cudaMemsetAsync(ctx->d_wf_counters, 0, kWfCounterWords * sizeof(uint32_t), ctx->stream));
        p.qc = q_even; p.qn = q_odd;
        // the depth-0 queue is implicit (slot i = camera ray of path i, recomputed by its consumers) unless it goes to the persistent tracer
        const bool implicit0 = pt_depth > 0;
        if (!implicit0) { wf_gen_primary<<<grid256, 256, 0, ctx->stream>>>(p); count_launch(ctx, "wf_gen_primary"); }
        for (uint32_t depth = 0; depth <= cfg.max_bounces; depth++) {
            p.depth = depth;
            p.qc = depth & 1 ? q_odd : q_even; p.qn = depth & 1 ? q_even : q_odd;
            p.qc.utmin = depth == 0 ? fp.cam.near : 0.1f;                  // (a materialised depth-0 queue holds camera rays: tmin = near)
            if (depth == 0 && implicit0) { p.qc.a = nullptr; p.qc.b = nullptr; p.qc.path = nullptr; p.qc.stride = 0; p.qc.kind = 0; }
            // queue sizes live on the device: launch for the largest possible queue, surplus CTAs exit at once
            const uint32_t *n_closest = depth ? p.counts + 4 * (depth - 1) + 2 : (implicit0 ? nullptr : p.counts + 3);
            if (depth == 0 && implicit0 && fuse0) {
                // depth 0 in one kernel: trace + shade, the hit record stays in registers
                if (nl == 3) { if (stats) wf_trace_shade0<3, true><<<grid128, DUST_THREADS, 0, ctx->stream>>>(p); else wf_trace_shade0<3, false><<<grid128, DUST_THREADS, 0, ctx->stream>>>(p); }
                else if (nl == 4) { if (stats) wf_trace_shade0<4, true><<<grid128, DUST_THREADS, 0, ctx->stream>>>(p); else wf_trace_shade0<4, false><<<grid128, DUST_THREADS, 0, ctx->stream>>>(p); }
                else { if (stats) wf_trace_shade0<0, true><<<grid128, DUST_THREADS, 0, ctx->stream>>>(p); else wf_trace_shade0<0, false><<<grid128, DUST_THREADS, 0, ctx->stream>>>(p); }
                count_launch(ctx, "wf_trace_shade0");
            } else {
                if (depth >= pt_depth) trace_pt(p.qc, n_closest, 2 * depth, grid128); else trace(p.qc, n_closest, nullptr, grid128);
                if (stats) wf_shade<true><<<grid128, DUST_THREADS, 0, ctx->stream>>>(p); else wf_shade<false><<<grid128, DUST_THREADS, 0, ctx->stream>>>(p);
                count_launch(ctx, "wf_shade");
            }
            if (depth >= pt_depth) trace_pt(p.qs, p.counts + 4 * depth + 1, 2 * depth + 1, grid128); else trace(p.qs, p.counts + 4 * depth + 1, nullptr, grid128);
        }
        wf_resolve<<<(n_slots + 255) / 256, 256, 0, ctx->stream>>>(p);
        count_launch(ctx, "wf_resolve");
        if (stats && std::getenv("DUST_DEBUG_STATS")) {                                           // queue sizes of this batch (diagnostics)
            std::vector<uint32_t> h(kWfCounterWords);
            CU_TRY(ctx, cudaMemcpyAsync(h.data(), ctx->d_wf_counters, kWfCounterWords * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
            CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            for (uint32_t d = 0; d <= cfg.max_bounces; d++)
                fprintf(stderr, "[dust dbg] depth %u: shadow rays %u, bounce rays %u, slow-path rays closest %u / shadow %u\n", d, h[4 * d + 1], h[4 * d + 2],
                        h[4 * kWfMaxDepths + 2 * d], h[4 * kWfMaxDepths + 2 * d + 1]);
        }
        CU_TRY(ctx, pt_err);
        CU_TRY(ctx, cudaGetLastError());
    }
    return DUST_OK;
}

// offsets of the sections of a tree blob: [n16 | blocks | mids | leaves | palette_lin | palette_raw | materials], 256-byte aligned
struct BlobLayout { size_t n16, blocks, mids, leaves, pal, praw, mat, total; };
BlobLayout blob_layout(size_t n_n16, size_t n_mids, size_t n_leaves, size_t n_materials)
{
    BlobLayout L;
    L.n16 = 0;
    L.blocks = align_up(L.n16 + n_n16 * sizeof(DNode16), 256);
    L.mids = align_up(L.blocks + n_n16 * sizeof(uint2), 256);
    L.leaves = align_up(L.mids + n_mids * sizeof(uint4), 256);
    L.pal = align_up(L.leaves + n_leaves * sizeof(uint4), 256);
    L.praw = align_up(L.pal + 256 * sizeof(float4), 256);
    L.mat = align_up(L.praw + 1024, 256);
    L.total = align_up(L.mat + n_materials + 16, 256);
    return L;
}

void tree_set_pointers(DustTree *t, const BlobLayout &L)
{
    uint8_t *base = static_cast<uint8_t *>(t->blob);
    t->blob_bytes = L.total;
    t->n16 = reinterpret_cast<const DNode16 *>(base + L.n16);
    t->blocks = reinterpret_cast<const uint2 *>(base + L.blocks);
    t->mids = reinterpret_cast<const uint4 *>(base + L.mids);
    t->leaves = reinterpret_cast<const uint4 *>(base + L.leaves);
    t->palette_lin = reinterpret_cast<const float4 *>(base + L.pal);
    t->palette_raw = reinterpret_cast<const uchar4 *>(base + L.praw);
    t->materials = base + L.mat;
}

void palette_tables(const uint8_t *palette, float4 pal[256], uint8_t praw[1024])
{
    memset(pal, 0, 256 * sizeof(float4));
    for (int i = 0; i < 255; i++) {
        const uint8_t *c = palette + 4 * i;
        pal[i] = make_float4(srgb_to_linear((float)c[0] / 255.0f), srgb_to_linear((float)c[1] / 255.0f), srgb_to_linear((float)c[2] / 255.0f), 0.0f);
    }
    memset(praw, 0, 1024);
    memcpy(praw, palette, 255 * 4);
}

}  // namespace

extern "C" {

int32_t dust_abi_version(void) { return DUST_ABI_VERSION; }

int32_t dust_ctx_create(int32_t device, DustCtx **out)
{
    if (!out) return DUST_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return DUST_ERR_NO_DEVICE;
    DustCtx *ctx = new (std::nothrow) DustCtx();
    if (!ctx) return DUST_ERR_OUT_OF_MEMORY;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&ctx->d_counters, sizeof(DCounters)) != cudaSuccess) {
        delete ctx;
        return DUST_ERR_CUDA;
    }
    cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, device);
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return DUST_OK;
}

int32_t dust_ctx_destroy(DustCtx *ctx)
{
    if (!ctx) return DUST_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    dust_comm_destroy(ctx);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->wf_scratch) cudaFree(ctx->wf_scratch);
    if (ctx->d_wf_counters) cudaFree(ctx->d_wf_counters);
    if (ctx->d_tile_lists) cudaFree(ctx->d_tile_lists);
    if (ctx->d_sun_lists) cudaFree(ctx->d_sun_lists);
    for (cudaEvent_t ev : ctx->ev_pool) cudaEventDestroy(ev);
    if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
    if (ctx->h_snaps) cudaFreeHost(ctx->h_snaps);
    for (int k = 0; k < 2; k++) if (ctx->side[k]) cudaStreamDestroy(ctx->side[k]);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    for (DustCtx::FrameSlot &fs : ctx->slots) {
        if (fs.scratch) cudaFree(fs.scratch);
        if (fs.traced) cudaEventDestroy(fs.traced);
        if (fs.done) cudaEventDestroy(fs.done);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return DUST_OK;
}

int32_t dust_ctx_set_stream(DustCtx *ctx, void *cuda_stream, int32_t external)
{
    if (!ctx) return DUST_ERR_INVALID_ARGUMENT;
    ctx->stream = external ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return DUST_OK;
}

int32_t dust_ctx_synchronize(DustCtx *ctx)
{
    if (!ctx) return DUST_ERR_INVALID_ARGUMENT;
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return DUST_OK;
}

const char *dust_last_error_string(const DustCtx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

// every tree gets the dense root table and a host copy of the root's block mask (derived data: the blob contract is untouched)
static cudaError_t tree_make_dense(DustCtx *ctx, DustTree *t)
{
    cudaError_t e = cudaMalloc(&t->dense_root, 4096 * sizeof(uint4));
    if (e != cudaSuccess) return e;
    vb_dense_root<<<16, 256, 0, ctx->stream>>>(t->n16, t->blocks, t->mids, t->n_levels, t->dense_root);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(&t->root_blocks, t->blocks, sizeof(uint2), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { cudaFree(t->dense_root); t->dense_root = nullptr; }
    return e;
}

int32_t dust_tree_upload(DustCtx *ctx, const DustTreeDesc *desc, DustTree **out)
{
    if (!ctx || !desc || !out) return DUST_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    HostTree ht{ desc, { 0, 0, 0, 0 } };
    std::string why;
    if (!ht.ok(why)) return fail(ctx, desc->n_levels == 3 || desc->n_levels == 4 ? DUST_ERR_INVALID_ARGUMENT : DUST_ERR_UNSUPPORTED, why);
    Relayout r;
    relayout(ht, r);
    if (r.bad_ptr) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "child pointer outside its pool");
    // material_ptr + popcount(mask) must stay inside `materials`
    for (const uint4 &l : r.leaves) {
        const uint64_t need = (uint64_t)l.z + (uint64_t)__builtin_popcountll(((uint64_t)l.y << 32) | l.x);
        if (need > desc->n_materials) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "leaf material_ptr runs past the material array");
    }
    if (desc->n_materials && !desc->materials) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "materials is NULL");

    float4 pal[256];
    uint8_t praw[1024];
    palette_tables(desc->palette, pal, praw);
    for (uint4 &l : r.leaves) l.w = leaf_avg_albedo(((uint64_t)l.y << 32) | l.x, desc->materials + l.z, praw);   // praw: 256 entries, the last one zero

    // block masks: bit bx<<4|by<<2|bz of node n = "some child bit set in the 4x4x4 block (bx,by,bz) of its 16^3 children"
    std::vector<uint2> blocks(r.n16.size());
    for (size_t n = 0; n < r.n16.size(); n++) {
        uint64_t bm = 0;
        for (uint32_t w = 0; w < 64; w++) {                           // word w holds children x = w >> 2, y = 4 * (w & 3) + (bit >> 4), z = bit & 15
            const uint64_t m = ((uint64_t)r.n16[n].word[w].y << 32) | r.n16[n].word[w].x;
            for (uint32_t bz = 0; bz < 4; bz++)
                if (m & (0x000F000F000F000Full << (4 * bz))) bm |= 1ull << (((w >> 4) << 4) | ((w & 3) << 2) | bz);
        }
        blocks[n] = make_uint2((uint32_t)bm, (uint32_t)(bm >> 32));
    }

    // one allocation: [n16 | blocks | mids | leaves | palette_lin | palette_raw | materials], each 256-byte aligned
    const BlobLayout L = blob_layout(r.n16.size(), r.mids.size(), r.leaves.size(), (size_t)desc->n_materials);
    const size_t off_n16 = L.n16, off_blocks = L.blocks, off_mids = L.mids, off_leaves = L.leaves, off_pal = L.pal, off_praw = L.praw, off_mat = L.mat, total = L.total;

    CU_TRY(ctx, cudaSetDevice(ctx->device));
    DustTree *t = new (std::nothrow) DustTree();
    if (!t) return DUST_ERR_OUT_OF_MEMORY;
    cudaError_t e = cudaMalloc(&t->blob, total);
    if (e != cudaSuccess) { delete t; return fail(ctx, DUST_ERR_OUT_OF_MEMORY, cudaGetErrorString(e)); }
    uint8_t *base = static_cast<uint8_t *>(t->blob);
    auto put = [&](size_t off, const void *src, size_t bytes) {
        return bytes ? cudaMemcpyAsync(base + off, src, bytes, cudaMemcpyHostToDevice, ctx->stream) : cudaSuccess;
    };
    e = put(off_n16, r.n16.data(), r.n16.size() * sizeof(DNode16));
    if (e == cudaSuccess) e = put(off_blocks, blocks.data(), blocks.size() * sizeof(uint2));
    if (e == cudaSuccess) e = put(off_mids, r.mids.data(), r.mids.size() * sizeof(uint4));
    if (e == cudaSuccess) e = put(off_leaves, r.leaves.data(), r.leaves.size() * sizeof(uint4));
    if (e == cudaSuccess) e = put(off_pal, pal, sizeof pal);
    if (e == cudaSuccess) e = put(off_praw, praw, sizeof praw);
    if (e == cudaSuccess) e = put(off_mat, desc->materials, (size_t)desc->n_materials);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);     // host vectors die at return
    if (e != cudaSuccess) { cudaFree(t->blob); delete t; return fail(ctx, DUST_ERR_CUDA, cudaGetErrorString(e)); }
    t->n_levels = desc->n_levels;
    t->blob_bytes = total;
    t->n16 = reinterpret_cast<const DNode16 *>(base + off_n16);
    t->blocks = reinterpret_cast<const uint2 *>(base + off_blocks);
    t->mids = reinterpret_cast<const uint4 *>(base + off_mids);
    t->leaves = reinterpret_cast<const uint4 *>(base + off_leaves);
    t->palette_lin = reinterpret_cast<const float4 *>(base + off_pal);
    t->palette_raw = reinterpret_cast<const uchar4 *>(base + off_praw);
    t->materials = base + off_mat;
    t->n_leaves = r.leaves.size(); t->n_mids = r.mids.size(); t->n_n16 = r.n16.size(); t->n_materials = desc->n_materials;
    if (r.leaves.empty()) { for (int a = 0; a < 3; a++) { t->bmin[a] = 0; t->bmax[a] = 0; } }
    else for (int a = 0; a < 3; a++) { t->bmin[a] = r.bmin[a]; t->bmax[a] = r.bmax[a]; }
    e = tree_make_dense(ctx, t);
    if (e != cudaSuccess) { cudaFree(t->blob); delete t; return fail(ctx, e == cudaErrorMemoryAllocation ? DUST_ERR_OUT_OF_MEMORY : DUST_ERR_CUDA, cudaGetErrorString(e)); }
    *out = t;
    return DUST_OK;
}

// ---- device-side build (build.cuh) ---------------------------------------------------------------------

int32_t dust_tree_build_voxels(DustCtx *ctx, const DustVoxel *voxels, uint64_t n, const uint32_t *size, const uint8_t *palette,
                               uint32_t memory, DustTree **out)
{
    if (!ctx || !out || !size || !palette || (n && !voxels)) return DUST_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (size[0] > 256 || size[1] > 256 || size[2] > 256) return fail(ctx, DUST_ERR_UNSUPPORTED, "model larger than 256^3 (loader.rs:365)");
    if (n > (1ull << 31)) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "too many voxels");
    if (memory != DUST_MEM_HOST && memory != DUST_MEM_DEVICE) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "bad memory space");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    // scratch: [grid u64 x 16 Mi | block_counts -> mat_ptr | pop_block -> mat_pos | flag_tree -> leaf_slot | leaf_mask | mid_mask | mid_flag | mid_slot | counts | voxels]
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_grid = take((size_t)VB_BLOCKS * 64 * 8), o_cnt = take(VB_BLOCKS * 4), o_pop = take(VB_BLOCKS * 4), o_flag = take(VB_BLOCKS * 4);
    const size_t o_lmask = take(VB_BLOCKS * 8), o_mmask = take(VB_MIDS * 8), o_mflag = take(VB_MIDS * 4), o_mslot = take(VB_MIDS * 4);
    const size_t o_counts = take(sizeof(VbCounts)), o_partial = take(1024 * 4), o_vox = take(memory == DUST_MEM_HOST ? (size_t)n * 4 : 0);
    const int32_t rc = ensure_scratch(ctx, off);
    if (rc != DUST_OK) return rc;
    uint8_t *sc = static_cast<uint8_t *>(ctx->scratch);
    unsigned long long *grid = reinterpret_cast<unsigned long long *>(sc + o_grid), *lmask = reinterpret_cast<unsigned long long *>(sc + o_lmask),
                       *mmask = reinterpret_cast<unsigned long long *>(sc + o_mmask);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(sc + o_cnt), *pop = reinterpret_cast<uint32_t *>(sc + o_pop), *flag = reinterpret_cast<uint32_t *>(sc + o_flag),
             *mflag = reinterpret_cast<uint32_t *>(sc + o_mflag), *mslot = reinterpret_cast<uint32_t *>(sc + o_mslot);
    VbCounts *d_counts = reinterpret_cast<VbCounts *>(sc + o_counts);
    uint32_t *partial = reinterpret_cast<uint32_t *>(sc + o_partial);
    const uchar4 *d_vox = reinterpret_cast<const uchar4 *>(voxels);
    cudaStream_t st = ctx->stream;
    if (memory == DUST_MEM_HOST && n) {
        CU_TRY(ctx, cudaMemcpyAsync(sc + o_vox, voxels, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        d_vox = reinterpret_cast<const uchar4 *>(sc + o_vox);
    }
    CU_TRY(ctx, cudaMemsetAsync(sc + o_grid, 0, o_pop - o_grid, st));                      // grid + block_counts
    VbCounts h_counts;
    memset(&h_counts, 0, sizeof h_counts);
    for (int a = 0; a < 3; a++) { h_counts.bmin[a] = INT32_MAX; h_counts.bmax[a] = INT32_MIN; }
    CU_TRY(ctx, cudaMemcpyAsync(d_counts, &h_counts, sizeof h_counts, cudaMemcpyHostToDevice, st));
    if (n) { vb_scatter<<<(uint32_t)((n + 255) / 256), 256, 0, st>>>(d_vox, (uint32_t)n, size[1], grid, cnt); count_launch(ctx); }
    vb_leaves<<<VB_BLOCKS / 256, 256, 0, st>>>(grid, lmask, flag, pop); count_launch(ctx);
    vb_mid_masks<<<VB_MIDS / 256, 256, 0, st>>>(flag, mmask, mflag); count_launch(ctx);
    CU_TRY(ctx, cudaMemcpyAsync(mslot, mflag, VB_MIDS * 4, cudaMemcpyDeviceToDevice, st));
    auto scan = [&](uint32_t *a, uint32_t n_elems, uint32_t *total) {
        vb_scan_partials<<<n_elems / 1024, 1024, 0, st>>>(a, partial);
        vb_scan_apply<<<n_elems / 1024, 1024, 0, st>>>(a, partial, total);
        count_launch(ctx); count_launch(ctx);
    };
    scan(cnt, VB_BLOCKS, nullptr);                                                         // material_ptr (collector.rs:76-82)
    scan(pop, VB_BLOCKS, &d_counts->n_materials);                                          // where each block's bytes go
    scan(flag, VB_BLOCKS, &d_counts->n_leaves);                                            // leaf slots, Tree::iter_leaf order
    scan(mslot, VB_MIDS, &d_counts->n_mids);
    CU_TRY(ctx, cudaGetLastError());
    CU_TRY(ctx, cudaMemcpyAsync(&h_counts, d_counts, sizeof h_counts, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));                                                // sizes of the blob

    const BlobLayout L = blob_layout(1, h_counts.n_mids, h_counts.n_leaves, h_counts.n_materials);
    DustTree *t = new (std::nothrow) DustTree();
    if (!t) return DUST_ERR_OUT_OF_MEMORY;
    cudaError_t e = cudaMalloc(&t->blob, L.total);
    if (e != cudaSuccess) { delete t; return fail(ctx, DUST_ERR_OUT_OF_MEMORY, cudaGetErrorString(e)); }
    tree_set_pointers(t, L);
    uint8_t *base = static_cast<uint8_t *>(t->blob);
    float4 pal[256];
    uint8_t praw[1024];
    palette_tables(palette, pal, praw);
    e = cudaMemcpyAsync(base + L.pal, pal, sizeof pal, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(base + L.praw, praw, sizeof praw, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        vb_emit_leaves<<<VB_BLOCKS / 256, 256, 0, st>>>(grid, lmask, flag, cnt, pop, reinterpret_cast<uint4 *>(base + L.leaves), base + L.mat, d_counts);
        vb_emit_mids<<<VB_MIDS / 256, 256, 0, st>>>(mmask, mslot, flag, reinterpret_cast<uint4 *>(base + L.mids));
        vb_emit_root<<<1, 64, 0, st>>>(mflag, mslot, reinterpret_cast<DNode16 *>(base + L.n16), reinterpret_cast<uint2 *>(base + L.blocks));
        count_launch(ctx); count_launch(ctx); count_launch(ctx);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_counts, d_counts, sizeof h_counts, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);                                   // bounds; pal / praw die at return
    if (e != cudaSuccess) { cudaFree(t->blob); delete t; return fail(ctx, DUST_ERR_CUDA, cudaGetErrorString(e)); }
    if (h_counts.max_mat_end > h_counts.n_materials) {                                    // duplicate voxels: same verdict as dust_tree_upload
        cudaFree(t->blob); delete t;
        return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "leaf material_ptr runs past the material array (duplicate voxels in the model)");
    }
    if (h_counts.n_leaves) {
        // avg_albedo of every leaf: integer sums + libm powf on the host, so that both build paths store the same word
        std::vector<uint4> hl(h_counts.n_leaves);
        std::vector<uint8_t> hm((size_t)h_counts.n_materials + 16);
        e = cudaMemcpyAsync(hl.data(), base + L.leaves, hl.size() * sizeof(uint4), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hm.data(), base + L.mat, (size_t)h_counts.n_materials, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) {
            for (uint4 &l : hl) l.w = leaf_avg_albedo(((uint64_t)l.y << 32) | l.x, hm.data() + l.z, praw);
            e = cudaMemcpyAsync(base + L.leaves, hl.data(), hl.size() * sizeof(uint4), cudaMemcpyHostToDevice, st);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { cudaFree(t->blob); delete t; return fail(ctx, DUST_ERR_CUDA, cudaGetErrorString(e)); }
    }
    t->n_levels = 3;
    t->n_leaves = h_counts.n_leaves; t->n_mids = h_counts.n_mids; t->n_n16 = 1; t->n_materials = h_counts.n_materials;
    for (int a = 0; a < 3; a++) { t->bmin[a] = h_counts.n_leaves ? h_counts.bmin[a] : 0; t->bmax[a] = h_counts.n_leaves ? h_counts.bmax[a] : 0; }
    e = tree_make_dense(ctx, t);
    if (e != cudaSuccess) { cudaFree(t->blob); delete t; return fail(ctx, e == cudaErrorMemoryAllocation ? DUST_ERR_OUT_OF_MEMORY : DUST_ERR_CUDA, cudaGetErrorString(e)); }
    *out = t;
    return DUST_OK;
}

int32_t dust_tree_layout(const DustTree *tree, DustTreeLayout *out)
{
    if (!tree || !out) return DUST_ERR_INVALID_ARGUMENT;
    const BlobLayout L = blob_layout(tree->n_n16, tree->n_mids, tree->n_leaves, tree->n_materials);
    out->n_levels = tree->n_levels;
    out->n_nodes16 = tree->n_n16; out->n_mids = tree->n_mids; out->n_leaves = tree->n_leaves; out->n_materials = tree->n_materials;
    out->off_nodes16 = L.n16; out->off_blocks = L.blocks; out->off_mids = L.mids; out->off_leaves = L.leaves;
    out->off_palette_linear = L.pal; out->off_palette = L.praw; out->off_materials = L.mat; out->total_bytes = L.total;
    for (int a = 0; a < 3; a++) { out->bounds_min[a] = tree->bmin[a]; out->bounds_max[a] = tree->bmax[a]; }
    return DUST_OK;
}

int32_t dust_tree_download(DustCtx *ctx, const DustTree *tree, void *dst, uint64_t capacity)
{
    if (!ctx || !tree || !dst) return DUST_ERR_INVALID_ARGUMENT;
    if (capacity < tree->blob_bytes) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "dust_tree_download: buffer smaller than DustTreeLayout.total_bytes");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaMemcpyAsync(dst, tree->blob, tree->blob_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return DUST_OK;
}

int32_t dust_tree_destroy(DustCtx *ctx, DustTree *tree)
{
    if (!ctx || !tree) return DUST_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (tree->blob) cudaFree(tree->blob);
    if (tree->dense_root) cudaFree(tree->dense_root);
    delete tree;
    return DUST_OK;
}

int32_t dust_tree_leaf_count(const DustTree *tree, uint64_t *out)
{
    if (!tree || !out) return DUST_ERR_INVALID_ARGUMENT;
    *out = tree->n_leaves;
    return DUST_OK;
}

int32_t dust_tree_device_bytes(const DustTree *tree, uint64_t *out)
{
    if (!tree || !out) return DUST_ERR_INVALID_ARGUMENT;
    *out = tree->blob_bytes + (tree->dense_root ? 4096 * sizeof(uint4) : 0);            // the blob + the derived dense root table
    return DUST_OK;
}

// Instance records, padded world boxes and the SAH BVH over them for `n` (tree, transform) pairs.
// what the kernels walk: the wide nodes, or (-DDUST_BINARY_BVH) the binary nodes as they are
struct BvhBlob {
    std::vector<DBvhNode> nodes;                                         // binary SAH tree (the source of the collapse)
    std::vector<uint4> wide;
    BvhGrid grid{};
    const void *data() const {
#ifdef DUST_BINARY_BVH
        return nodes.data();
#else
        return wide.data();
#endif
    }
    size_t bytes() const {
#ifdef DUST_BINARY_BVH
        return nodes.size() * sizeof(DBvhNode);
#else
        return wide.size() * sizeof(uint4);
#endif
    }
    static size_t capacity(uint32_t n) { return (size_t)n * sizeof(DBvhNode); }   // n leaves: at most n - 1 nodes of either kind, 64 bytes each
};

static int32_t scene_arrays(DustCtx *ctx, uint32_t n, const DustTree *const *trees, const float *o2w, std::vector<DInstance> &inst,
                            BvhBlob &bvh, Box &root_box, int &root_ref, uint32_t &max_levels, uint32_t &min_levels,
                            std::vector<float4> &iboxes)
{
    std::vector<DBvhNode> &nodes = bvh.nodes;
    inst.assign(n, DInstance{});
    std::vector<Box> boxes(n);
    max_levels = 3; min_levels = 4;
    for (uint32_t k = 0; k < n; k++) {
        const DustTree *t = trees[k];
        if (!t) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "instance without a tree");
        DInstance &d = inst[k];
        const float *m = o2w + 12 * (size_t)k;
        if (!world2obj(m, d.w2o)) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "instance transform is singular or not finite");
        for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) d.o2w[3 * r + c] = m[4 * r + c]; d.o2w_t[r] = m[4 * r + 3]; }
        d.n_levels = t->n_levels;
        d.n16 = t->n16; d.blocks = t->blocks; d.mids = t->mids; d.leaves = t->leaves; d.materials = t->materials; d.palette_lin = t->palette_lin; d.palette_raw = t->palette_raw;
        // occupied bounds are multiples of 4 voxels (whole leaves): leaf units, lo | size << 16
        for (int a = 0; a < 3; a++) { d.bounds_lo[a] = (float)t->bmin[a]; d.bounds_hi[a] = (float)t->bmax[a]; }
        d._pad1 = 0;
        d.dense_root = t->dense_root; d.root_blocks = t->root_blocks; d._pad2[0] = d._pad2[1] = 0;
        max_levels = std::max(max_levels, t->n_levels);
        min_levels = std::min(min_levels, t->n_levels);
        // world box of the occupied object-space bounds, padded so that culling can never remove
        // a hit the exact per-instance traversal would report (DESIGN.md "Instance BVH")
        Box b;
        for (int a = 0; a < 3; a++) { b.lo[a] = INFINITY; b.hi[a] = -INFINITY; }
        float big = 0.0f;
        for (int corner = 0; corner < 8; corner++) {
            const float p[3] = { (float)(corner & 1 ? t->bmax[0] : t->bmin[0]), (float)(corner & 2 ? t->bmax[1] : t->bmin[1]),
                                 (float)(corner & 4 ? t->bmax[2] : t->bmin[2]) };
            for (int a = 0; a < 3; a++) {
                const float w = m[4 * a] * p[0] + m[4 * a + 1] * p[1] + m[4 * a + 2] * p[2] + m[4 * a + 3];
                b.lo[a] = std::min(b.lo[a], w); b.hi[a] = std::max(b.hi[a], w);
                big = std::max(big, std::fabs(w));
            }
        }
        const float pad = 0.02f + 1e-5f * big;
        for (int a = 0; a < 3; a++) { b.lo[a] -= pad; b.hi[a] += pad; }
        boxes[k] = b;
    }
    if (n > (1u << 20)) return fail(ctx, DUST_ERR_UNSUPPORTED, "more than 2^20 instances");
    iboxes.resize(2 * (size_t)n);                                       // the same padded boxes, for the per-tile candidate lists
    for (uint32_t k = 0; k < n; k++) {
        iboxes[2 * (size_t)k] = make_float4(boxes[k].lo[0], boxes[k].lo[1], boxes[k].lo[2], 0.0f);
        iboxes[2 * (size_t)k + 1] = make_float4(boxes[k].hi[0], boxes[k].hi[1], boxes[k].hi[2], 0.0f);
    }
    nodes.clear();
    root_box = Box{};
    root_ref = -1;
    int depth = 0;
    if (n) {
        std::vector<uint32_t> ids(n);
        for (uint32_t k = 0; k < n; k++) ids[k] = k;
        nodes.reserve((size_t)n);
        root_ref = build_bvh(nodes, ids, boxes, 0, n, root_box, 0, depth);
        if (depth > DUST_MAX_BVH_DEPTH) return fail(ctx, DUST_ERR_UNSUPPORTED, "instance BVH deeper than the traversal stack");
#ifndef DUST_BINARY_BVH
        bvh.wide.clear();
        if (root_ref >= 0) {
            int wdepth = 0;
            bool ok = true;
            bvh.wide.reserve((size_t)n * 4);
            bvh.grid = bvh_grid(root_box);
            root_ref = build_wide(bvh.wide, nodes, bvh.grid, root_ref, 0, wdepth, ok);
            if (!ok) return fail(ctx, DUST_ERR_UNSUPPORTED, "instance boxes do not fit the wide BVH's 16-bit grid");
            // one stack entry per level, and an empty candidate list must always find room for one more expansion (cull_room)
            if (wdepth + 1 + DUST_BVH_WIDTH + 1 > DUST_SCRATCH_WORDS) return fail(ctx, DUST_ERR_UNSUPPORTED, "instance BVH deeper than the traversal stack");
        }
#endif
    }
    return DUST_OK;
}

static void scene_set_view(DustScene *s, const Box &root_box, int root_ref, const BvhBlob &bvh)
{
    for (int a = 0; a < 3; a++) { s->view.grid_org[a] = bvh.grid.org[a]; s->view.grid_step[a] = (float)bvh.grid.step[a]; }
    s->view.inst = s->d_inst; s->view.bvh = s->d_bvh; s->view.prev_o2w = s->d_prev; s->view.inst_boxes = s->d_boxes;
    for (int a = 0; a < 3; a++) { s->view.root_lo[a] = root_box.lo[a]; s->view.root_hi[a] = root_box.hi[a]; }
    s->view.root_ref = root_ref; s->view.n_inst = s->n_inst;
    s->view.id_bits = 8;                                   // candidate key = entry t (high bits) | instance (low id_bits)
    while ((1u << s->view.id_bits) < s->n_inst) s->view.id_bits++;
}

int32_t dust_scene_create(DustCtx *ctx, const DustInstance *instances, uint32_t n, DustScene **out)
{
    if (!ctx || !out || (n && !instances)) return DUST_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    std::vector<const DustTree *> trees(n);
    std::vector<float> o2w(12 * (size_t)n);
    for (uint32_t k = 0; k < n; k++) { trees[k] = instances[k].tree; memcpy(&o2w[12 * (size_t)k], instances[k].obj2world, 12 * sizeof(float)); }
    std::vector<DInstance> inst;
    BvhBlob bvh;
    Box root_box{};
    int root_ref = -1;
    uint32_t max_levels = 3, min_levels = 4;
    std::vector<float4> iboxes;
    const int32_t rc = scene_arrays(ctx, n, trees.data(), o2w.data(), inst, bvh, root_box, root_ref, max_levels, min_levels, iboxes);
    if (rc != DUST_OK) return rc;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    DustScene *s = new (std::nothrow) DustScene();
    if (!s) return DUST_ERR_OUT_OF_MEMORY;
    s->n_inst = n; s->n_bvh = (uint32_t)bvh.nodes.size(); s->max_levels = max_levels; s->min_levels = std::min(min_levels, max_levels);
    if (n) {
        cudaError_t e = cudaMalloc(&s->d_inst, sizeof(DInstance) * n);
        if (e == cudaSuccess) e = cudaMalloc(&s->d_bvh, BvhBlob::capacity(n));
        if (e == cudaSuccess) e = cudaMemcpyAsync(s->d_inst, inst.data(), sizeof(DInstance) * n, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMalloc(&s->d_prev, sizeof(float) * 12 * n);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s->d_prev, o2w.data(), sizeof(float) * 12 * n, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess && bvh.bytes()) e = cudaMemcpyAsync(s->d_bvh, bvh.data(), bvh.bytes(), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMalloc(&s->d_boxes, sizeof(float4) * iboxes.size());
        if (e == cudaSuccess) e = cudaMemcpyAsync(s->d_boxes, iboxes.data(), sizeof(float4) * iboxes.size(), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            if (s->d_boxes) cudaFree(s->d_boxes);
            if (s->d_inst) cudaFree(s->d_inst);
            if (s->d_bvh) cudaFree(s->d_bvh);
            if (s->d_prev) cudaFree(s->d_prev);
            delete s;
            return fail(ctx, DUST_ERR_CUDA, cudaGetErrorString(e));
        }
    }
    s->trees = std::move(trees);
    s->o2w = std::move(o2w);
    scene_set_view(s, root_box, root_ref, bvh);
    *out = s;
    return DUST_OK;
}

int32_t dust_scene_update_transforms(DustCtx *ctx, DustScene *scene, const float *obj2world, uint32_t n)
{
    if (!ctx || !scene || (n && !obj2world)) return DUST_ERR_INVALID_ARGUMENT;
    if (n != scene->n_inst) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "one transform per instance");
    if (!n) return DUST_OK;
    std::vector<DInstance> inst;
    BvhBlob bvh;
    Box root_box{};
    int root_ref = -1;
    uint32_t max_levels = 3, min_levels = 4;
    std::vector<float4> iboxes;
    const int32_t rc = scene_arrays(ctx, n, scene->trees.data(), obj2world, inst, bvh, root_box, root_ref, max_levels, min_levels, iboxes);
    if (rc != DUST_OK) return rc;                                        // the scene is unchanged on error
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    // this frame's transforms become "last frame" for DustGBuffer.motion (PreviousFrameGlobalTransform, standard.rs:845-854)
    CU_TRY(ctx, cudaMemcpyAsync(scene->d_prev, scene->o2w.data(), sizeof(float) * 12 * n, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(scene->d_inst, inst.data(), sizeof(DInstance) * n, cudaMemcpyHostToDevice, ctx->stream));
    if (bvh.bytes()) CU_TRY(ctx, cudaMemcpyAsync(scene->d_bvh, bvh.data(), bvh.bytes(), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(scene->d_boxes, iboxes.data(), sizeof(float4) * iboxes.size(), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));                     // stream-ordered after the frames in flight; host arrays die at return
    memcpy(scene->o2w.data(), obj2world, sizeof(float) * 12 * n);
    scene->n_bvh = (uint32_t)bvh.nodes.size();
    scene_set_view(scene, root_box, root_ref, bvh);
    return DUST_OK;
}

int32_t dust_scene_destroy(DustCtx *ctx, DustScene *scene)
{
    if (!ctx || !scene) return DUST_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (scene->d_inst) cudaFree(scene->d_inst);
    if (scene->d_bvh) cudaFree(scene->d_bvh);
    if (scene->d_boxes) cudaFree(scene->d_boxes);
    if (scene->d_prev) cudaFree(scene->d_prev);
    delete scene;
    return DUST_OK;
}

int32_t dust_scene_set_previous_transforms(DustCtx *ctx, DustScene *scene, const float *obj2world_prev, uint32_t n)
{
    if (!ctx || !scene || (n && !obj2world_prev)) return DUST_ERR_INVALID_ARGUMENT;
    if (n != scene->n_inst) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "one previous transform per instance");
    if (!n) return DUST_OK;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaMemcpyAsync(scene->d_prev, obj2world_prev, sizeof(float) * 12 * n, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));                    // the caller's array may die at return
    return DUST_OK;
}

int32_t dust_shard_layout(uint32_t width, uint32_t height, uint32_t shard_count, uint32_t *tiles_per_shard, uint64_t *pixels_per_shard)
{
    if (!width || !height || !shard_count) return DUST_ERR_INVALID_ARGUMENT;
    const uint32_t tiles = ((width + 31) / 32) * ((height + 31) / 32);
    const uint32_t per = (tiles + shard_count - 1) / shard_count;
    if (tiles_per_shard) *tiles_per_shard = per;
    if (pixels_per_shard) *pixels_per_shard = (uint64_t)per * 1024u;
    return DUST_OK;
}

// dust_trace (slot == NULL: the call returns when the host buffers are complete) and dust_trace_async (slot: the frame's staging
// buffer; kernels on the context's stream, D2H copies on the copy stream behind them, nothing waits on the host)
static int32_t trace_impl(DustCtx *ctx, const DustScene *scene, const DustCamera *camera, const DustSky *sky,
                          const DustTraceConfig *config, const DustFrame *frame, DustCtx::FrameSlot *slot)
{
    if (!ctx || !scene || !camera || !sky || !config || !frame) return DUST_ERR_INVALID_ARGUMENT;
    const DustTraceConfig &cfg = *config;
    if (!cfg.width || !cfg.height || !cfg.spp || !cfg.shard_count || cfg.shard_index >= cfg.shard_count || !frame->radiance)
        return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "bad trace config");
    if ((uint64_t)cfg.width * cfg.height > 0x7FFFFFFFull) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "frame too large");
    const bool want_hits = (cfg.flags & DUST_TRACE_WRITE_HITS) != 0;
    if (want_hits && (cfg.shard_count != 1 || !frame->hit_instance || !frame->hit_leaf || !frame->hit_voxel || !frame->hit_t))
        return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "hit buffers need shard_count == 1 and all four arrays");
    if (frame->radiance_format != DUST_FMT_RGBA32F && frame->radiance_format != DUST_FMT_RGBA16F && frame->radiance_format != DUST_FMT_YCOCG16F)
        return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "bad radiance format");
    const DustGBuffer *gb = frame->gbuffer;
    if (gb && cfg.shard_count != 1) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "the G-buffer needs shard_count == 1");
    CU_TRY(ctx, cudaSetDevice(ctx->device));

    uint32_t tiles_per_shard = 0; uint64_t shard_pixels = 0;
    dust_shard_layout(cfg.width, cfg.height, cfg.shard_count, &tiles_per_shard, &shard_pixels);
    const uint64_t out_pixels = cfg.shard_count > 1 ? shard_pixels : (uint64_t)cfg.width * cfg.height;
    const size_t px_bytes = frame->radiance_format == DUST_FMT_RGBA32F ? 16 : 8;

    FrameParams p;
    memset(&p, 0, sizeof p);
    p.scene = scene->view;
    p.cam = *camera; p.sky = *sky; p.cfg = cfg;
    dust_sun_terms(sky, p.sun_dir, &p.sun_ar2, p.sun_nee);
    p.tiles_x = (cfg.width + 31) / 32; p.tiles_y = (cfg.height + 31) / 32; p.n_local_tiles = tiles_per_shard;
    p.ycocg = frame->radiance_format == DUST_FMT_YCOCG16F;

    uint8_t *scratch = nullptr;
    size_t off_hi = 0, off_hl = 0, off_hv = 0, off_ht = 0, off_gb[5] = { 0, 0, 0, 0, 0 };
    const uint64_t npx = (uint64_t)cfg.width * cfg.height;
    if (frame->memory == DUST_MEM_HOST) {
        size_t need = align_up(out_pixels * px_bytes, 256);
        if (want_hits) {
            off_hi = need; need += align_up(npx * 4, 256);
            off_hl = need; need += align_up(npx * 4, 256);
            off_ht = need; need += align_up(npx * 4, 256);
            off_hv = need; need += align_up(npx, 256);
        }
        if (gb) {
            const void *src[5] = { gb->albedo, gb->depth, gb->normal, gb->motion, gb->voxel_id };
            const size_t bytes[5] = { 4, 4, 4, 8, 4 };
            for (int k = 0; k < 5; k++) if (src[k]) { off_gb[k] = need; need += align_up(npx * bytes[k], 256); }
        }
        if (slot) {
            if (need > slot->scratch_bytes) {
                if (slot->scratch) { CU_TRY(ctx, cudaFree(slot->scratch)); slot->scratch = nullptr; slot->scratch_bytes = 0; }
                CU_TRY(ctx, cudaMalloc(&slot->scratch, need));
                slot->scratch_bytes = need;
            }
            scratch = static_cast<uint8_t *>(slot->scratch);
        } else {
            int32_t rc = ensure_scratch(ctx, need);
            if (rc != DUST_OK) return rc;
            scratch = static_cast<uint8_t *>(ctx->scratch);
        }
        if (frame->radiance_format == DUST_FMT_RGBA32F) p.radiance32 = reinterpret_cast<float4 *>(scratch);
        else p.radiance16 = reinterpret_cast<__half *>(scratch);
        if (want_hits) {
            p.hit_instance = reinterpret_cast<uint32_t *>(scratch + off_hi); p.hit_leaf = reinterpret_cast<uint32_t *>(scratch + off_hl);
            p.hit_t = reinterpret_cast<float *>(scratch + off_ht); p.hit_voxel = scratch + off_hv;
        }
        if (gb) {
            if (gb->albedo) p.gb.albedo = reinterpret_cast<uint32_t *>(scratch + off_gb[0]);
            if (gb->depth) p.gb.depth = reinterpret_cast<float *>(scratch + off_gb[1]);
            if (gb->normal) p.gb.normal = reinterpret_cast<uint32_t *>(scratch + off_gb[2]);
            if (gb->motion) p.gb.motion = reinterpret_cast<uint16_t *>(scratch + off_gb[3]);
            if (gb->voxel_id) p.gb.voxel_id = reinterpret_cast<uint32_t *>(scratch + off_gb[4]);
            // the reference leaves normal / voxel_id of miss pixels untouched: start from the caller's contents
            if (gb->normal) CU_TRY(ctx, cudaMemcpyAsync(p.gb.normal, gb->normal, npx * 4, cudaMemcpyHostToDevice, ctx->stream));
            if (gb->voxel_id) CU_TRY(ctx, cudaMemcpyAsync(p.gb.voxel_id, gb->voxel_id, npx * 4, cudaMemcpyHostToDevice, ctx->stream));
        }
    } else if (frame->memory == DUST_MEM_DEVICE) {
        if (!aligned16(frame->radiance)) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "device radiance buffer must be 16-byte aligned");
        if (frame->radiance_format == DUST_FMT_RGBA32F) p.radiance32 = static_cast<float4 *>(frame->radiance);
        else p.radiance16 = static_cast<__half *>(frame->radiance);
        if (want_hits) { p.hit_instance = frame->hit_instance; p.hit_leaf = frame->hit_leaf; p.hit_voxel = frame->hit_voxel; p.hit_t = frame->hit_t; }
        if (gb) p.gb = *gb;
    } else return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "bad memory space");

    const bool stats = (cfg.flags & DUST_TRACE_STATS) != 0;
    p.counters = ctx->d_counters;
    if (stats) CU_TRY(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof(DCounters), ctx->stream));
    ctx->timing_on = (cfg.flags & DUST_TRACE_TIMING) != 0;
    if (ctx->timing_on) {
        ctx->launch_recs.clear();
        ctx->launch_dropped = 0;
        if (!ctx->ev_begin) CU_TRY(ctx, cudaEventCreate(&ctx->ev_begin));
        if (!ctx->h_snaps) CU_TRY(ctx, cudaMallocHost(reinterpret_cast<void **>(&ctx->h_snaps), kMaxLaunchRecs * sizeof(DCounters)));
        ctx->snaps_valid = stats;
    }
    if (cfg.shard_count > 1 && tiles_per_shard * cfg.shard_count != p.tiles_x * p.tiles_y)
        CU_TRY(ctx, cudaMemsetAsync(p.radiance32 ? (void *)p.radiance32 : (void *)p.radiance16, 0, out_pixels * px_bytes, ctx->stream));

    // 2^slot_log2 threads per pixel, one sample each per round (frame_kernel's thread mapping)
    p.slot_log2 = cfg.spp >= 3 ? 2u : (cfg.spp == 2 ? 1u : 0u);
    const uint64_t n_threads = ((uint64_t)tiles_per_shard * 1024u) << p.slot_log2;
    if (n_threads / DUST_THREADS > 0x7FFFFFFFull) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "frame too large");
    const uint32_t grid = (uint32_t)(n_threads / DUST_THREADS);
    const int nl = scene->n_inst == 0 || scene->min_levels == scene->max_levels ? (scene->n_inst ? (int)scene->max_levels : 3) : 0;
    // Pipeline: frames with bounces re-order their rays between path vertices (wavefront.cuh); coherent
    // frames (primary + sun only) and small frames stay in the single-launch megakernel.
    bool wavefront = cfg.max_bounces >= 1 && (uint64_t)tiles_per_shard * 1024u * cfg.spp >= (1u << 18);
    if (cfg.flags & DUST_TRACE_PIPELINE_WAVEFRONT) wavefront = true;
    if (cfg.flags & DUST_TRACE_PIPELINE_MEGAKERNEL) wavefront = false;
    if (ctx->timing_on) CU_TRY(ctx, cudaEventRecord(ctx->ev_begin, ctx->stream));
#ifndef DUST_NO_TILE_LISTS
    if (scene->n_inst > 1 && scene->d_boxes) {
        // camera rays: one conservative frustum test per tile and instance instead of a BVH walk per ray (trace_core.cuh cull_from_list)
        const size_t words = (size_t)tiles_per_shard * DUST_TILE_LIST_WORDS;
        if (words > ctx->tile_lists_words) {
            if (ctx->d_tile_lists) { CU_TRY(ctx, cudaStreamSynchronize(ctx->stream)); CU_TRY(ctx, cudaFree(ctx->d_tile_lists)); ctx->d_tile_lists = nullptr; ctx->tile_lists_words = 0; }
            CU_TRY(ctx, cudaMalloc(&ctx->d_tile_lists, words * sizeof(uint32_t)));
            ctx->tile_lists_words = words;
        }
        wf_tile_cull<<<(tiles_per_shard + 3) / 4, 128, 0, ctx->stream>>>(p, ctx->d_tile_lists);
        count_launch(ctx, "wf_tile_cull");
        p.tile_lists = ctx->d_tile_lists;
    }
#endif
#ifndef DUST_NO_SUN_GRID
    if (wavefront && scene->n_inst > 1 && scene->d_boxes) {
        // sun shadow rays: a 2D look-up by where the origin projects along the sun direction instead of a BVH walk per ray (kernels.cuh sun_list_of)
        const double sx = p.sun_dir[0], sy = p.sun_dir[1], sz = p.sun_dir[2];
        double ax[3] = { 1.0, 0.0, 0.0 };
        if (std::fabs(sx) > std::fabs(sy) || std::fabs(sx) > std::fabs(sz)) { ax[0] = 0.0; if (std::fabs(sy) <= std::fabs(sz)) ax[1] = 1.0; else ax[2] = 1.0; }
        double u[3] = { sy * ax[2] - sz * ax[1], sz * ax[0] - sx * ax[2], sx * ax[1] - sy * ax[0] };
        const double ul = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        if (std::isfinite(ul) && ul > 1e-6) {
            for (double &c : u) c /= ul;
            const double v[3] = { sy * u[2] - sz * u[1], sz * u[0] - sx * u[2], sx * u[1] - sy * u[0] };
            double umin = INFINITY, umax = -INFINITY, vmin = INFINITY, vmax = -INFINITY, big = 0.0;
            for (int corner = 0; corner < 8; corner++) {
                const double c[3] = { corner & 1 ? scene->view.root_hi[0] : scene->view.root_lo[0], corner & 2 ? scene->view.root_hi[1] : scene->view.root_lo[1],
                                      corner & 4 ? scene->view.root_hi[2] : scene->view.root_lo[2] };
                const double pu = c[0] * u[0] + c[1] * u[1] + c[2] * u[2], pv = c[0] * v[0] + c[1] * v[1] + c[2] * v[2];
                umin = std::min(umin, pu); umax = std::max(umax, pu); vmin = std::min(vmin, pv); vmax = std::max(vmax, pv);
                big = std::max({ big, std::fabs(c[0]), std::fabs(c[1]), std::fabs(c[2]) });
            }
            const double eu = 1e-3 * (umax - umin) + 1e-6, ev = 1e-3 * (vmax - vmin) + 1e-6;
            umin -= eu; umax += eu; vmin -= ev; vmax += ev;
            const double du = (umax - umin) / DUST_SUN_GRID, dv = (vmax - vmin) / DUST_SUN_GRID;
            if (std::isfinite(du) && std::isfinite(dv) && du > 0.0 && dv > 0.0) {
                if (!ctx->d_sun_lists) CU_TRY(ctx, cudaMalloc(&ctx->d_sun_lists, (size_t)DUST_SUN_GRID * DUST_SUN_GRID * DUST_TILE_LIST_WORDS * sizeof(uint32_t)));
                for (int k = 0; k < 3; k++) { p.sun_u[k] = (float)u[k]; p.sun_v[k] = (float)v[k]; }
                p.sun_u0 = (float)umin; p.sun_v0 = (float)vmin; p.sun_du = (float)du; p.sun_dv = (float)dv;
                p.sun_inv_du = (float)(1.0 / du); p.sun_inv_dv = (float)(1.0 / dv);
                // margin: float rounding of a ray's projection (a few 2^-24 of its coordinates), of the cell index, of the float basis
                p.sun_eps = (float)(1e-3 * std::max(du, dv) + 2e-5 * (big + (umax - umin) + (vmax - vmin)));
                wf_sun_grid<<<DUST_SUN_GRID * DUST_SUN_GRID / 4, 128, 0, ctx->stream>>>(p, ctx->d_sun_lists);
                count_launch(ctx, "wf_sun_grid");
                p.sun_lists = ctx->d_sun_lists;
            }
        }
    }
#endif
    if (wavefront) {
        const int32_t rc = trace_wavefront(ctx, p, tiles_per_shard, nl, stats);
        if (rc != DUST_OK) { ctx->timing_on = false; return rc; }
    } else {
        auto launch = [&](auto kernel) { kernel<<<grid, DUST_THREADS, 0, ctx->stream>>>(p); };
        if (nl == 3) { if (stats) launch(frame_kernel<3, true>); else launch(frame_kernel<3, false>); }
        else if (nl == 4) { if (stats) launch(frame_kernel<4, true>); else launch(frame_kernel<4, false>); }
        else { if (stats) launch(frame_kernel<0, true>); else launch(frame_kernel<0, false>); }
        count_launch(ctx, "frame_kernel");
    }
    ctx->timing_on = false;
    CU_TRY(ctx, cudaGetLastError());

    if (frame->memory == DUST_MEM_HOST) {
        cudaStream_t cs = ctx->stream;
        if (slot) {                                                       // the copies leave the trace stream: frame k+1 traces while frame k downloads
            CU_TRY(ctx, cudaEventRecord(slot->traced, ctx->stream));
            CU_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, slot->traced, 0));
            cs = ctx->copy_stream;
        }
        CU_TRY(ctx, cudaMemcpyAsync(frame->radiance, scratch, out_pixels * px_bytes, cudaMemcpyDeviceToHost, cs));
        if (want_hits) {
            CU_TRY(ctx, cudaMemcpyAsync(frame->hit_instance, scratch + off_hi, npx * 4, cudaMemcpyDeviceToHost, cs));
            CU_TRY(ctx, cudaMemcpyAsync(frame->hit_leaf, scratch + off_hl, npx * 4, cudaMemcpyDeviceToHost, cs));
            CU_TRY(ctx, cudaMemcpyAsync(frame->hit_t, scratch + off_ht, npx * 4, cudaMemcpyDeviceToHost, cs));
            CU_TRY(ctx, cudaMemcpyAsync(frame->hit_voxel, scratch + off_hv, npx, cudaMemcpyDeviceToHost, cs));
        }
        if (gb) {
            void *dst[5] = { gb->albedo, gb->depth, gb->normal, gb->motion, gb->voxel_id };
            const size_t bytes[5] = { 4, 4, 4, 8, 4 };
            for (int k = 0; k < 5; k++)
                if (dst[k]) CU_TRY(ctx, cudaMemcpyAsync(dst[k], scratch + off_gb[k], npx * bytes[k], cudaMemcpyDeviceToHost, cs));
        }
        if (slot) { CU_TRY(ctx, cudaEventRecord(slot->done, cs)); return DUST_OK; }
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (stats) return fetch_counters(ctx);
    return DUST_OK;
}

int32_t dust_trace(DustCtx *ctx, const DustScene *scene, const DustCamera *camera, const DustSky *sky,
                   const DustTraceConfig *config, const DustFrame *frame)
{
    return trace_impl(ctx, scene, camera, sky, config, frame, nullptr);
}

int32_t dust_trace_async(DustCtx *ctx, const DustScene *scene, const DustCamera *camera, const DustSky *sky,
                         const DustTraceConfig *config, const DustFrame *frame, uint64_t *ticket_out)
{
    if (!ctx || !frame || !config || !ticket_out) return DUST_ERR_INVALID_ARGUMENT;
    if (frame->memory != DUST_MEM_HOST) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "dust_trace_async downloads into host buffers (DUST_MEM_HOST); device frames are asynchronous already");
    if (config->flags & (DUST_TRACE_STATS | DUST_TRACE_TIMING)) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "dust_trace_async: STATS / TIMING need the synchronous call");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if (!ctx->copy_stream) CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    DustCtx::FrameSlot &fs = ctx->slots[ctx->next_ticket % 3];
    if (!fs.traced) CU_TRY(ctx, cudaEventCreateWithFlags(&fs.traced, cudaEventDisableTiming));
    if (!fs.done) CU_TRY(ctx, cudaEventCreateWithFlags(&fs.done, cudaEventDisableTiming));
    // three frames in flight at most: the slot's previous frame must have left its staging buffer
    if (fs.busy) { CU_TRY(ctx, cudaEventSynchronize(fs.done)); fs.busy = false; }
    const int32_t rc = trace_impl(ctx, scene, camera, sky, config, frame, &fs);
    if (rc != DUST_OK) return rc;
    fs.busy = true; fs.ticket = ctx->next_ticket;
    *ticket_out = ctx->next_ticket++;
    return DUST_OK;
}

int32_t dust_frame_wait(DustCtx *ctx, uint64_t ticket)
{
    if (!ctx || !ticket || ticket >= ctx->next_ticket) return DUST_ERR_INVALID_ARGUMENT;
    DustCtx::FrameSlot &fs = ctx->slots[ticket % 3];
    if (fs.ticket != ticket || !fs.busy) return DUST_OK;                 // already waited for, or its slot was recycled (which waited)
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaEventSynchronize(fs.done));
    fs.busy = false;
    return DUST_OK;
}

// One chunk (<= 16 Mi rays) of a ray batch on `stream`: wf_cull (multi-instance scenes) + wf_trace_pt + the overflow pass.
// `counters`: 4 words (overflow count, pad, 64-bit work counter); `cand`: cn * (4 * WF_CANDS + 4) bytes of scratch when need_cull.
static cudaError_t launch_ray_chunk(DustCtx *ctx, const DustScene *scene, const DustRay *d_rays, DustHit *d_hits, uint32_t cn, cudaStream_t stream,
                                    uint32_t *counters, uint8_t *cand, bool need_cull, bool plain, bool stats, int nl, bool rough)
{
    WfParams p;
    memset(&p, 0, sizeof p);
    p.f.scene = scene->view;
    p.f.counters = ctx->d_counters;
    p.n_paths = cn;
    p.hits_out = d_hits;
    WfQueue q{};
    q.a = reinterpret_cast<float4 *>(const_cast<DustRay *>(d_rays)); q.b = q.a + 1; q.path = nullptr; q.stride = 2;
    const uint32_t grid = (cn + DUST_THREADS - 1) / DUST_THREADS;
    uint32_t *ovf_count = counters;
    unsigned long long *work = reinterpret_cast<unsigned long long *>(counters + 2);
    // kernel variants by (levels, stats | rough): DUST_RAYS_ROUGH has its own instantiations (without the STATS counters)
    const int variant = rough ? 2 : (stats ? 1 : 0);
#define DUST_BY_NL(KERNEL, ...)                                                                                  \
    do {                                                                                                         \
        if (nl == 3) { if (variant == 2) KERNEL<3, false, ##__VA_ARGS__, true> LAUNCH; else if (variant == 1) KERNEL<3, true, ##__VA_ARGS__> LAUNCH; else KERNEL<3, false, ##__VA_ARGS__> LAUNCH; } \
        else if (nl == 4) { if (variant == 2) KERNEL<4, false, ##__VA_ARGS__, true> LAUNCH; else if (variant == 1) KERNEL<4, true, ##__VA_ARGS__> LAUNCH; else KERNEL<4, false, ##__VA_ARGS__> LAUNCH; } \
        else { if (variant == 2) KERNEL<0, false, ##__VA_ARGS__, true> LAUNCH; else if (variant == 1) KERNEL<0, true, ##__VA_ARGS__> LAUNCH; else KERNEL<0, false, ##__VA_ARGS__> LAUNCH; } \
    } while (0)
    if (plain) {
#define LAUNCH <<<grid, DUST_THREADS, 0, stream>>>(p, q, nullptr, nullptr)
        DUST_BY_NL(wf_trace);
#undef LAUNCH
        count_launch(ctx);
        return cudaGetLastError();
    }
    cudaError_t le = cudaMemsetAsync(counters, 0, 16, stream);
    if (le != cudaSuccess) return le;
    if (need_cull) {
        p.cand = reinterpret_cast<uint4 *>(cand);
        p.overflow = reinterpret_cast<uint32_t *>(cand + align_up((size_t)cn * 4 * WF_CANDS, 256));
        if (stats) wf_cull<true><<<grid, DUST_THREADS, 0, stream>>>(p, q, nullptr, ovf_count);
        else wf_cull<false><<<grid, DUST_THREADS, 0, stream>>>(p, q, nullptr, ovf_count);
        count_launch(ctx);
    }
    int per_sm = 0;
    {
        // occupancy of the variant about to be launched (all variants of one NL share their launch bounds)
        if (nl == 3) le = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wf_trace_pt<3, false, true>, DUST_THREADS, 0);
        else if (nl == 4) le = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wf_trace_pt<4, false, true>, DUST_THREADS, 0);
        else le = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wf_trace_pt<0, false, true>, DUST_THREADS, 0);
    }
    const uint32_t pgrid = std::min<uint32_t>(grid, (uint32_t)(ctx->sms * std::max(per_sm, 1)));
#define LAUNCH <<<pgrid, DUST_THREADS, 0, stream>>>(p, q, nullptr, work, ovf_count)
    if (need_cull) DUST_BY_NL(wf_trace_pt, true); else DUST_BY_NL(wf_trace_pt, false);
#undef LAUNCH
    count_launch(ctx);
    if (need_cull) {                                                   // rays whose candidate list was not conclusive: one CTA per ray
        const uint32_t sgrid = std::min<uint32_t>(grid, (uint32_t)ctx->sms * 8u);
#define LAUNCH <<<sgrid, DUST_THREADS, 0, stream>>>(p, q, ovf_count, p.overflow)
        DUST_BY_NL(wf_trace_slow);
#undef LAUNCH
        count_launch(ctx);
    }
#undef DUST_BY_NL
    return le != cudaSuccess ? le : cudaGetLastError();
}

int32_t dust_trace_rays(DustCtx *ctx, const DustScene *scene, const DustRay *rays, uint64_t n, DustHit *hits,
                        uint32_t memory, uint32_t flags)
{
    if (!ctx || !scene || (n && (!rays || !hits))) return DUST_ERR_INVALID_ARGUMENT;
    if (n == 0) return DUST_OK;
    if (memory != DUST_MEM_HOST && memory != DUST_MEM_DEVICE) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "bad memory space");
    if (memory == DUST_MEM_DEVICE && (!aligned16(rays) || !aligned16(hits))) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "device ray / hit buffers must be 16-byte aligned");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const bool host = memory == DUST_MEM_HOST;
    const bool stats = (flags & DUST_TRACE_STATS) != 0;
    if (stats) CU_TRY(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof(DCounters), ctx->stream));
    const int nl = scene->n_inst == 0 || scene->min_levels == scene->max_levels ? (scene->n_inst ? (int)scene->max_levels : 3) : 0;
    if (!ctx->d_wf_counters) CU_TRY(ctx, cudaMalloc(&ctx->d_wf_counters, kWfCounterWords * sizeof(uint32_t)));

    // Ray batches are incoherent by nature: persistent lanes that claim the next ray when theirs is finished
    // (wf_trace_pt), fed by pre-culled candidate lists (wf_cull) when the scene has more than one instance.
    const bool plain = std::getenv("DUST_RAYS_PLAIN") != nullptr;                         // A/B only: one thread per ray
    const bool need_cull = !plain && scene->n_inst > 1;
    // Device buffers: chunks of <= 16 Mi rays, back to back on the context's stream.  Host buffers: chunks of 2 Mi
    // rays ping-pong over two side streams, so the H2D copy of chunk k+1 and the D2H copy of chunk k-1 overlap the
    // kernels of chunk k (both copy engines + the SMs busy; needs pinned host memory to overlap, works without).
    uint64_t chunk_max = host ? (2ull << 20) : (need_cull ? (16ull << 20) : (1ull << 30));
    if (const char *e = std::getenv("DUST_RAYS_CHUNK")) chunk_max = std::max<uint64_t>(1024, strtoull(e, nullptr, 10));   // tests: force several chunks
    const int n_slots = host ? 2 : 1;
    const size_t cand_slot = need_cull ? align_up(align_up((size_t)std::min<uint64_t>(n, chunk_max) * 4 * WF_CANDS, 256) + (size_t)std::min<uint64_t>(n, chunk_max) * 4, 256) : 0;
    if (cand_slot * n_slots > ctx->wf_scratch_bytes) {
        if (ctx->wf_scratch) { CU_TRY(ctx, cudaStreamSynchronize(ctx->stream)); CU_TRY(ctx, cudaFree(ctx->wf_scratch)); ctx->wf_scratch = nullptr; ctx->wf_scratch_bytes = 0; }
        CU_TRY(ctx, cudaMalloc(&ctx->wf_scratch, cand_slot * n_slots));
        ctx->wf_scratch_bytes = cand_slot * n_slots;
    }
    const DustRay *d_rays = rays;
    DustHit *d_hits = hits;
    cudaEvent_t ev = nullptr;
    if (host) {
        if (!ctx->ev_join) CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
        ev = ctx->ev_join;
        const size_t ray_bytes = align_up((size_t)n * sizeof(DustRay), 256);
        const int32_t rc = ensure_scratch(ctx, ray_bytes + (size_t)n * sizeof(DustHit));
        if (rc != DUST_OK) return rc;
        d_rays = reinterpret_cast<const DustRay *>(ctx->scratch);
        d_hits = reinterpret_cast<DustHit *>(static_cast<uint8_t *>(ctx->scratch) + ray_bytes);
        for (int k = 0; k < 2; k++)
            if (!ctx->side[k]) CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->side[k], cudaStreamNonBlocking));
        CU_TRY(ctx, cudaEventRecord(ev, ctx->stream));                                     // side streams start after the work already queued
        for (int k = 0; k < 2; k++) CU_TRY(ctx, cudaStreamWaitEvent(ctx->side[k], ev, 0));
    }
    cudaError_t le = cudaSuccess;
    uint32_t slot = 0;
    for (uint64_t first = 0; first < n && le == cudaSuccess; first += chunk_max, slot ^= 1u) {
        const uint32_t cn = (uint32_t)std::min<uint64_t>(chunk_max, n - first);
        cudaStream_t st = host ? ctx->side[slot] : ctx->stream;
        const uint32_t sl = host ? slot : 0u;
        if (host) le = cudaMemcpyAsync(const_cast<DustRay *>(d_rays) + first, rays + first, (size_t)cn * sizeof(DustRay), cudaMemcpyHostToDevice, st);
        if (le == cudaSuccess)
            le = launch_ray_chunk(ctx, scene, d_rays + first, d_hits + first, cn, st, ctx->d_wf_counters + 4 * sl,
                                  static_cast<uint8_t *>(ctx->wf_scratch) + cand_slot * sl, need_cull, plain, stats, nl, (flags & DUST_RAYS_ROUGH) != 0);
        if (host && le == cudaSuccess) le = cudaMemcpyAsync(hits + first, d_hits + first, (size_t)cn * sizeof(DustHit), cudaMemcpyDeviceToHost, st);
    }
    if (host) {
        for (int k = 0; k < 2; k++) {                                                      // join the side streams back into the context's stream
            if (cudaEventRecord(ev, ctx->side[k]) == cudaSuccess) cudaStreamWaitEvent(ctx->stream, ev, 0);
            cudaStreamSynchronize(ctx->side[k]);
        }
    }
    CU_TRY(ctx, le);
    CU_TRY(ctx, cudaGetLastError());
    if (host) CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (stats) return fetch_counters(ctx);
    return DUST_OK;
}

int32_t dust_resolve_shards(DustCtx *ctx, uint32_t width, uint32_t height, uint32_t shard_count, const void *gathered, void *out)
{
    if (!ctx || !gathered || !out || !width || !height || !shard_count) return DUST_ERR_INVALID_ARGUMENT;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    uint32_t tiles_per_shard = 0;
    dust_shard_layout(width, height, shard_count, &tiles_per_shard, nullptr);
    dim3 grid((width + 31) / 32, (height + 7) / 8);
    resolve_kernel<<<grid, 256, 0, ctx->stream>>>(static_cast<const float4 *>(gathered), static_cast<float4 *>(out), width, height,
                                                 shard_count, (width + 31) / 32, tiles_per_shard);
    CU_TRY(ctx, cudaGetLastError());
    count_launch(ctx);
    return DUST_OK;
}

// ---- post chain (post.cuh) ---------------------------------------------------------------------------

int32_t dust_exposure_create(DustCtx *ctx, DustExposure **out)
{
    if (!ctx || !out) return DUST_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    DustExposure *e = new (std::nothrow) DustExposure();
    if (!e) return fail(ctx, DUST_ERR_OUT_OF_MEMORY, "host allocation failed");
    if (cudaMalloc(&e->d, sizeof(DHistogram)) != cudaSuccess) { delete e; return fail(ctx, DUST_ERR_OUT_OF_MEMORY, "cudaMalloc(histogram)"); }
    const cudaError_t me = cudaMemsetAsync(e->d, 0, sizeof(DHistogram), ctx->stream);  // fill_buffer(&mut buffer, 0), auto_exposure.rs:116-118
    if (me != cudaSuccess) { cudaFree(e->d); delete e; return fail(ctx, DUST_ERR_CUDA, cudaGetErrorString(me)); }
    *out = e;
    return DUST_OK;
}

int32_t dust_exposure_destroy(DustCtx *ctx, DustExposure *exposure)
{
    if (!ctx || !exposure) return DUST_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (exposure->d) cudaFree(exposure->d);
    delete exposure;
    return DUST_OK;
}

int32_t dust_exposure_get(DustCtx *ctx, const DustExposure *exposure, float *avg_out)
{
    if (!ctx || !exposure || !avg_out) return DUST_ERR_INVALID_ARGUMENT;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaMemcpyAsync(avg_out, &exposure->d->avg, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return DUST_OK;
}

int32_t dust_exposure_set(DustCtx *ctx, DustExposure *exposure, float avg)
{
    if (!ctx || !exposure) return DUST_ERR_INVALID_ARGUMENT;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaMemcpyAsync(&exposure->d->avg, &avg, sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return DUST_OK;
}

int32_t dust_auto_exposure(DustCtx *ctx, DustExposure *exposure, const void *illuminance, uint32_t width, uint32_t height,
                           uint32_t memory, const DustExposureParams *params, float *avg_out)
{
    if (!ctx || !exposure || !illuminance || !params || !width || !height) return DUST_ERR_INVALID_ARGUMENT;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const uint64_t n = (uint64_t)width * height;
    const uint2 *img = static_cast<const uint2 *>(illuminance);
    if (memory == DUST_MEM_HOST) {
        const int32_t rc = ensure_scratch(ctx, n * 8);
        if (rc != DUST_OK) return rc;
        CU_TRY(ctx, cudaMemcpyAsync(ctx->scratch, illuminance, n * 8, cudaMemcpyHostToDevice, ctx->stream));
        img = static_cast<const uint2 *>(ctx->scratch);
    } else if (memory != DUST_MEM_DEVICE) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "bad memory space");
    else if (!aligned16(illuminance)) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "device illuminance image must be 16-byte aligned");
    const uint32_t grid = (uint32_t)std::min<uint64_t>((n / 2 + 255) / 256 + 1, (uint64_t)ctx->sms * 8);
    exposure_histogram_kernel<<<grid, 256, 0, ctx->stream>>>(img, n, params->min_log_luminance, 1.0f / params->log_luminance_range, exposure->d);
    count_launch(ctx);
    exposure_average_kernel<<<1, 256, 0, ctx->stream>>>(exposure->d, width * height, params->min_log_luminance, params->log_luminance_range,
                                                        params->time_coefficient);
    count_launch(ctx);
    CU_TRY(ctx, cudaGetLastError());
    if (avg_out) {
        CU_TRY(ctx, cudaMemcpyAsync(avg_out, &exposure->d->avg, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return DUST_OK;
}

int32_t dust_tone_map(DustCtx *ctx, const DustExposure *exposure, const void *illuminance, const uint32_t *albedo,
                      uint32_t width, uint32_t height, uint32_t memory, const DustToneMapParams *params, void *dst)
{
    if (!ctx || !exposure || !illuminance || !albedo || !params || !dst || !width || !height) return DUST_ERR_INVALID_ARGUMENT;
    if (params->dst_format != DUST_FMT_RGBA32F && params->dst_format != DUST_FMT_RGBA16F) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "bad dst_format");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const uint64_t n = (uint64_t)width * height;
    const size_t dst_bytes = n * (params->dst_format == DUST_FMT_RGBA32F ? 16 : 8);
    const uint2 *img = static_cast<const uint2 *>(illuminance);
    const uint32_t *alb = albedo;
    void *d_dst = dst;
    if (memory == DUST_MEM_HOST) {
        const size_t off_alb = align_up(n * 8, 256), off_dst = off_alb + align_up(n * 4, 256);
        const int32_t rc = ensure_scratch(ctx, off_dst + dst_bytes);
        if (rc != DUST_OK) return rc;
        uint8_t *s = static_cast<uint8_t *>(ctx->scratch);
        CU_TRY(ctx, cudaMemcpyAsync(s, illuminance, n * 8, cudaMemcpyHostToDevice, ctx->stream));
        CU_TRY(ctx, cudaMemcpyAsync(s + off_alb, albedo, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        img = reinterpret_cast<const uint2 *>(s); alb = reinterpret_cast<const uint32_t *>(s + off_alb); d_dst = s + off_dst;
    } else if (memory != DUST_MEM_DEVICE) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "bad memory space");
    else if (!aligned16(illuminance) || !aligned16(dst)) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "device illuminance / dst images must be 16-byte aligned");
    ToneMapParams tp;
    memcpy(tp.csc, params->color_space_conversion, sizeof tp.csc);
    tp.oetf = params->output_transfer_function; tp.dst_format = params->dst_format;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sms * 8);
    tone_map_kernel<<<grid, 256, 0, ctx->stream>>>(img, alb, n, exposure->d, tp, d_dst);
    count_launch(ctx);
    CU_TRY(ctx, cudaGetLastError());
    if (memory == DUST_MEM_HOST) {
        CU_TRY(ctx, cudaMemcpyAsync(dst, d_dst, dst_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return DUST_OK;
}

// ---- multi-GPU exchange step: one ncclAllGather of the compact per-rank tile buffers ----
// NCCL is bound at run time (dlopen "libnccl.so.2": the copy the host process already loaded, else the
// system one), so single-GPU users of this library carry no NCCL dependency.  The declarations below
// restate the public NCCL 2.x C API (nccl.h) for the five entry points used.

namespace {
struct NcclUniqueId { char internal[DUST_COMM_ID_BYTES]; };
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*CommGetAsyncError)(void *, int *) = nullptr;                  // optional (NCCL >= 2.4)
};
NcclApi *nccl_api(std::string &why)
{
    static NcclApi api;
    static std::string err;
    static std::once_flag once;                                        // contexts of different GPUs may be driven from different threads
    std::call_once(once, [] {
        const char *override_path = std::getenv("DUST_NCCL_LIB");
        api.lib = dlopen(override_path ? override_path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.lib) { err = std::string("dlopen libnccl.so.2: ") + dlerror(); return; }
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.lib, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.lib, "ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.lib, "ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.lib, "ncclAllGather"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.lib, "ncclGetErrorString"));
        api.CommGetAsyncError = reinterpret_cast<decltype(api.CommGetAsyncError)>(dlsym(api.lib, "ncclCommGetAsyncError"));
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GetErrorString) {
            err = "libnccl.so.2 lacks a required symbol";
            dlclose(api.lib);
            api.lib = nullptr;
        }
    });
    why = err;
    return api.lib ? &api : nullptr;
}
#define NCCL_TRY(ctx, api, expr)                                                                  \
    do {                                                                                          \
        const int r_ = (expr);                                                                    \
        if (r_ != 0) return fail(ctx, DUST_ERR_COMM, std::string(#expr) + ": " + (api)->GetErrorString(r_)); \
    } while (0)
}  // namespace

int32_t dust_comm_unique_id(uint8_t *id_out)
{
    if (!id_out) return DUST_ERR_INVALID_ARGUMENT;
    std::string why;
    NcclApi *api = nccl_api(why);
    if (!api) return DUST_ERR_COMM;
    NcclUniqueId id;
    if (api->GetUniqueId(&id) != 0) return DUST_ERR_COMM;
    memcpy(id_out, id.internal, DUST_COMM_ID_BYTES);
    return DUST_OK;
}

int32_t dust_comm_init(DustCtx *ctx, const uint8_t *id, int32_t rank, int32_t nranks)
{
    if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) return DUST_ERR_INVALID_ARGUMENT;
    if (ctx->nccl_comm) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "dust_comm_init: this context already has a communicator");
    std::string why;
    NcclApi *api = nccl_api(why);
    if (!api) return fail(ctx, DUST_ERR_COMM, why);
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    NcclUniqueId uid;
    memcpy(uid.internal, id, DUST_COMM_ID_BYTES);
    NCCL_TRY(ctx, api, api->CommInitRank(&ctx->nccl_comm, nranks, uid, rank));
    ctx->comm_rank = rank; ctx->comm_size = nranks;
    return DUST_OK;
}

int32_t dust_comm_destroy(DustCtx *ctx)
{
    if (!ctx) return DUST_ERR_INVALID_ARGUMENT;
    if (!ctx->nccl_comm) return DUST_OK;
    std::string why;
    NcclApi *api = nccl_api(why);
    if (!api) return fail(ctx, DUST_ERR_COMM, why);
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    NCCL_TRY(ctx, api, api->CommDestroy(ctx->nccl_comm));
    ctx->nccl_comm = nullptr; ctx->comm_rank = 0; ctx->comm_size = 1;
    return DUST_OK;
}

int32_t dust_allgather_framebuffer(DustCtx *ctx, const void *shard, uint64_t bytes_per_rank, void *gathered)
{
    if (!ctx || !shard || !gathered || !bytes_per_rank) return DUST_ERR_INVALID_ARGUMENT;
    if (!ctx->nccl_comm) return fail(ctx, DUST_ERR_INVALID_ARGUMENT, "dust_allgather_framebuffer: call dust_comm_init first");
    std::string why;
    NcclApi *api = nccl_api(why);
    if (!api) return fail(ctx, DUST_ERR_COMM, why);
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    NCCL_TRY(ctx, api, api->AllGather(shard, gathered, (size_t)bytes_per_rank, /* ncclUint8 */ 1, ctx->nccl_comm, ctx->stream));
    // failure detection: an asynchronous error of the communicator (a peer died, a link went down) surfaces here
    // instead of as a hang in the next stream synchronisation.  ncclInProgress (7) is the normal state of an enqueued collective.
    if (api->CommGetAsyncError) {
        int async_err = 0;
        NCCL_TRY(ctx, api, api->CommGetAsyncError(ctx->nccl_comm, &async_err));
        if (async_err != 0 && async_err != 7)
            return fail(ctx, DUST_ERR_COMM, std::string("ncclCommGetAsyncError: ") + api->GetErrorString(async_err));
    }
    return DUST_OK;
}

int32_t dust_get_launch_info(DustCtx *ctx, DustLaunchInfo *out, uint32_t capacity, uint32_t *n_out)
{
    if (!ctx || !n_out || (capacity && !out)) return DUST_ERR_INVALID_ARGUMENT;
    // n_out counts every launch of the call: a value above the records returned (capacity, or the library's own limit of
    // kMaxLaunchRecs) tells the caller the list is truncated
    *n_out = (uint32_t)ctx->launch_recs.size() + ctx->launch_dropped;
    if (ctx->launch_recs.empty()) return DUST_OK;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    DCounters prev;
    memset(&prev, 0, sizeof prev);
    cudaEvent_t before = ctx->ev_begin;
    for (uint32_t k = 0; k < ctx->launch_recs.size() && k < capacity; k++) {
        DustLaunchInfo &o = out[k];
        memset(&o, 0, sizeof o);
        strncpy(o.kernel, ctx->launch_recs[k].name, sizeof o.kernel - 1);
        CU_TRY(ctx, cudaEventElapsedTime(&o.ms, before, ctx->launch_recs[k].ev));
        before = ctx->launch_recs[k].ev;
        if (ctx->snaps_valid) {                                             // counters are cumulative over the call: differences per launch
            const DCounters &c = ctx->h_snaps[k];
            o.primary_rays = c.primary_rays - prev.primary_rays; o.bounce_rays = c.bounce_rays - prev.bounce_rays;
            o.shadow_rays = c.shadow_rays - prev.shadow_rays; o.mask_words = c.mask_words - prev.mask_words;
            o.child_ptrs = c.child_ptrs - prev.child_ptrs; o.hits_shaded = c.hits_shaded - prev.hits_shaded;
            o.mask_words_strict = c.mask_words_strict - prev.mask_words_strict;
            prev = c;
        }
    }
    return DUST_OK;
}

int32_t dust_get_stats(DustCtx *ctx, DustStats *out)
{
    if (!ctx || !out) return DUST_ERR_INVALID_ARGUMENT;
    *out = ctx->stats;
    return DUST_OK;
}

}  // extern "C"

// the radiance cache + dust_trace_gi: same translation unit (the kernels of the shared headers are defined once)
#include "dust_gi.cu"
