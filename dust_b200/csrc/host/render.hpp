// Host-side mirror of the pieces of `crates/render` that feed the hot path:
//   Sunlight::bake                 crates/render/src/pipeline/sky.rs:90-268
//   PinholeProjection              crates/render/src/projection.rs:3-29
//   CameraSettings construction    crates/render/src/pipeline/standard.rs:277-302
//   StandardPipeline::render       crates/render/src/pipeline/standard.rs:228-240  -> CudaPipeline::render (path-traced bounces) /
//                                                                                      CudaPipeline::render_gi (the reference's own radiance-cache GI)
#pragma once
#include "../../../include/dust_b200.h"
#include "hosek_wilkie_xyz_tables.h"

#include <cmath>
#include <cstring>

namespace dust::render {

struct Vec3 { float x, y, z; };

// crates/render/src/pipeline/sky.rs:6-23
struct Sunlight {
    float turbidity = 1.0f;
    Vec3 albedo{ 0.2f, 0.2f, 0.2f };
    Vec3 direction{ 0.0f, 0.80114365f, -0.5984721f };

    DustSky bake() const;
};

namespace detail {

constexpr float kPi = 3.14159265358979323846f;
constexpr float kHalfPi = 1.57079632679489661923f;

inline float ipow(float x, int n) { float r = 1.0f; while (n-- > 0) r *= x; return r; }

// sky.rs:135-143 — degree-5 Bernstein blend over the six elevation control points
inline void bezier5(const float *ctrl /* [6][3] */, float s, float out[3])
{
    const float r = 1.0f - s;
    const float w[6] = { ipow(r, 5), 5.0f * ipow(r, 4) * s, 10.0f * ipow(r, 3) * ipow(s, 2),
                         10.0f * ipow(r, 2) * ipow(s, 3), 5.0f * r * ipow(s, 4), ipow(s, 5) };
    for (int c = 0; c < 3; ++c) {
        float v = w[0] * ctrl[c];
        for (int k = 1; k < 6; ++k) v = v + w[k] * ctrl[3 * k + c];
        out[c] = v;
    }
}

// sky.rs:145-227 — bilinear blend in (albedo, turbidity)
inline void albedo_turbidity_blend(const float *low, const float *high, unsigned turb_stride, float turbidity,
                                   const Vec3 &albedo, float s, float out[3])
{
    const unsigned it = static_cast<unsigned>(turbidity);
    const float rem = turbidity - static_cast<float>(it);
    const float alb[3] = { albedo.x, albedo.y, albedo.z };
    float c[3];
    bezier5(low + (it - 1) * turb_stride, s, c);
    for (int k = 0; k < 3; ++k) out[k] = (1.0f - alb[k]) * (1.0f - rem) * c[k];
    bezier5(high + (it - 1) * turb_stride, s, c);
    for (int k = 0; k < 3; ++k) out[k] += alb[k] * (1.0f - rem) * c[k];
    if (it < 10) {
        bezier5(low + it * turb_stride, s, c);
        for (int k = 0; k < 3; ++k) out[k] += (1.0f - alb[k]) * rem * c[k];
        bezier5(high + it * turb_stride, s, c);
        for (int k = 0; k < 3; ++k) out[k] += alb[k] * rem * c[k];
    }
}

// sky.rs:229-253 — piecewise cubic in elevation (45 pieces x 4 coefficients per turbidity)
inline void solar_piece(unsigned turbidity, float elevation, float out[3])
{
    unsigned pos = static_cast<unsigned>(std::pow(2.0f * elevation / kPi, 1.0f / 3.0f) * 45.0f);
    if (pos > 44u) pos = 44u;
    const float brk = ipow(static_cast<float>(pos) / 45.0f, 3) * kHalfPi;
    const float x = elevation - brk;
    const float *coefs = &HW_SOLAR_DATASET[3 * (4u * 45u * turbidity + 4u * pos)];
    float xe = 1.0f;
    out[0] = out[1] = out[2] = 0.0f;
    for (int k = 3; k >= 0; --k) {
        for (int c = 0; c < 3; ++c) out[c] += coefs[3 * k + c] * xe;
        xe *= x;
    }
}

}  // namespace detail

inline DustSky Sunlight::bake() const
{
    using namespace detail;
    DustSky s;
    std::memset(&s, 0, sizeof s);
    const float theta = std::asin(direction.y);
    const float se = std::pow(theta / kHalfPi, 1.0f / 3.0f);
    DustSkyChannel *ch[3] = { &s.r, &s.g, &s.b };
    for (int i = 0; i < 9; ++i) {
        float v[3];
        albedo_turbidity_blend(HW_CONFIG_LOW_ALBEDO + i * 18, HW_CONFIG_HIGH_ALBEDO + i * 18, 9 * 18, turbidity, albedo, se, v);
        for (int c = 0; c < 3; ++c) ch[c]->config[i] = v[c];
    }
    float rad[3];
    albedo_turbidity_blend(HW_RAD_LOW_ALBEDO, HW_RAD_HIGH_ALBEDO, 18, turbidity, albedo, se, rad);
    // sky.rs:254-268
    unsigned turb_low = static_cast<unsigned>(turbidity) - 1u;
    float turb_frac = turbidity - static_cast<float>(turb_low + 1u);
    if (turb_low == 9u) { turb_low = 8u; turb_frac = 1.0f; }
    float lo[3], hi[3];
    solar_piece(turb_low, theta, lo);
    solar_piece(turb_low + 1u, theta, hi);
    for (int c = 0; c < 3; ++c) {
        ch[c]->radiance = rad[c];
        ch[c]->ld_coefficient0 = HW_SOLAR_LD_COEFS[c];
        ch[c]->ld_coefficient1 = HW_SOLAR_LD_COEFS[3 + c];
        for (int k = 0; k < 4; ++k) ch[c]->ld_coefficient2[k] = HW_SOLAR_LD_COEFS[3 * (2 + k) + c];
        s.sunlight_intensity[c] = (1.0f - turb_frac) * lo[c] + turb_frac * hi[c];
    }
    s.direction[0] = direction.x; s.direction[1] = direction.y; s.direction[2] = direction.z; s.direction[3] = 0.0f;
    s.sunlight_intensity[3] = (0.51f * (kPi / 180.0f)) / 2.0f;
    return s;
}

// crates/render/src/projection.rs:21-28
struct PinholeProjection {
    float fov = 0.78539816339744830962f;   // FRAC_PI_4
    float near = 0.1f;
    float far = 10000.0f;
};

// GlobalTransform of the camera entity: rotation columns + translation
struct CameraTransform {
    Vec3 x_axis{ 1, 0, 0 }, y_axis{ 0, 1, 0 }, z_axis{ 0, 0, 1 };
    Vec3 translation{ 0, 0, 0 };

    // bevy Transform::looking_at (the FpsCameraBundle of examples/castle.rs:118-128 starts there)
    static CameraTransform looking_at(Vec3 eye, Vec3 target, Vec3 up)
    {
        auto sub = [](Vec3 a, Vec3 b) { return Vec3{ a.x - b.x, a.y - b.y, a.z - b.z }; };
        auto cross = [](Vec3 a, Vec3 b) { return Vec3{ a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; };
        auto norm = [](Vec3 a) { const float l = std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); return Vec3{ a.x / l, a.y / l, a.z / l }; };
        CameraTransform t;
        const Vec3 back = norm(sub(eye, target));
        const Vec3 right = norm(cross(up, back));
        t.x_axis = right; t.y_axis = cross(back, right); t.z_axis = back; t.translation = eye;
        return t;
    }
};

// standard.rs:277-302.  view_proj / inverse_view_proj are carried for layout fidelity only (the
// traced passes never read them: camera.glsl:1-16); they are filled with
// perspective_infinite_reverse_rh * view.
inline DustCamera camera_settings(const PinholeProjection &proj, const CameraTransform &xf, uint32_t width, uint32_t height)
{
    DustCamera c;
    std::memset(&c, 0, sizeof c);
    const float f = 1.0f / std::tan(0.5f * proj.fov);
    const float aspect = static_cast<float>(width) / static_cast<float>(height);
    // column-major 4x4: P = [[f/aspect,0,0,0],[0,f,0,0],[0,0,0,-1],[0,0,near,0]] (glam perspective_infinite_reverse_rh)
    const float P[16] = { f / aspect, 0, 0, 0, 0, f, 0, 0, 0, 0, 0, -1, 0, 0, proj.near, 0 };
    // view = inverse(rigid transform): R^T, -R^T t
    const Vec3 ax[3] = { xf.x_axis, xf.y_axis, xf.z_axis };
    float V[16] = { 0 };
    for (int r = 0; r < 3; ++r) {
        V[0 * 4 + r] = ax[r].x; V[1 * 4 + r] = ax[r].y; V[2 * 4 + r] = ax[r].z;
        V[3 * 4 + r] = -(ax[r].x * xf.translation.x + ax[r].y * xf.translation.y + ax[r].z * xf.translation.z);
    }
    V[15] = 1.0f;
    for (int col = 0; col < 4; ++col) for (int row = 0; row < 4; ++row) {
        float v = 0.0f;
        for (int k = 0; k < 4; ++k) v += P[k * 4 + row] * V[col * 4 + k];
        c.view_proj[col * 4 + row] = v;
    }
    // inverse_view_proj = view^-1 * P^-1; P^-1 = [[aspect/f,0,0,0],[0,1/f,0,0],[0,0,0,1/near],[0,0,-1,0]]
    const float Pi[16] = { aspect / f, 0, 0, 0, 0, 1.0f / f, 0, 0, 0, 0, 0, 1.0f / proj.near, 0, 0, -1, 0 };
    const float Vi[16] = { xf.x_axis.x, xf.x_axis.y, xf.x_axis.z, 0, xf.y_axis.x, xf.y_axis.y, xf.y_axis.z, 0,
                           xf.z_axis.x, xf.z_axis.y, xf.z_axis.z, 0, xf.translation.x, xf.translation.y, xf.translation.z, 1 };
    for (int col = 0; col < 4; ++col) for (int row = 0; row < 4; ++row) {
        float v = 0.0f;
        for (int k = 0; k < 4; ++k) v += Vi[k * 4 + row] * Pi[col * 4 + k];
        c.inverse_view_proj[col * 4 + row] = v;
    }
    c.camera_view_col0[0] = xf.x_axis.x; c.camera_view_col0[1] = xf.x_axis.y; c.camera_view_col0[2] = xf.x_axis.z;
    c.camera_view_col1[0] = xf.y_axis.x; c.camera_view_col1[1] = xf.y_axis.y; c.camera_view_col1[2] = xf.y_axis.z;
    c.camera_view_col2[0] = xf.z_axis.x; c.camera_view_col2[1] = xf.z_axis.y; c.camera_view_col2[2] = xf.z_axis.z;
    c.position_x = xf.translation.x; c.position_y = xf.translation.y; c.position_z = xf.translation.z;
    c.tan_half_fov = std::tan(proj.fov / 2.0f);
    c.near = proj.near; c.far = proj.far; c._padding = 0.0f;
    return c;
}

// The plugin-facing object: one per GPU, owns nothing but handles.  `render` has the argument
// meaning of StandardPipeline::render (standard.rs:228-240): G-buffer out, scene (the TLAS
// stand-in), per-frame params (sunlight, push constants), camera (projection + transform).
// Returns false ("skip this frame", the reference's Option::None at :254-266) only on error;
// the message is available from dust_last_error_string.
struct RenderParams {
    Sunlight sunlight;
    uint32_t spp = 1, max_bounces = 0;
    uint32_t rand = 0;            // PushConstants.rand  (standard.rs:195-209)
    uint32_t frame_index = 0;     // PushConstants.frame_index
    uint32_t flags = 0;
    uint32_t shard_index = 0, shard_count = 1;
};

class CudaPipeline {
public:
    explicit CudaPipeline(DustCtx *ctx) : ctx_(ctx) {}
    bool render(const DustFrame &gbuffer, uint32_t width, uint32_t height, const DustScene *scene,
                const RenderParams &params, const PinholeProjection &projection, const CameraTransform &camera)
    {
        const DustCamera cam = camera_settings(projection, camera, width, height);
        const DustSky sky = params.sunlight.bake();                       // standard.rs:325
        DustTraceConfig cfg{ width, height, params.spp, params.max_bounces, params.rand, params.frame_index,
                             params.shard_index, params.shard_count, params.flags };
        last_status_ = dust_trace(ctx_, scene, &cam, &sky, &cfg, &gbuffer);
        if (last_status_ == DUST_OK) ++frame_index_;                      // standard.rs:252
        return last_status_ == DUST_OK;
    }
    // The same frame with the reference's own GI (the ambient-occlusion, final-gather and surfel dispatches of StandardPipeline::render,
    // standard.rs:564-789).  The pipeline owns the persistent state — spatial_hash_buffer (32 Mi entries) and surfel_pool_buffer
    // (720 x 480) — for its lifetime, the way use_shared_state keeps them across frames (standard.rs:334-358).  `gbuffer.radiance`
    // receives img_illuminance (DUST_FMT_YCOCG16F), `gbuffer.gbuffer` the primary-hit images.
    bool render_gi(const DustFrame &gbuffer, uint32_t width, uint32_t height, const DustScene *scene, const RenderParams &params,
                   const PinholeProjection &projection, const CameraTransform &camera, DustGiStats *stats = nullptr,
                   uint32_t cache_capacity = 32u * 1024u * 1024u, uint32_t surfel_pool_size = 720u * 480u)
    {
        if (!cache_) {
            last_status_ = dust_cache_create(ctx_, cache_capacity, surfel_pool_size, &cache_);
            if (last_status_ != DUST_OK) return false;
        }
        const DustCamera cam = camera_settings(projection, camera, width, height);
        const DustSky sky = params.sunlight.bake();
        DustTraceConfig cfg{ width, height, 1, 0, params.rand, params.frame_index, 0, 1, 0 };
        last_status_ = dust_trace_gi(ctx_, scene, cache_, &cam, &sky, &cfg, &gbuffer, stats);
        if (last_status_ == DUST_OK) ++frame_index_;
        return last_status_ == DUST_OK;
    }
    ~CudaPipeline() { if (cache_) dust_cache_destroy(ctx_, cache_); }
    CudaPipeline(const CudaPipeline &) = delete;
    CudaPipeline &operator=(const CudaPipeline &) = delete;
    int32_t last_status() const { return last_status_; }
    uint32_t frames_rendered() const { return frame_index_; }

private:
    DustCtx *ctx_;
    DustRadianceCache *cache_ = nullptr;
    int32_t last_status_ = DUST_OK;
    uint32_t frame_index_ = 0;
};

// ExposureSettings (crates/render/src/pipeline/auto_exposure.rs:229-247) and the two post passes with
// the reference's names.  AutoExposurePipeline owns the Histogram buffer the way the reference keeps
// it in a shared state (auto_exposure.rs:108-118); ToneMappingPipeline reads it (tone_mapping.rs).
struct ExposureSettings {
    float min_log_luminance = -6.0f, max_log_luminance = 8.5f;
    float time_coefficient = 0.2f;
    float default_value = 0.0f, current = 0.0f;
};

class AutoExposurePipeline {
public:
    explicit AutoExposurePipeline(DustCtx *ctx) : ctx_(ctx) { status_ = dust_exposure_create(ctx, &buffer_); }
    ~AutoExposurePipeline() { if (buffer_) dust_exposure_destroy(ctx_, buffer_); }
    AutoExposurePipeline(const AutoExposurePipeline &) = delete;
    AutoExposurePipeline &operator=(const AutoExposurePipeline &) = delete;
    // illuminance: device image in DUST_FMT_YCOCG16F; enqueues both passes, returns the buffer (like the
    // reference's RenderRes<SharedDeviceState<ResidentBuffer>>) or nullptr on error
    DustExposure *render(const void *illuminance_image, uint32_t width, uint32_t height, const ExposureSettings &settings)
    {
        if (!buffer_) return nullptr;
        const DustExposureParams p{ settings.min_log_luminance, settings.max_log_luminance - settings.min_log_luminance, settings.time_coefficient };
        status_ = dust_auto_exposure(ctx_, buffer_, illuminance_image, width, height, DUST_MEM_DEVICE, &p, nullptr);
        return status_ == DUST_OK ? buffer_ : nullptr;
    }
    int32_t last_status() const { return status_; }

private:
    DustCtx *ctx_;
    DustExposure *buffer_ = nullptr;
    int32_t status_ = DUST_OK;
};

class ToneMappingPipeline {
public:
    // color_space_conversion / transfer function: the specialisation constants the reference derives from
    // the swapchain's colour space (tone_map.comp:12-23)
    ToneMappingPipeline(DustCtx *ctx, const float color_space_conversion[9], uint32_t output_transfer_function) : ctx_(ctx)
    {
        for (int i = 0; i < 9; ++i) params_.color_space_conversion[i] = color_space_conversion[i];
        params_.output_transfer_function = output_transfer_function;
        params_.dst_format = DUST_FMT_RGBA16F;
    }
    bool render(const void *src_illuminance, const uint32_t *albedo, void *dst, uint32_t dst_format, uint32_t width, uint32_t height,
                const DustExposure *exposure)
    {
        params_.dst_format = dst_format;
        status_ = dust_tone_map(ctx_, exposure, src_illuminance, albedo, width, height, DUST_MEM_DEVICE, &params_, dst);
        return status_ == DUST_OK;
    }
    int32_t last_status() const { return status_; }

private:
    DustCtx *ctx_;
    DustToneMapParams params_{};
    int32_t status_ = DUST_OK;
};

}  // namespace dust::render
