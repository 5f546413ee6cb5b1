// The frame loop of the reference's examples/castle.rs, on the CUDA path, written against the C++ host mirror
// (dust_b200/csrc/host/*.hpp = the Rust crates' vdb / vox / render with the reference's names) and the C ABI:
//
//   VoxLoader::load            (crates/vox/src/loader.rs:322-415)   VoxFile::parse + per-model trees + instances
//   tlas_system per frame      (crates/render/src/accel_struct/tlas.rs:79-132)   dust_scene_update_transforms
//   StandardPipeline::render   (crates/render/src/pipeline/standard.rs:228-240)  CudaPipeline::render
//   AutoExposurePipeline / ToneMappingPipeline (examples/castle.rs:200-222)      the two post passes
//
//   castle <scene.vox> <out-prefix> [frames] [width] [height] [--device-build] [--gi]
//
// Every frame one instance (the first one, standing in for castle.rs's moving teapot, :287-292) is translated, the
// frame is traced demodulated into the img_illuminance encoding with a G-buffer, exposure adapts, the tone-mapped
// image is written as <out-prefix>_<frame>.ppm and a line with the adapted luminance is printed.  No Python, no
// oracle; without a CUDA device the first call fails with DUST_ERR_NO_DEVICE.
#include "../dust_b200/csrc/host/vox.hpp"
#include "../dust_b200/csrc/host/render.hpp"

#include <cuda_runtime_api.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <map>
#include <string>
#include <vector>

using namespace dust;

#define CHECK(ctx, expr)                                                                                   \
    do {                                                                                                   \
        const int32_t rc_ = (expr);                                                                        \
        if (rc_ != DUST_OK) { std::fprintf(stderr, "%s -> %d %s\n", #expr, rc_, dust_last_error_string(ctx)); return 1; } \
    } while (0)

int main(int argc, char **argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: castle <scene.vox> <out-prefix> [frames] [width] [height] [--device-build] [--gi]\n"); return 2; }
    const uint32_t frames = argc > 3 ? (uint32_t)std::atoi(argv[3]) : 3, W = argc > 4 ? (uint32_t)std::atoi(argv[4]) : 320,
                   H = argc > 5 ? (uint32_t)std::atoi(argv[5]) : 180;
    bool device_build = false, verbose = false;
    bool gi = false;                                                        // --gi: the reference's own GI (radiance cache + surfels) instead of path-traced bounces
    for (int i = 3; i < argc; i++) {
        device_build |= std::strcmp(argv[i], "--device-build") == 0; verbose |= std::strcmp(argv[i], "--verbose") == 0;
        gi |= std::strcmp(argv[i], "--gi") == 0;
    }
    auto note = [&](const char *what) { if (verbose) std::fprintf(stderr, "[castle] %s\n", what); };

    std::ifstream in(argv[1], std::ios::binary);
    std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    std::string err;
    auto file = vox::VoxFile::parse(bytes.data(), bytes.size(), err);
    if (!file) { std::fprintf(stderr, "parse: %s\n", err.c_str()); return 1; }

    note("parsed");
    DustCtx *ctx = nullptr;
    if (dust_ctx_create(0, &ctx) != DUST_OK) { std::fprintf(stderr, "no CUDA device: there is no CPU path\n"); return 1; }

    // one tree per distinct model (loader.rs:360-372), one instance per scene-graph shape
    const std::vector<vox::Instance> instances = file->instances();
    std::map<uint32_t, DustTree *> trees;
    for (const vox::Instance &inst : instances) {
        if (trees.count(inst.model_id)) continue;
        DustTree *t = nullptr;
        if (device_build) {
            t = vox::build_model_on_device(ctx, *file, inst.model_id, err);
            if (!t) { std::fprintf(stderr, "device build: %s\n", err.c_str()); return 1; }
        } else {
            auto geom = vox::load_model(*file, inst.model_id, err);
            if (!geom) { std::fprintf(stderr, "load_model: %s\n", err.c_str()); return 1; }
            DustTreeDesc desc;
            geom->tree.describe(desc, geom->materials.data(), geom->materials.size(), file->palette[0].data());
            CHECK(ctx, dust_tree_upload(ctx, &desc, &t));                   // host memory may be freed after return
        }
        trees[inst.model_id] = t;
    }
    note("trees uploaded");
    std::vector<DustInstance> di(instances.size());
    std::vector<float> o2w(12 * instances.size());
    for (size_t k = 0; k < instances.size(); k++) {
        std::memcpy(di[k].obj2world, instances[k].obj2world.m, sizeof di[k].obj2world);
        std::memcpy(&o2w[12 * k], instances[k].obj2world.m, 12 * sizeof(float));
        di[k].tree = trees[instances[k].model_id];
    }
    DustScene *scene = nullptr;
    CHECK(ctx, dust_scene_create(ctx, di.data(), (uint32_t)di.size(), &scene));

    note("scene created");
    // device images: illuminance (YCoCg16F), albedo (A2R10G10B10), tone-mapped output (RGBA32F)
    const size_t npx = (size_t)W * H;
    void *d_illum = nullptr, *d_albedo = nullptr, *d_ldr = nullptr;
    if (cudaMalloc(&d_illum, npx * 8) || cudaMalloc(&d_albedo, npx * 4) || cudaMalloc(&d_ldr, npx * 16)) { std::fprintf(stderr, "cudaMalloc failed\n"); return 1; }
    DustGBuffer gb{};
    gb.albedo = static_cast<uint32_t *>(d_albedo);
    DustFrame frame{};
    frame.radiance = d_illum; frame.memory = DUST_MEM_DEVICE; frame.radiance_format = DUST_FMT_YCOCG16F; frame.gbuffer = &gb;

    {   // the pipelines own device state of `ctx`: they go out of scope before the context is destroyed
    render::CudaPipeline pipeline(ctx);
    render::AutoExposurePipeline exposure(ctx);
    const float identity[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
    render::ToneMappingPipeline tone_map(ctx, identity, 1 /* sRGB OETF */);
    render::ExposureSettings exposure_settings;

    render::RenderParams params;                                            // Sunlight::default(), castle.rs sun at scroll 0
    params.spp = 2; params.max_bounces = 1; params.flags = DUST_TRACE_DEMODULATE;
    render::PinholeProjection projection;                                   // projection.rs:21-28
    const render::CameraTransform camera = render::CameraTransform::looking_at({ 122.0f, 300.61f, 54.45f }, { 0, 0, 0 }, { 0, 1, 0 });   // castle.rs:126-128

    std::vector<float> ldr(npx * 4);
    for (uint32_t f = 0; f < frames; f++) {
        if (f && !instances.empty()) {                                      // the moving instance (castle.rs:287-292)
            o2w[3] += 4.0f;
            CHECK(ctx, dust_scene_update_transforms(ctx, scene, o2w.data(), (uint32_t)instances.size()));
        }
        note("frame");
        params.rand = f * 7919u; params.frame_index = f;
        DustGiStats gs{};
        const bool ok = gi ? pipeline.render_gi(frame, W, H, scene, params, projection, camera, &gs, 1u << 20, 16384u)
                           : pipeline.render(frame, W, H, scene, params, projection, camera);
        if (!ok) { std::fprintf(stderr, "render: %s\n", dust_last_error_string(ctx)); return 1; }
        if (gi) std::printf("gi frame %u: near %llu far %llu miss %llu cache hits %llu inserts %llu surfels %llu\n", f, (unsigned long long)gs.near_hits,
                            (unsigned long long)gs.far_hits, (unsigned long long)gs.far_misses, (unsigned long long)gs.cache_found,
                            (unsigned long long)gs.cache_inserts, (unsigned long long)gs.surfel_rays);
        DustExposure *buffer = exposure.render(d_illum, W, H, exposure_settings);
        if (!buffer) { std::fprintf(stderr, "auto exposure: %s\n", dust_last_error_string(ctx)); return 1; }
        if (!tone_map.render(d_illum, static_cast<const uint32_t *>(d_albedo), d_ldr, DUST_FMT_RGBA32F, W, H, buffer)) {
            std::fprintf(stderr, "tone map: %s\n", dust_last_error_string(ctx)); return 1;
        }
        note("post chain enqueued");
        float avg = 0.0f;
        CHECK(ctx, dust_exposure_get(ctx, buffer, &avg));                   // synchronises the context's stream
        if (cudaMemcpy(ldr.data(), d_ldr, npx * 16, cudaMemcpyDeviceToHost)) { std::fprintf(stderr, "cudaMemcpy failed\n"); return 1; }
        const std::string path = std::string(argv[2]) + "_" + std::to_string(f) + ".ppm";
        std::ofstream out(path, std::ios::binary);
        out << "P6\n" << W << " " << H << "\n255\n";
        double sum = 0.0;
        for (size_t p = 0; p < npx; p++)
            for (int c = 0; c < 3; c++) {
                const float v = ldr[4 * p + c];
                sum += v;
                out.put((char)(uint8_t)(std::fmin(std::fmax(v, 0.0f), 1.0f) * 255.0f + 0.5f));
            }
        std::printf("frame %u adapted_luminance %.6g mean_ldr %.6f %s\n", f, avg, sum / (3.0 * (double)npx), path.c_str());
    }
    }
    DustStats st;
    CHECK(ctx, dust_get_stats(ctx, &st));
    std::printf("kernel_launches %llu\n", (unsigned long long)st.kernel_launches);

    cudaFree(d_illum); cudaFree(d_albedo); cudaFree(d_ldr);
    dust_scene_destroy(ctx, scene);
    for (auto &kv : trees) dust_tree_destroy(ctx, kv.second);
    dust_ctx_destroy(ctx);
    return 0;
}
