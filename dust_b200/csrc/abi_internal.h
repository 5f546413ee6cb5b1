// Internal to libdust_b200.so: the objects behind the opaque handles of include/dust_b200.h and the few helpers the translation
// units of the library share (dust_abi.cu: context, trees, scenes, path tracer; dust_gi.cu: the radiance-cache GI passes).
#pragma once
#include "wavefront.cuh"
#include "post.cuh"

#include <string>
#include <vector>

using namespace dust;

struct DustCtx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t side[2] = { nullptr, nullptr };   // dust_trace_rays(DUST_MEM_HOST): copy / compute overlap across chunks
    cudaEvent_t ev_join = nullptr;                 // fork / join of the side streams (created once, lives with the context)
    std::string err;
    // scratch for DUST_MEM_HOST frames / ray batches
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    DCounters *d_counters = nullptr;
    // wavefront pipeline: path state + ray queues (grown on demand), queue / bin counters (fixed size)
    void *wf_scratch = nullptr;
    size_t wf_scratch_bytes = 0;
    uint32_t *d_wf_counters = nullptr;
    uint32_t *d_tile_lists = nullptr;        // per local tile: candidate instances of its camera frustum (wf_tile_cull)
    size_t tile_lists_words = 0;
    uint32_t *d_sun_lists = nullptr;         // per cell of the sun grid: candidate instances of the sun shadow rays (wf_sun_grid)
    int sms = 148;
    DustStats stats{};
    // DUST_TRACE_TIMING: one event after every launch of the last dust_trace call (+ a counter snapshot with DUST_TRACE_STATS)
    struct LaunchRec { const char *name; cudaEvent_t ev; };
    bool timing_on = false;
    cudaEvent_t ev_begin = nullptr;
    std::vector<LaunchRec> launch_recs;
    uint32_t launch_dropped = 0;             // launches past kMaxLaunchRecs of the last timed call (reported, not silently lost)
    std::vector<cudaEvent_t> ev_pool;
    DCounters *h_snaps = nullptr;            // pinned, kMaxLaunchRecs entries
    bool snaps_valid = false;
    // dust_trace_async: frames in flight (the reference keeps up to 3, crates/rhyolite_bevy/src/lib.rs:58): each slot owns a device
    // staging buffer; the D2H copy of frame k runs on `copy_stream` behind the event `traced` while frame k+1 traces
    struct FrameSlot { void *scratch = nullptr; size_t scratch_bytes = 0; cudaEvent_t traced = nullptr, done = nullptr; uint64_t ticket = 0; bool busy = false; };
    FrameSlot slots[3];
    cudaStream_t copy_stream = nullptr;
    uint64_t next_ticket = 1;
    // multi-GPU: NCCL communicator of the shard group (dust_comm_init); NCCL is dlopen'ed on first use
    void *nccl_comm = nullptr;
    int comm_rank = 0, comm_size = 1;
};

struct DustTree {
    uint32_t n_levels = 0;
    void *blob = nullptr;            // one device allocation
    size_t blob_bytes = 0;
    const DNode16 *n16 = nullptr;
    const uint2 *blocks = nullptr;
    const uint4 *mids = nullptr;
    const uint4 *leaves = nullptr;
    const uint8_t *materials = nullptr;
    const float4 *palette_lin = nullptr;
    const uchar4 *palette_raw = nullptr;
    uint4 *dense_root = nullptr;     // own allocation of 4096 words (device_layout.h DInstance::dense_root)
    uint2 root_blocks = { 0u, 0u };  // host copy of blocks[0]
    uint64_t n_leaves = 0, n_mids = 0, n_n16 = 0, n_materials = 0;
    int32_t bmin[3] = { 0, 0, 0 }, bmax[3] = { 0, 0, 0 };   // occupied bounds in object voxels
};

struct DustExposure {
    DHistogram *d = nullptr;
};

struct DustScene {
    DInstance *d_inst = nullptr;
    DBvhNode *d_bvh = nullptr;
    float4 *d_boxes = nullptr;       // padded world boxes, 2 float4 per instance
    float4 *d_prev = nullptr;        // last frame's object -> world rows (3 per instance)
    DSceneView view{};               // device pointers + root box, passed to the kernels by value
    uint32_t n_inst = 0, n_bvh = 0, max_levels = 3, min_levels = 3;
    std::vector<const DustTree *> trees;   // per instance, for dust_scene_update_transforms
    std::vector<float> o2w;                // current object -> world rows (12 per instance)
};

// ---- shared helpers (defined in dust_abi.cu) ----
int32_t dust_fail(DustCtx *ctx, int32_t code, const std::string &msg);
int32_t dust_ensure_scratch(DustCtx *ctx, size_t bytes);
void dust_count_launch(DustCtx *ctx, const char *name = nullptr);
void dust_sun_terms(const DustSky *sky, float sun_dir[3], float *sun_ar2, float sun_nee[3]);
inline size_t dust_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

#define CU_TRY(ctx, expr)                                                                         \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return dust_fail(ctx, e_ == cudaErrorMemoryAllocation ? DUST_ERR_OUT_OF_MEMORY : DUST_ERR_CUDA, \
                             std::string(#expr) + ": " + cudaGetErrorString(e_));                 \
    } while (0)
