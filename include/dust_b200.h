/*
 * dust_b200 — C ABI of the B200-native replacement for Dust's ray-traversal hot path.
 *
 * This is the drop-in boundary: the reference (dust-engine/dust @ 7ea772b) has no FFI of its own;
 * the path is entered through Rust traits and one method,
 *     StandardPipeline::render(&mut self, gbuffer, tlas, params, camera)
 *                                         crates/render/src/pipeline/standard.rs:228-240
 * fed by VoxGeometry::from_tree (crates/vox/src/geometry.rs:55-179), PaletteMaterial::parameters
 * (crates/vox/src/material.rs:106-118) and tlas_system (crates/render/src/accel_struct/tlas.rs:79-132).
 * Each entry point below names the reference interface it replaces.  INTEGRATION.md shows the
 * Rust `extern "C"` binding a maintainer would add.
 *
 * Conventions: plain pointers and sizes only; every call returns an int32 status (DUST_OK = 0),
 * never throws or aborts; out-params by pointer; caller owns host memory, the library owns
 * device memory behind opaque handles; one DustCtx per GPU, calls on a ctx serialised by the
 * caller (the reference has one recorder thread per frame).  There is NO CPU fallback: without a
 * CUDA device dust_ctx_create fails with DUST_ERR_NO_DEVICE.
 *
 * Alignment of DUST_MEM_DEVICE buffers: the kernels read rays, hit records and half-float images with 128-bit
 * loads, so device pointers passed to dust_trace_rays (rays, hits), dust_auto_exposure / dust_tone_map
 * (illuminance, dst) and DustFrame.radiance must be 16-byte aligned (cudaMalloc and torch allocations are; a
 * sub-view at an odd element offset may not be).  A misaligned pointer is rejected with
 * DUST_ERR_INVALID_ARGUMENT.  Host buffers (DUST_MEM_HOST) have no alignment requirement.
 */
#ifndef DUST_B200_H
#define DUST_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DUST_ABI_VERSION 4

enum {
    DUST_OK = 0,
    DUST_ERR_INVALID_ARGUMENT = 1,
    DUST_ERR_NO_DEVICE = 2,
    DUST_ERR_CUDA = 3,
    DUST_ERR_UNSUPPORTED = 4,     /* e.g. a hierarchy other than (4,2,2) / (4,4,2,2) */
    DUST_ERR_OUT_OF_MEMORY = 5,
    DUST_ERR_COMM = 6             /* NCCL missing or a collective failed (dust_last_error_string has the NCCL text) */
};

typedef struct DustCtx DustCtx;
typedef struct DustTree DustTree;
typedef struct DustScene DustScene;

/*
 * A `dust_vdb::Tree<hierarchy!(..)>` exactly as it lies in host memory
 * (crates/vdb/src/tree.rs:7-13): the inline root node and, per non-root LEVEL (0 = leaf), the
 * chunk base pointers of `Pool` (crates/vdb/src/pool.rs:3-17; chunks of 1024 items, item `ptr`
 * lives at chunks[ptr >> 10] + (ptr & 1023) * item_bytes).  Node byte layouts are the
 * #[repr(C)] ones: InternalNode {child_mask: [u64; 8^f/64], child_ptrs: [u32; 8^f]}
 * (node/internal.rs:22-32), LeafNode {occupancy: u64, active: u64, material_ptr: u32, pad}
 * = 24 bytes (node/leaf.rs:13-25).  `materials` / `palette` are the two other buffers the
 * reference puts in a hit-group record (crates/vox/src/material.rs:30-41).
 */
typedef struct DustTreeDesc {
    uint32_t n_levels;                 /* 3 for hierarchy!(4,2,2); leaf level included */
    uint32_t fanout_log2[4];           /* root first, as written in hierarchy!() */
    const void *root;
    uint64_t root_bytes;
    const void *const *pool_chunks[3]; /* [LEVEL] -> n_chunks[LEVEL] chunk base pointers */
    uint32_t n_chunks[3];
    uint32_t item_bytes[3];
    uint32_t n_items[3];               /* Pool::top — slots ever handed out */
    const uint8_t *materials;          /* 1 palette index per voxel (loader.rs:274-288) */
    uint64_t n_materials;
    const uint8_t *palette;            /* 255 x RGBA8 (loader.rs:208-218) */
} DustTreeDesc;

/* One TLAS instance: tlas.rs:99-117 (object->world 3x4 row-major) + the BLAS/SBT reference */
typedef struct DustInstance {
    float obj2world[12];
    const DustTree *tree;
} DustInstance;

/* CameraSettings, crates/render/src/pipeline/standard.rs:813-827 (std430, 192 bytes).
 * Only the view columns, position, tan_half_fov, near and far are read
 * (assets/shaders/headers/camera.glsl:1-16, primary/primary.rgen:16-19). */
typedef struct DustCamera {
    float view_proj[16];
    float inverse_view_proj[16];
    float camera_view_col0[3]; float position_x;
    float camera_view_col1[3]; float position_y;
    float camera_view_col2[3]; float position_z;
    float tan_half_fov, far, near, _padding;
} DustCamera;

/* SkyModelState, crates/render/src/pipeline/sky.rs:66-85 (std430, 224 bytes) — the output of
 * Sunlight::bake(), uploaded as the `sunlight_config` uniform (headers/layout.playout:35-51). */
typedef struct DustSkyChannel {
    float config[9];
    float radiance;
    float ld_coefficient0, ld_coefficient1;
    float ld_coefficient2[4];
} DustSkyChannel;
typedef struct DustSky {
    DustSkyChannel r, g, b;
    float direction[4];
    float sunlight_intensity[4];
} DustSky;

enum {
    DUST_TRACE_WRITE_HITS = 1u,   /* fill the hit_* arrays of DustFrame (sample-0 primary ray) */
    DUST_TRACE_STATS      = 2u,   /* count rays / visited words; read back with dust_get_stats */
    /* pipeline override (default: chosen per frame).  Results are bit-identical either way. */
    DUST_TRACE_PIPELINE_MEGAKERNEL = 4u,   /* one launch, a path stays in its thread */
    DUST_TRACE_PIPELINE_WAVEFRONT  = 8u,   /* rays re-ordered between path vertices (frames with bounces) */
    /* Leave the primary-hit albedo factor out of the radiance (the reference's img_illuminance holds the
     * incoming light; tone_map.comp:192-203 multiplies DustGBuffer.albedo back in): radiance for
     * dust_tone_map.  Default (flag clear): the product, i.e. tone_map's `mapped` before exposure. */
    DUST_TRACE_DEMODULATE = 16u,
    /* Measurement: drop a CUDA event behind every kernel launch of this call (and, together with DUST_TRACE_STATS,
     * a snapshot of the counters), readable with dust_get_launch_info: per-kernel device times of the live run. */
    DUST_TRACE_TIMING = 32u,
    /* dust_trace_rays only: block-granular leaves — the reference's far-field intersection shader
     * (assets/shaders/final_gather/rough.rint:46-58, the rays of final_gather.rgen:38-51): every non-empty leaf is a solid
     * 4x4x4 box, hit where the ray enters it; DustHit.voxel = 0, DustHit.leaf = the block (gl_PrimitiveID).  First piece of
     * SURVEY.md 8f rank 2 (the far-field gather); the radiance cache those hits look up is not built. */
    DUST_RAYS_ROUGH = 64u
};

/* Replaces the push constants + launch size of the four vkCmdTraceRaysKHR dispatches
 * (standard.rs:447-463, 477-490).  Shards: the frame is cut into 32x32 pixel tiles, tile t
 * (row-major) belongs to shard t % shard_count and is that shard's local tile t / shard_count. */
typedef struct DustTraceConfig {
    uint32_t width, height;
    uint32_t spp;
    uint32_t max_bounces;
    uint32_t seed, frame_index;
    uint32_t shard_index, shard_count;   /* 0,1 = whole frame, row-major output */
    uint32_t flags;
} DustTraceConfig;

enum { DUST_MEM_HOST = 0, DUST_MEM_DEVICE = 1 };
/* radiance formats: ACEScg RGBA as float / half, or the reference's img_illuminance encoding
 * REBLUR_FrontEnd_PackRadianceAndNormHitDist (headers/nrd.glsl:99-147): half (Y, Co, Cg, 1) */
enum { DUST_FMT_RGBA32F = 0, DUST_FMT_RGBA16F = 1, DUST_FMT_YCOCG16F = 2 };

/* The G-buffer images of the primary hit that hit.rchit:57-94 / miss.rmiss:13-16 write for the downstream
 * consumers (NRD, tone map; formats: GBuffer, standard.rs:880-917), for sample 0 of every pixel.  Row-major
 * width*height, same memory space as the frame, shard_count == 1 only; NULL members are skipped.
 * On a miss albedo = (1,1,1,1), depth = +inf, motion = 0 and — like the reference — normal and voxel_id are
 * left untouched. */
typedef struct DustGBuffer {
    uint32_t *albedo;          /* A2R10G10B10_UNORM_PACK32: vec4(palette[..].rgb / 255, 1)                  hit.rchit:71-76 */
    float *depth;              /* R32_SFLOAT: gl_HitTEXT                                                    hit.rchit:81 */
    uint32_t *normal;          /* A2R10G10B10_UNORM_PACK32: NRD_FrontEnd_PackNormalAndRoughness(normalWorld, 1.0, palette_index)  :83 */
    uint16_t *motion;          /* R16G16B16A16_SFLOAT: last-frame world position - world position, a = 0     hit.rchit:90-94 */
    uint32_t *voxel_id;        /* R32_UINT: voxel << 24 | palette_index << 16 | instance & 0xFFFF            hit.rchit:87 */
} DustGBuffer;

/* Replaces the GBuffer images the reference writes (standard.rs:880-917).  With shard_count == 1
 * every array is row-major width*height.  With shard_count > 1 `radiance` is the compact
 * per-shard tile buffer (dust_shard_layout) and the hit arrays must be NULL. */
typedef struct DustFrame {
    void *radiance;            /* radiance_format: RGBA32F / RGBA16F (ACEScg rgb, a = 1) or YCOCG16F */
    uint32_t *hit_instance;    /* 0xFFFFFFFF = miss */
    uint32_t *hit_leaf;        /* index in Tree::iter_leaf order == gl_PrimitiveID (geometry.rs:68) */
    uint8_t *hit_voxel;        /* x<<4 | y<<2 | z inside the leaf (hit.rint:30-32) */
    float *hit_t;              /* gl_HitTEXT; +inf on miss (miss.rmiss:15) */
    uint32_t memory;           /* DUST_MEM_HOST or DUST_MEM_DEVICE */
    uint32_t radiance_format;  /* DUST_FMT_* */
    const DustGBuffer *gbuffer; /* optional (NULL: none) */
} DustFrame;

typedef struct DustRay { float o[3]; float tmin; float d[3]; float tmax; } DustRay;       /* 32 B */
typedef struct DustHit { float t; uint32_t instance, leaf, voxel; } DustHit;              /* 16 B */

typedef struct DustStats {
    uint64_t primary_rays, bounce_rays, shadow_rays;
    uint64_t mask_words;       /* W: 64-bit mask words the traversal had to read */
    uint64_t child_ptrs;       /* P: child links followed */
    uint64_t hits_shaded;      /* H */
    uint64_t kernel_launches;  /* launches of this library's kernels since ctx creation */
    uint64_t mask_words_strict; /* W by the letter of SURVEY.md 8d: node words tested + mid nodes entered + leaves entered;
                                 * mask_words also counts the block-mask words and the words of a node already in registers */
} DustStats;

int32_t dust_abi_version(void);

/* Context = one CUDA device + one stream.  Replaces the device/queue resources of
 * crates/rhyolite_bevy/src/lib.rs:143-191 for this path. */
int32_t dust_ctx_create(int32_t device, DustCtx **out);
int32_t dust_ctx_destroy(DustCtx *ctx);
/* external != 0: run on the caller's cudaStream_t `cuda_stream` (e.g. torch's current stream; the
 * legacy default stream is the NULL handle, hence the explicit flag).  external == 0: back to the
 * context's own stream. */
int32_t dust_ctx_set_stream(DustCtx *ctx, void *cuda_stream, int32_t external);
int32_t dust_ctx_synchronize(DustCtx *ctx);
const char *dust_last_error_string(const DustCtx *ctx);

/* Replaces VoxGeometry::from_tree + the material/palette uploads (geometry.rs:55-179,
 * loader.rs:208-236, 274-288): the tree is re-laid-out for the GPU and uploaded once. */
int32_t dust_tree_upload(DustCtx *ctx, const DustTreeDesc *desc, DustTree **out);
/* Device-side build of ONE .vox model: replaces VoxLoader::load_model's per-voxel
 * Tree::set_value + ModelIndexCollector (loader.rs:238-308, collector.rs:2-87) + VoxGeometry::from_tree +
 * dust_tree_upload by a handful of kernels over the voxel list.  `voxels` = the model's dot_vox voxels as the loader
 * sees them (.vox axes, i = palette entry), `size` = dot_vox model size {x, y, z} (each <= 256), `palette` = 255 x
 * RGBA8.  The result is section for section the blob dust_tree_upload produces for the same model (hierarchy!(4,2,2)). */
typedef struct DustVoxel { uint8_t x, y, z, i; } DustVoxel;
int32_t dust_tree_build_voxels(DustCtx *ctx, const DustVoxel *voxels, uint64_t n, const uint32_t *size, const uint8_t *palette,
                               uint32_t memory, DustTree **out);
/* Inspection: section offsets / counts of a tree's device blob, and a copy of it (tests compare both build paths).
 * Besides the blob every tree owns a derived 64 KB look-up table of its root node's 4096 cells (built on the device from the
 * blob; the kernels descend through it).  It is not part of the blob, dust_tree_layout and dust_tree_download do not show it;
 * dust_tree_device_bytes counts it. */
typedef struct DustTreeLayout {
    uint32_t n_levels, _pad;
    uint64_t n_nodes16, n_mids, n_leaves, n_materials;
    uint64_t off_nodes16, off_blocks, off_mids, off_leaves, off_palette_linear, off_palette, off_materials, total_bytes;
    int32_t bounds_min[3], bounds_max[3];     /* occupied bounds in object voxels (multiples of 4) */
} DustTreeLayout;
int32_t dust_tree_layout(const DustTree *tree, DustTreeLayout *out);
int32_t dust_tree_download(DustCtx *ctx, const DustTree *tree, void *dst, uint64_t capacity);
int32_t dust_tree_destroy(DustCtx *ctx, DustTree *tree);
int32_t dust_tree_leaf_count(const DustTree *tree, uint64_t *out);
int32_t dust_tree_device_bytes(const DustTree *tree, uint64_t *out);

/* Replaces BLAS build + tlas_system + TLASStore::accel_struct (blas.rs:125-230, tlas.rs:37-132) */
int32_t dust_scene_create(DustCtx *ctx, const DustInstance *instances, uint32_t n, DustScene **out);
int32_t dust_scene_destroy(DustCtx *ctx, DustScene *scene);
/* Last frame's object->world matrices (n x 12 floats, row-major 3x4), the `instances[]` buffer of
 * hit.rchit:92 (PreviousFrameGlobalTransform, standard.rs:845-854); used only for DustGBuffer.motion.
 * Until called, last frame == this frame. */
/* Moving instances (examples/castle.rs:287-292 moves the teapot every frame; tlas_system rebuilds the TLAS from
 * the entities' GlobalTransforms each frame, tlas.rs:79-132): new object->world matrices (n x 12 floats) for the
 * SAME trees.  Instance records and the instance BVH are rebuilt on the host and re-uploaded, stream-ordered
 * after the frames in flight; the transforms in force until now become "last frame" for DustGBuffer.motion. */
int32_t dust_scene_update_transforms(DustCtx *ctx, DustScene *scene, const float *obj2world, uint32_t n);
int32_t dust_scene_set_previous_transforms(DustCtx *ctx, DustScene *scene, const float *obj2world_prev, uint32_t n);

/* Replaces StandardPipeline::render (standard.rs:228-810) for the primary + bounce passes. */
int32_t dust_trace(DustCtx *ctx, const DustScene *scene, const DustCamera *camera, const DustSky *sky,
                   const DustTraceConfig *config, const DustFrame *frame);

/* Frames in flight.  The reference submits a frame and moves on (up to 3 in flight behind fences:
 * crates/rhyolite_bevy/src/lib.rs:58, queue.rs:222-241).  dust_trace_async is dust_trace for a DUST_MEM_HOST frame without the
 * final wait: the kernels are queued on the context's stream, the device -> host copies follow on a second stream, and the call
 * returns a ticket; the host buffers of `frame` (pinned memory, or the copies serialise) belong to the library until
 * dust_frame_wait(ticket) returns.  The download of frame k overlaps the kernels of frame k+1.  At most 3 tickets are
 * outstanding: the 4th call first waits for the oldest.  DUST_TRACE_STATS / DUST_TRACE_TIMING are not accepted here. */
int32_t dust_trace_async(DustCtx *ctx, const DustScene *scene, const DustCamera *camera, const DustSky *sky,
                         const DustTraceConfig *config, const DustFrame *frame, uint64_t *ticket_out);
int32_t dust_frame_wait(DustCtx *ctx, uint64_t ticket);

/* Explicit ray batch, closest hit (BASELINE.json config 5).  `memory` = DUST_MEM_*. */
int32_t dust_trace_rays(DustCtx *ctx, const DustScene *scene, const DustRay *rays, uint64_t n,
                        DustHit *hits, uint32_t memory, uint32_t flags);

/* Sharded frames: sizes of the compact per-shard buffer, and the un-permute that follows the
 * all-gather (gathered = shard_count consecutive shard buffers; out = row-major width*height).
 * Both device pointers, RGBA32F. */
int32_t dust_shard_layout(uint32_t width, uint32_t height, uint32_t shard_count,
                          uint32_t *tiles_per_shard, uint64_t *pixels_per_shard);
int32_t dust_resolve_shards(DustCtx *ctx, uint32_t width, uint32_t height, uint32_t shard_count,
                            const void *gathered, void *out);

/* Multi-GPU exchange step (the reference renders on one GPU): one process and one DustCtx
 * per GPU, every rank traces its shard (DustTraceConfig.shard_index / shard_count) into a compact device
 * buffer, ONE all-gather over NVLink/NVSwitch (ncclAllGather on the context's stream) concatenates the
 * shard buffers in rank order, dust_resolve_shards un-permutes.  Rank 0 obtains the 128-byte id with
 * dust_comm_unique_id and hands it to the other ranks over the host's own channel; dust_comm_init is
 * collective (all ranks call it).  NCCL is bound with dlopen("libnccl.so.2") at the first call. */
#define DUST_COMM_ID_BYTES 128
int32_t dust_comm_unique_id(uint8_t *id_out);
int32_t dust_comm_init(DustCtx *ctx, const uint8_t *id, int32_t rank, int32_t nranks);
int32_t dust_comm_destroy(DustCtx *ctx);
/* shard, gathered: device pointers; gathered holds nranks * bytes_per_rank bytes */
int32_t dust_allgather_framebuffer(DustCtx *ctx, const void *shard, uint64_t bytes_per_rank, void *gathered);

/* ---- post chain: the two compute passes the reference runs on the traced frame ----
 * DustExposure = the reference's Histogram buffer {uint histogram[256]; float avg;} (auto_exposure.playout),
 * zero-filled at creation like auto_exposure.rs:116-118.  dust_auto_exposure = auto_exposure.comp +
 * auto_exposure_avg.comp on an illuminance image in DUST_FMT_YCOCG16F (AutoExposurePipeline::render,
 * crates/render/src/pipeline/auto_exposure.rs:96-224); dust_tone_map = tone_map.comp (ToneMappingPipeline::render,
 * crates/render/src/pipeline/tone_mapping.rs).  `memory` says where the images live; `avg_out` (may be NULL) is a
 * host float that receives the adapted luminance — passing it makes the call synchronous. */
typedef struct DustExposure DustExposure;
typedef struct DustExposureParams {     /* ExposureSettings, auto_exposure.rs:229-247 (defaults -6.0, 8.5 - -6.0, 0.2) */
    float min_log_luminance;
    float log_luminance_range;          /* max_log_luminance - min_log_luminance */
    float time_coefficient;
} DustExposureParams;
typedef struct DustToneMapParams {
    float color_space_conversion[9];    /* COLOR_SPACE_CONVERSION_0..8 (tone_map.comp:15-23), column-major mat3 */
    uint32_t output_transfer_function;  /* OUTPUT_TRANSFER_FUNCTION 0..8 (tone_map.comp:12, 98-167) */
    uint32_t dst_format;                /* DUST_FMT_RGBA32F or DUST_FMT_RGBA16F */
} DustToneMapParams;
int32_t dust_exposure_create(DustCtx *ctx, DustExposure **out);
int32_t dust_exposure_destroy(DustCtx *ctx, DustExposure *exposure);
/* read / seed the adapted luminance (Histogram.avg); both synchronise the context's stream */
int32_t dust_exposure_get(DustCtx *ctx, const DustExposure *exposure, float *avg_out);
int32_t dust_exposure_set(DustCtx *ctx, DustExposure *exposure, float avg);
int32_t dust_auto_exposure(DustCtx *ctx, DustExposure *exposure, const void *illuminance, uint32_t width, uint32_t height,
                           uint32_t memory, const DustExposureParams *params, float *avg_out);
/* albedo: DustGBuffer.albedo (A2R10G10B10_UNORM_PACK32); dst: width*height pixels of dst_format */
int32_t dust_tone_map(DustCtx *ctx, const DustExposure *exposure, const void *illuminance, const uint32_t *albedo,
                      uint32_t width, uint32_t height, uint32_t memory, const DustToneMapParams *params, void *dst);

/* ---- the reference's own global illumination ------------------------------------------------------------
 * StandardPipeline::render's four ray-tracing dispatches for one frame (crates/render/src/pipeline/standard.rs:477-789) with their
 * persistent state (standard.rs:334-358): the world-space radiance cache `spatial_hash` (SpatialHashEntry, 3-probe open addressing,
 * LogLuv radiance, capped running mean, LRU eviction: assets/shaders/headers/spatial_hash.glsl:128-220) and the `surfel_pool`.
 *   primary            primary/{primary.rgen, hit.rchit, miss.rmiss}                      -> G-buffer, illuminance = 0
 *   ambient occlusion  final_gather/ambient_occlusion.{rgen, rint, rchit, rmiss}, nee.rmiss -> sun term + the near segment [0.1, 8]
 *   final gather       final_gather/final_gather.{rgen, rchit, rmiss}, rough.rint          -> the far segment [8, far]: block-granular hit,
 *                                                                                            cache look-up, surfel scheduling, sky on a miss
 *   surfel             surfel/{surfel.rgen, surfel.rchit, surfel.rmiss, nee.rmiss}         -> one ray per pool entry, cache inserts
 * The reference's invocations race on the cache and the pool; this library runs ONE of the admitted interleavings, the same on every
 * run: look-ups of a pass see the table as the previous pass left it, the inserts of the surfel pass are applied in ascending surfel
 * index (sorted by probe window, independent windows in parallel), the highest pixel that maps to a pool slot wins.  Noise comes from
 * pcg streams seeded by DustTraceConfig.seed / frame_index instead of the STBN textures (Git-LFS pointers in the reference).
 *
 * dust_trace_gi: `config` = width, height, seed (push_constants.rand), frame_index (spp, max_bounces, shards, flags are ignored).
 * `frame->radiance` must be DUST_FMT_YCOCG16F: the reference's img_illuminance (YCoCg + normHitDist, half floats) for pixels with a
 * primary hit and img_illuminance_denoised = pack((sky + sun) / 3.14, 100000) for the others; `frame->gbuffer` (optional) receives
 * the primary-hit images; hit arrays are not written.  `stats_out` may be NULL (passing it synchronises). */
typedef struct DustRadianceCache DustRadianceCache;
typedef struct DustCacheEntry {          /* SpatialHashEntry, headers/layout.playout (scalar layout, 12 bytes) */
    uint32_t fingerprint;                /* 0 = empty */
    uint32_t radiance;                   /* LogLuv: 14 bits log2 Y, 9 + 9 bits chroma */
    uint16_t last_accessed_frame;
    uint16_t sample_count;
} DustCacheEntry;
typedef struct DustSurfel {              /* SurfelEntry, headers/layout.playout (16 bytes) */
    float position[3];                   /* world-space centre of the 4x4x4 block */
    uint32_t direction;                  /* face 0..5; >= 6: empty slot (the pool starts filled with 0xFF bytes, standard.rs:347-349) */
} DustSurfel;
typedef struct DustGiStats {
    uint64_t near_hits;                  /* pixels resolved by the ambient-occlusion ray */
    uint64_t far_hits, far_misses;       /* final-gather rays */
    uint64_t cache_found;                /* look-ups (final gather + surfel pass) that found their key */
    uint64_t surfels_scheduled;          /* pool write requests of this frame */
    uint64_t surfel_rays;                /* valid pool entries traced */
    uint64_t cache_inserts;
    uint64_t insert_segments;            /* independent probe-window groups the inserts were applied in */
} DustGiStats;
/* capacity: SpatialHashCapacity (reference: 32 * 1024 * 1024, spatial_hash.glsl:1); surfel_pool_size: SurfelPoolSize (720 * 480) */
int32_t dust_cache_create(DustCtx *ctx, uint32_t capacity, uint32_t surfel_pool_size, DustRadianceCache **out);
int32_t dust_cache_destroy(DustCtx *ctx, DustRadianceCache *cache);
int32_t dust_cache_clear(DustCtx *ctx, DustRadianceCache *cache);
/* host copies of the state (tests compare it with the CPU checker): n_entries <= capacity + 2 entries, n_surfels <= pool size */
int32_t dust_cache_download(DustCtx *ctx, const DustRadianceCache *cache, DustCacheEntry *entries_out, uint64_t n_entries,
                            DustSurfel *surfels_out, uint64_t n_surfels);
int32_t dust_cache_upload(DustCtx *ctx, DustRadianceCache *cache, const DustCacheEntry *entries, uint64_t n_entries,
                          const DustSurfel *surfels, uint64_t n_surfels);
int32_t dust_trace_gi(DustCtx *ctx, const DustScene *scene, DustRadianceCache *cache, const DustCamera *camera, const DustSky *sky,
                      const DustTraceConfig *config, const DustFrame *frame, DustGiStats *stats_out);

int32_t dust_get_stats(DustCtx *ctx, DustStats *out);
/* The launches of the last dust_trace call made with DUST_TRACE_TIMING, in order: kernel name, device time between the
 * event behind the previous launch and the one behind this launch, and (when that call also had DUST_TRACE_STATS) what
 * this launch added to the counters.  Synchronises the context's stream.  *n_out = number of launches of that call;
 * at most min(capacity, 1024) records are returned, so *n_out above that means the list is truncated. */
typedef struct DustLaunchInfo {
    char kernel[24];
    float ms;
    uint32_t _pad;
    uint64_t primary_rays, bounce_rays, shadow_rays, mask_words, child_ptrs, hits_shaded, mask_words_strict;
} DustLaunchInfo;
int32_t dust_get_launch_info(DustCtx *ctx, DustLaunchInfo *out, uint32_t capacity, uint32_t *n_out);

#ifdef __cplusplus
}
#endif
#endif
