//! Raw bindings of `include/dust_b200.h` (ABI version 4).  One `#[repr(C)]` struct per C struct, one declaration per entry point —
//! `tests/test_rust_shim.py` checks this file against the header (every function, every struct, every field, in order).
//! NOT compiled in the build container (no Rust toolchain there); it is the file a Dust maintainer adds.
#![allow(non_camel_case_types, non_snake_case)]
use core::ffi::{c_char, c_void};

pub const DUST_ABI_VERSION: i32 = 4;

pub const DUST_OK: i32 = 0;
pub const DUST_ERR_INVALID_ARGUMENT: i32 = 1;
pub const DUST_ERR_NO_DEVICE: i32 = 2;
pub const DUST_ERR_CUDA: i32 = 3;
pub const DUST_ERR_UNSUPPORTED: i32 = 4;
pub const DUST_ERR_OUT_OF_MEMORY: i32 = 5;
pub const DUST_ERR_COMM: i32 = 6;

pub const DUST_TRACE_WRITE_HITS: u32 = 1;
pub const DUST_TRACE_STATS: u32 = 2;
pub const DUST_TRACE_PIPELINE_MEGAKERNEL: u32 = 4;
pub const DUST_TRACE_PIPELINE_WAVEFRONT: u32 = 8;
pub const DUST_TRACE_DEMODULATE: u32 = 16;
pub const DUST_TRACE_TIMING: u32 = 32;
pub const DUST_RAYS_ROUGH: u32 = 64;

pub const DUST_MEM_HOST: u32 = 0;
pub const DUST_MEM_DEVICE: u32 = 1;
pub const DUST_FMT_RGBA32F: u32 = 0;
pub const DUST_FMT_RGBA16F: u32 = 1;
pub const DUST_FMT_YCOCG16F: u32 = 2;
pub const DUST_COMM_ID_BYTES: usize = 128;

#[repr(C)] pub struct DustCtx { _p: [u8; 0] }
#[repr(C)] pub struct DustTree { _p: [u8; 0] }
#[repr(C)] pub struct DustScene { _p: [u8; 0] }
#[repr(C)] pub struct DustExposure { _p: [u8; 0] }
#[repr(C)] pub struct DustRadianceCache { _p: [u8; 0] }

/// `dust_vdb::Tree<hierarchy!(..)>` as it lies in host memory (crates/vdb/src/tree.rs:7-13, pool.rs:3-17).
#[repr(C)]
pub struct DustTreeDesc {
    pub n_levels: u32,
    pub fanout_log2: [u32; 4],
    pub root: *const c_void,
    pub root_bytes: u64,
    pub pool_chunks: [*const *const c_void; 3],
    pub n_chunks: [u32; 3],
    pub item_bytes: [u32; 3],
    pub n_items: [u32; 3],
    pub materials: *const u8,
    pub n_materials: u64,
    pub palette: *const u8,
}

/// One TLAS instance (crates/render/src/accel_struct/tlas.rs:99-117).
#[repr(C)]
pub struct DustInstance {
    pub obj2world: [f32; 12],
    pub tree: *const DustTree,
}

/// == `CameraSettings` (crates/render/src/pipeline/standard.rs:813-827, std430, 192 bytes).
#[repr(C)]
pub struct DustCamera {
    pub view_proj: [f32; 16],
    pub inverse_view_proj: [f32; 16],
    pub camera_view_col0: [f32; 3],
    pub position_x: f32,
    pub camera_view_col1: [f32; 3],
    pub position_y: f32,
    pub camera_view_col2: [f32; 3],
    pub position_z: f32,
    pub tan_half_fov: f32,
    pub far: f32,
    pub near: f32,
    pub _padding: f32,
}

#[repr(C)]
pub struct DustSkyChannel {
    pub config: [f32; 9],
    pub radiance: f32,
    pub ld_coefficient0: f32,
    pub ld_coefficient1: f32,
    pub ld_coefficient2: [f32; 4],
}

/// == `SkyModelState` (crates/render/src/pipeline/sky.rs:66-85, std430, 224 bytes).
#[repr(C)]
pub struct DustSky {
    pub r: DustSkyChannel,
    pub g: DustSkyChannel,
    pub b: DustSkyChannel,
    pub direction: [f32; 4],
    pub sunlight_intensity: [f32; 4],
}

#[repr(C)]
pub struct DustTraceConfig {
    pub width: u32,
    pub height: u32,
    pub spp: u32,
    pub max_bounces: u32,
    pub seed: u32,
    pub frame_index: u32,
    pub shard_index: u32,
    pub shard_count: u32,
    pub flags: u32,
}

#[repr(C)]
pub struct DustGBuffer {
    pub albedo: *mut u32,
    pub depth: *mut f32,
    pub normal: *mut u32,
    pub motion: *mut u16,
    pub voxel_id: *mut u32,
}

#[repr(C)]
pub struct DustFrame {
    pub radiance: *mut c_void,
    pub hit_instance: *mut u32,
    pub hit_leaf: *mut u32,
    pub hit_voxel: *mut u8,
    pub hit_t: *mut f32,
    pub memory: u32,
    pub radiance_format: u32,
    pub gbuffer: *const DustGBuffer,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct DustRay {
    pub o: [f32; 3],
    pub tmin: f32,
    pub d: [f32; 3],
    pub tmax: f32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct DustHit {
    pub t: f32,
    pub instance: u32,
    pub leaf: u32,
    pub voxel: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct DustStats {
    pub primary_rays: u64,
    pub bounce_rays: u64,
    pub shadow_rays: u64,
    pub mask_words: u64,
    pub child_ptrs: u64,
    pub hits_shaded: u64,
    pub kernel_launches: u64,
    pub mask_words_strict: u64,
}

/// == `dot_vox::Voxel` (4 x u8).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct DustVoxel {
    pub x: u8,
    pub y: u8,
    pub z: u8,
    pub i: u8,
}

#[repr(C)]
pub struct DustTreeLayout {
    pub n_levels: u32,
    pub _pad: u32,
    pub n_nodes16: u64,
    pub n_mids: u64,
    pub n_leaves: u64,
    pub n_materials: u64,
    pub off_nodes16: u64,
    pub off_blocks: u64,
    pub off_mids: u64,
    pub off_leaves: u64,
    pub off_palette_linear: u64,
    pub off_palette: u64,
    pub off_materials: u64,
    pub total_bytes: u64,
    pub bounds_min: [i32; 3],
    pub bounds_max: [i32; 3],
}

#[repr(C)]
pub struct DustExposureParams {
    pub min_log_luminance: f32,
    pub log_luminance_range: f32,
    pub time_coefficient: f32,
}

#[repr(C)]
pub struct DustToneMapParams {
    pub color_space_conversion: [f32; 9],
    pub output_transfer_function: u32,
    pub dst_format: u32,
}

#[repr(C)]
pub struct DustLaunchInfo {
    pub kernel: [c_char; 24],
    pub ms: f32,
    pub _pad: u32,
    pub primary_rays: u64,
    pub bounce_rays: u64,
    pub shadow_rays: u64,
    pub mask_words: u64,
    pub child_ptrs: u64,
    pub hits_shaded: u64,
    pub mask_words_strict: u64,
}

/// == `SpatialHashEntry` of assets/shaders/headers/layout.playout (scalar layout).
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct DustCacheEntry {
    pub fingerprint: u32,
    pub radiance: u32,
    pub last_accessed_frame: u16,
    pub sample_count: u16,
}

/// == `SurfelEntry` (vec3 position + u32 direction, 16 bytes; standard.rs:334-346 allocates 720 * 480 of them).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct DustSurfel {
    pub position: [f32; 3],
    pub direction: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct DustGiStats {
    pub near_hits: u64,
    pub far_hits: u64,
    pub far_misses: u64,
    pub cache_found: u64,
    pub surfels_scheduled: u64,
    pub surfel_rays: u64,
    pub cache_inserts: u64,
    pub insert_segments: u64,
}

extern "C" {
    pub fn dust_abi_version() -> i32;
    pub fn dust_ctx_create(device: i32, out: *mut *mut DustCtx) -> i32;
    pub fn dust_ctx_destroy(ctx: *mut DustCtx) -> i32;
    pub fn dust_ctx_set_stream(ctx: *mut DustCtx, cuda_stream: *mut c_void, external: i32) -> i32;
    pub fn dust_ctx_synchronize(ctx: *mut DustCtx) -> i32;
    pub fn dust_last_error_string(ctx: *const DustCtx) -> *const c_char;
    pub fn dust_tree_upload(ctx: *mut DustCtx, desc: *const DustTreeDesc, out: *mut *mut DustTree) -> i32;
    pub fn dust_tree_build_voxels(ctx: *mut DustCtx, voxels: *const DustVoxel, n: u64, size: *const u32, palette: *const u8,
                                  memory: u32, out: *mut *mut DustTree) -> i32;
    pub fn dust_tree_layout(tree: *const DustTree, out: *mut DustTreeLayout) -> i32;
    pub fn dust_tree_download(ctx: *mut DustCtx, tree: *const DustTree, dst: *mut c_void, capacity: u64) -> i32;
    pub fn dust_tree_destroy(ctx: *mut DustCtx, tree: *mut DustTree) -> i32;
    pub fn dust_tree_leaf_count(tree: *const DustTree, out: *mut u64) -> i32;
    pub fn dust_tree_device_bytes(tree: *const DustTree, out: *mut u64) -> i32;
    pub fn dust_scene_create(ctx: *mut DustCtx, instances: *const DustInstance, n: u32, out: *mut *mut DustScene) -> i32;
    pub fn dust_scene_destroy(ctx: *mut DustCtx, scene: *mut DustScene) -> i32;
    pub fn dust_scene_update_transforms(ctx: *mut DustCtx, scene: *mut DustScene, obj2world: *const f32, n: u32) -> i32;
    pub fn dust_scene_set_previous_transforms(ctx: *mut DustCtx, scene: *mut DustScene, obj2world_prev: *const f32, n: u32) -> i32;
    pub fn dust_trace(ctx: *mut DustCtx, scene: *const DustScene, camera: *const DustCamera, sky: *const DustSky,
                      config: *const DustTraceConfig, frame: *const DustFrame) -> i32;
    pub fn dust_trace_async(ctx: *mut DustCtx, scene: *const DustScene, camera: *const DustCamera, sky: *const DustSky,
                            config: *const DustTraceConfig, frame: *const DustFrame, ticket_out: *mut u64) -> i32;
    pub fn dust_frame_wait(ctx: *mut DustCtx, ticket: u64) -> i32;
    pub fn dust_trace_rays(ctx: *mut DustCtx, scene: *const DustScene, rays: *const DustRay, n: u64, hits: *mut DustHit,
                           memory: u32, flags: u32) -> i32;
    pub fn dust_shard_layout(width: u32, height: u32, shard_count: u32, tiles_per_shard: *mut u32, pixels_per_shard: *mut u64) -> i32;
    pub fn dust_resolve_shards(ctx: *mut DustCtx, width: u32, height: u32, shard_count: u32, gathered: *const c_void, out: *mut c_void) -> i32;
    pub fn dust_comm_unique_id(id_out: *mut u8) -> i32;
    pub fn dust_comm_init(ctx: *mut DustCtx, id: *const u8, rank: i32, nranks: i32) -> i32;
    pub fn dust_comm_destroy(ctx: *mut DustCtx) -> i32;
    pub fn dust_allgather_framebuffer(ctx: *mut DustCtx, shard: *const c_void, bytes_per_rank: u64, gathered: *mut c_void) -> i32;
    pub fn dust_exposure_create(ctx: *mut DustCtx, out: *mut *mut DustExposure) -> i32;
    pub fn dust_exposure_destroy(ctx: *mut DustCtx, exposure: *mut DustExposure) -> i32;
    pub fn dust_exposure_get(ctx: *mut DustCtx, exposure: *const DustExposure, avg_out: *mut f32) -> i32;
    pub fn dust_exposure_set(ctx: *mut DustCtx, exposure: *mut DustExposure, avg: f32) -> i32;
    pub fn dust_auto_exposure(ctx: *mut DustCtx, exposure: *mut DustExposure, illuminance: *const c_void, width: u32, height: u32,
                              memory: u32, params: *const DustExposureParams, avg_out: *mut f32) -> i32;
    pub fn dust_tone_map(ctx: *mut DustCtx, exposure: *const DustExposure, illuminance: *const c_void, albedo: *const u32,
                         width: u32, height: u32, memory: u32, params: *const DustToneMapParams, dst: *mut c_void) -> i32;
    pub fn dust_get_stats(ctx: *mut DustCtx, out: *mut DustStats) -> i32;
    pub fn dust_get_launch_info(ctx: *mut DustCtx, out: *mut DustLaunchInfo, capacity: u32, n_out: *mut u32) -> i32;
    pub fn dust_cache_create(ctx: *mut DustCtx, capacity: u32, surfel_pool_size: u32, out: *mut *mut DustRadianceCache) -> i32;
    pub fn dust_cache_destroy(ctx: *mut DustCtx, cache: *mut DustRadianceCache) -> i32;
    pub fn dust_cache_clear(ctx: *mut DustCtx, cache: *mut DustRadianceCache) -> i32;
    pub fn dust_cache_download(ctx: *mut DustCtx, cache: *const DustRadianceCache, entries_out: *mut DustCacheEntry, n_entries: u64,
                               surfels_out: *mut DustSurfel, n_surfels: u64) -> i32;
    pub fn dust_cache_upload(ctx: *mut DustCtx, cache: *mut DustRadianceCache, entries: *const DustCacheEntry, n_entries: u64,
                             surfels: *const DustSurfel, n_surfels: u64) -> i32;
    pub fn dust_trace_gi(ctx: *mut DustCtx, scene: *const DustScene, cache: *mut DustRadianceCache, camera: *const DustCamera,
                         sky: *const DustSky, config: *const DustTraceConfig, frame: *const DustFrame, stats_out: *mut DustGiStats) -> i32;
}
