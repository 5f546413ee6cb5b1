//! Safe layer over `dust-cuda-sys`: what a Dust maintainer adds to put the B200 path behind the existing plugin surface.
//! Mirrors, name for name and argument for argument, the reference interfaces it replaces:
//!   CudaVoxGeometry::from_tree      <- VoxGeometry::from_tree            crates/vox/src/geometry.rs:55-179
//!   CudaVoxGeometry::from_voxels    <- VoxLoader::load_model             crates/vox/src/loader.rs:238-308 (built on the device)
//!   CudaScene::new / update         <- tlas_system                       crates/render/src/accel_struct/tlas.rs:79-132
//!   CudaPipeline::render            <- StandardPipeline::render          crates/render/src/pipeline/standard.rs:228-240
//!   CudaPipeline::render_gi         <- the same call with the reference's own GI (ambient occlusion + final gather + surfel pass
//!                                      over the world-space radiance cache, standard.rs:334-358, 564-789)
//! NOT compiled in the build container (no Rust toolchain).  The generic `Tree` / `Pool` accessors are the
//! ones INTEGRATION.md section 2 asks dust_vdb to make `pub`.
use dust_cuda_sys as sys;
use std::ffi::{c_void, CStr};
use std::sync::Arc;

#[derive(Debug)]
pub struct DustError { pub code: i32, pub message: String }

pub struct CudaCtx { raw: *mut sys::DustCtx }
unsafe impl Send for CudaCtx {}            // calls on one context are serialised by the owner (one recorder per frame, like the reference)

impl CudaCtx {
    pub fn new(device: i32) -> Result<Arc<Self>, DustError> {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { sys::dust_ctx_create(device, &mut raw) };
        if rc != sys::DUST_OK { return Err(DustError { code: rc, message: "dust_ctx_create (no CUDA device: there is no CPU fallback)".into() }); }
        assert_eq!(unsafe { sys::dust_abi_version() }, sys::DUST_ABI_VERSION);
        Ok(Arc::new(Self { raw }))
    }
    fn check(&self, rc: i32) -> Result<(), DustError> {
        if rc == sys::DUST_OK { return Ok(()); }
        let message = unsafe { CStr::from_ptr(sys::dust_last_error_string(self.raw)) }.to_string_lossy().into_owned();
        Err(DustError { code: rc, message })
    }
}
impl Drop for CudaCtx { fn drop(&mut self) { unsafe { sys::dust_ctx_destroy(self.raw); } } }

/// The tree + material buffers of one model on the device (replaces VoxGeometry + PaletteMaterial's buffers).
pub struct CudaVoxGeometry { ctx: Arc<CudaCtx>, tree: *mut sys::DustTree }

impl CudaVoxGeometry {
    /// `tree`: the host `dust_vox::Tree` (= Tree<hierarchy!(4,2,2)>, vox/src/lib.rs:19-20) exactly as VoxLoader::load_model built it.
    pub fn from_tree(ctx: &Arc<CudaCtx>, tree: &dust_vox::Tree, materials: &[u8], palette: &[[u8; 4]; 255]) -> Result<Self, DustError> {
        let pools = tree.pools();                                      // [LEVEL 0 = leaves, LEVEL 1 = 4^3 nodes]
        let chunks: Vec<Vec<*const c_void>> = pools.iter().map(|p| p.chunk_ptrs().iter().map(|c| *c as *const c_void).collect()).collect();
        let mut desc: sys::DustTreeDesc = unsafe { std::mem::zeroed() };
        desc.n_levels = 3;
        desc.fanout_log2 = [4, 2, 2, 0];
        desc.root = tree.root_bytes().as_ptr() as *const c_void;
        desc.root_bytes = tree.root_bytes().len() as u64;
        for (lv, p) in pools.iter().enumerate() {
            desc.pool_chunks[lv] = chunks[lv].as_ptr();
            desc.n_chunks[lv] = chunks[lv].len() as u32;
            desc.item_bytes[lv] = p.item_size() as u32;
            desc.n_items[lv] = p.top();
        }
        desc.materials = materials.as_ptr();
        desc.n_materials = materials.len() as u64;
        desc.palette = palette.as_ptr() as *const u8;
        let mut out = std::ptr::null_mut();
        ctx.check(unsafe { sys::dust_tree_upload(ctx.raw, &desc, &mut out) })?;      // host memory may be freed after return
        Ok(Self { ctx: ctx.clone(), tree: out })
    }

    /// The same device tree built from the model's voxel list by a handful of kernels (no host tree at all).
    pub fn from_voxels(ctx: &Arc<CudaCtx>, model: &dot_vox::Model, palette: &[[u8; 4]; 255]) -> Result<Self, DustError> {
        let size = [model.size.x, model.size.y, model.size.z];
        let mut out = std::ptr::null_mut();
        ctx.check(unsafe {
            sys::dust_tree_build_voxels(ctx.raw, model.voxels.as_ptr() as *const sys::DustVoxel, model.voxels.len() as u64, size.as_ptr(),
                                        palette.as_ptr() as *const u8, sys::DUST_MEM_HOST, &mut out)
        })?;
        Ok(Self { ctx: ctx.clone(), tree: out })
    }

    pub fn leaf_count(&self) -> u64 { let mut n = 0; unsafe { sys::dust_tree_leaf_count(self.tree, &mut n); } n }
}
impl Drop for CudaVoxGeometry { fn drop(&mut self) { unsafe { sys::dust_tree_destroy(self.ctx.raw, self.tree); } } }

/// The instance table + BVH (replaces the TLAS).  `instances`: (GlobalTransform matrix, geometry) per entity, as tlas_system sees them.
pub struct CudaScene { ctx: Arc<CudaCtx>, raw: *mut sys::DustScene, n: u32 }

fn rows_3x4(m: &glam::Mat4) -> [f32; 12] {
    let t = m.transpose().to_cols_array();                             // tlas.rs:100-105: transpose()[0..12] = row-major 3x4
    let mut r = [0.0f32; 12];
    r.copy_from_slice(&t[0..12]);
    r
}

impl CudaScene {
    pub fn new(ctx: &Arc<CudaCtx>, instances: &[(glam::Mat4, &CudaVoxGeometry)]) -> Result<Self, DustError> {
        let inst: Vec<sys::DustInstance> = instances.iter().map(|(m, g)| sys::DustInstance { obj2world: rows_3x4(m), tree: g.tree }).collect();
        let mut raw = std::ptr::null_mut();
        ctx.check(unsafe { sys::dust_scene_create(ctx.raw, inst.as_ptr(), inst.len() as u32, &mut raw) })?;
        Ok(Self { ctx: ctx.clone(), raw, n: inst.len() as u32 })
    }
    /// Moving instances (examples/castle.rs:287-292): the transforms in force until now become "last frame" for the motion vectors.
    pub fn update_transforms(&mut self, transforms: &[glam::Mat4]) -> Result<(), DustError> {
        assert_eq!(transforms.len() as u32, self.n);
        let flat: Vec<f32> = transforms.iter().flat_map(|m| rows_3x4(m)).collect();
        self.ctx.check(unsafe { sys::dust_scene_update_transforms(self.ctx.raw, self.raw, flat.as_ptr(), self.n) })
    }
}
impl Drop for CudaScene { fn drop(&mut self) { unsafe { sys::dust_scene_destroy(self.ctx.raw, self.raw); } } }

/// Device images of one frame: the analogue of `GBuffer<RenderImage<_>>` (standard.rs:880-917).
pub struct CudaGBuffer { pub width: u32, pub height: u32, pub frame: sys::DustFrame, pub gbuffer: Box<sys::DustGBuffer> }

pub struct RenderParams<'a> { pub sunlight: &'a dust_render::Sunlight, pub spp: u32, pub max_bounces: u32, pub rand: u32 }

pub struct CudaPipeline { ctx: Arc<CudaCtx>, frame_index: u32, cache: *mut sys::DustRadianceCache }

impl CudaPipeline {
    pub fn new(ctx: &Arc<CudaCtx>) -> Self { Self { ctx: ctx.clone(), frame_index: 0, cache: std::ptr::null_mut() } }

    fn config(&self, g: &CudaGBuffer, p: &RenderParams, flags: u32) -> sys::DustTraceConfig {
        sys::DustTraceConfig { width: g.width, height: g.height, spp: p.spp, max_bounces: p.max_bounces, seed: p.rand,
                               frame_index: self.frame_index, shard_index: 0, shard_count: 1, flags }
    }

    /// Same argument meaning as StandardPipeline::render (standard.rs:228-240); `None` = "skip this frame" (standard.rs:254-266).
    pub fn render(&mut self, gbuffer: &mut CudaGBuffer, scene: &CudaScene, params: &RenderParams,
                  camera: (&dust_render::PinholeProjection, &bevy::transform::components::GlobalTransform)) -> Option<()> {
        let cam = dust_render::CameraSettings::new(camera.0, camera.1, gbuffer.width, gbuffer.height);   // standard.rs:277-302, bytes == DustCamera
        let sky = params.sunlight.bake();                                                               // sky.rs:90-268, bytes == DustSky
        let cfg = self.config(gbuffer, params, sys::DUST_TRACE_DEMODULATE);
        let rc = unsafe { sys::dust_trace(self.ctx.raw, scene.raw, &cam as *const _ as *const sys::DustCamera,
                                          &sky as *const _ as *const sys::DustSky, &cfg, &gbuffer.frame) };
        if rc != sys::DUST_OK { return None; }
        self.frame_index += 1;                                                                          // standard.rs:252
        Some(())
    }

    /// The reference's own GI for this frame: primary + ambient-occlusion + final-gather + surfel passes over the persistent
    /// world-space radiance cache (32 Mi entries, 720 x 480 surfels: standard.rs:334-358).
    pub fn render_gi(&mut self, gbuffer: &mut CudaGBuffer, scene: &CudaScene, params: &RenderParams,
                     camera: (&dust_render::PinholeProjection, &bevy::transform::components::GlobalTransform)) -> Option<sys::DustGiStats> {
        if self.cache.is_null() {
            let rc = unsafe { sys::dust_cache_create(self.ctx.raw, 32 * 1024 * 1024, 720 * 480, &mut self.cache) };
            if rc != sys::DUST_OK { return None; }
        }
        let cam = dust_render::CameraSettings::new(camera.0, camera.1, gbuffer.width, gbuffer.height);
        let sky = params.sunlight.bake();
        let cfg = self.config(gbuffer, params, 0);
        let mut stats = sys::DustGiStats::default();
        let rc = unsafe { sys::dust_trace_gi(self.ctx.raw, scene.raw, self.cache, &cam as *const _ as *const sys::DustCamera,
                                             &sky as *const _ as *const sys::DustSky, &cfg, &gbuffer.frame, &mut stats) };
        if rc != sys::DUST_OK { return None; }
        self.frame_index += 1;
        Some(stats)
    }

    /// Frames in flight (crates/rhyolite_bevy/src/lib.rs:58): queue the frame and its download into pinned host memory, get a ticket.
    pub fn render_async(&mut self, host_frame: &sys::DustFrame, width: u32, height: u32, scene: &CudaScene, params: &RenderParams,
                        cam: &sys::DustCamera, sky: &sys::DustSky) -> Result<u64, DustError> {
        let cfg = sys::DustTraceConfig { width, height, spp: params.spp, max_bounces: params.max_bounces, seed: params.rand,
                                         frame_index: self.frame_index, shard_index: 0, shard_count: 1, flags: 0 };
        let mut ticket = 0u64;
        self.ctx.check(unsafe { sys::dust_trace_async(self.ctx.raw, scene.raw, cam, sky, &cfg, host_frame, &mut ticket) })?;
        self.frame_index += 1;
        Ok(ticket)
    }
    pub fn wait(&self, ticket: u64) -> Result<(), DustError> { self.ctx.check(unsafe { sys::dust_frame_wait(self.ctx.raw, ticket) }) }
}
impl Drop for CudaPipeline {
    fn drop(&mut self) { if !self.cache.is_null() { unsafe { sys::dust_cache_destroy(self.ctx.raw, self.cache); } } }
}
