"""Thin object wrappers over the C ABI (include/dust_b200.h): Context / DeviceTree / Scene and the
two trace entry points.  Host buffers are numpy arrays, device buffers are torch CUDA tensors
(torch is only the allocator / stream / collective plumbing; every kernel is ours).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import (PIPELINE_FLAGS, DUST_FMT_RGBA16F, DUST_FMT_RGBA32F, DUST_MEM_DEVICE, DUST_MEM_HOST, DUST_TRACE_STATS,
                   DUST_TRACE_WRITE_HITS, DustFrame, DustGBuffer, DustInstance, DustStats, DustTraceConfig, check)

MISS = 0xFFFFFFFF
HIT_DTYPE = np.dtype([("t", "<f4"), ("instance", "<u4"), ("leaf", "<u4"), ("voxel", "<u4")])
RAY_DTYPE = np.dtype([("o", "<f4", 3), ("tmin", "<f4"), ("d", "<f4", 3), ("tmax", "<f4")])
assert HIT_DTYPE.itemsize == 16 and RAY_DTYPE.itemsize == 32


class Context:
    """One per GPU (dust_ctx_create).  Raises DustError(DUST_ERR_NO_DEVICE) without a CUDA device."""

    def __init__(self, device: int = 0):
        self.lib = capi.load()
        h = C.c_void_p()
        check(None, self.lib.dust_ctx_create(device, C.byref(h)), "dust_ctx_create")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.dust_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_stream(self, cuda_stream_ptr):
        """cuda_stream_ptr: a cudaStream_t handle as int (0 = legacy default stream); None = the ctx's own stream."""
        ext = cuda_stream_ptr is not None
        check(self.h, self.lib.dust_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr or 0), int(ext)), "dust_ctx_set_stream")

    def use_torch_stream(self):
        import torch
        self.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def synchronize(self):
        check(self.h, self.lib.dust_ctx_synchronize(self.h), "dust_ctx_synchronize")

    def stats(self) -> dict:
        s = DustStats()
        check(self.h, self.lib.dust_get_stats(self.h, C.byref(s)), "dust_get_stats")
        return s.as_dict()

    def launch_info(self) -> list:
        """Per-launch device times (and counter deltas) of the last trace call made with timing=True."""
        n = C.c_uint32()
        buf = (capi.DustLaunchInfo * 512)()
        check(self.h, self.lib.dust_get_launch_info(self.h, buf, 512, C.byref(n)), "dust_get_launch_info")
        return [{"kernel": buf[k].kernel.decode(), "ms": float(buf[k].ms),
                 **{f: int(getattr(buf[k], f)) for f in ("primary_rays", "bounce_rays", "shadow_rays", "mask_words", "child_ptrs", "hits_shaded", "mask_words_strict")}}
                for k in range(min(int(n.value), 512))]

    # ---- multi-GPU exchange step (dust_comm_* / dust_allgather_framebuffer) ---------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """Rank 0 calls this and hands the 128 bytes to the other ranks over the host's own channel."""
        buf = (C.c_uint8 * capi.DUST_COMM_ID_BYTES)()
        check(None, capi.load().dust_comm_unique_id(buf), "dust_comm_unique_id")
        return bytes(buf)

    def comm_init(self, uid: bytes, rank: int, nranks: int):
        """Collective: every rank of the shard group calls it with the same id."""
        assert len(uid) == capi.DUST_COMM_ID_BYTES
        buf = (C.c_uint8 * capi.DUST_COMM_ID_BYTES).from_buffer_copy(uid)
        check(self.h, self.lib.dust_comm_init(self.h, buf, rank, nranks), "dust_comm_init")

    def comm_destroy(self):
        check(self.h, self.lib.dust_comm_destroy(self.h), "dust_comm_destroy")

    def allgather_framebuffer(self, shard, gathered):
        """shard: this rank's compact tile buffer (CUDA tensor); gathered: nranks x the same, rank order."""
        nbytes = shard.numel() * shard.element_size()
        check(self.h, self.lib.dust_allgather_framebuffer(self.h, _ptr(shard), nbytes, _ptr(gathered)), "dust_allgather_framebuffer")

    def upload_tree(self, desc: capi.DustTreeDesc) -> "DeviceTree":
        return DeviceTree(self, desc)

    def create_scene(self, instances) -> "Scene":
        return Scene(self, instances)


class DeviceTree:
    def __init__(self, ctx: Context, desc: capi.DustTreeDesc = None, _handle=None):
        self.ctx = ctx
        if _handle is not None:
            self.h = _handle
            return
        h = C.c_void_p()
        check(ctx.h, ctx.lib.dust_tree_upload(ctx.h, C.byref(desc), C.byref(h)), "dust_tree_upload")
        self.h = h

    @classmethod
    def build_voxels(cls, ctx: Context, voxels, size, palette) -> "DeviceTree":
        """dust_tree_build_voxels: the tree of one .vox model built on the device from its voxel list
        ([n, 4] uint8 x, y, z, i as parsed; numpy = host memory, torch CUDA tensor = device memory)."""
        n = int(voxels.shape[0])
        if isinstance(voxels, np.ndarray):
            voxels = np.ascontiguousarray(voxels, dtype=np.uint8)
            mem = DUST_MEM_HOST
        else:
            mem = DUST_MEM_DEVICE
        pal = np.ascontiguousarray(np.asarray(palette, dtype=np.uint8).reshape(-1)[:255 * 4])
        h = C.c_void_p()
        check(ctx.h, ctx.lib.dust_tree_build_voxels(ctx.h, _ptr(voxels) if n else None, n, (C.c_uint32 * 3)(*[int(v) for v in size]),
                                                   pal.ctypes.data, mem, C.byref(h)), "dust_tree_build_voxels")
        return cls(ctx, _handle=h)

    def layout(self) -> capi.DustTreeLayout:
        out = capi.DustTreeLayout()
        check(self.ctx.h, self.ctx.lib.dust_tree_layout(self.h, C.byref(out)), "dust_tree_layout")
        return out

    def download(self) -> dict:
        """The device blob, section by section (numpy copies): nodes16 [n,64,4] u32, blocks [n,2] u32, mids / leaves [n,4] u32,
        palette_linear [256,4] f32, palette [256,4] u8, materials [n] u8."""
        L = self.layout()
        buf = np.empty(int(L.total_bytes), np.uint8)
        check(self.ctx.h, self.ctx.lib.dust_tree_download(self.ctx.h, self.h, buf.ctypes.data, buf.size), "dust_tree_download")
        def sec(off, count, dtype, shape):
            nbytes = int(count) * np.dtype(dtype).itemsize
            return buf[int(off):int(off) + nbytes].view(dtype).reshape(shape).copy()
        return {
            "nodes16": sec(L.off_nodes16, L.n_nodes16 * 256, np.uint32, (-1, 64, 4)), "blocks": sec(L.off_blocks, L.n_nodes16 * 2, np.uint32, (-1, 2)),
            "mids": sec(L.off_mids, L.n_mids * 4, np.uint32, (-1, 4)), "leaves": sec(L.off_leaves, L.n_leaves * 4, np.uint32, (-1, 4)),
            "palette_linear": sec(L.off_palette_linear, 1024, np.float32, (256, 4)), "palette": sec(L.off_palette, 1024, np.uint8, (256, 4)),
            "materials": sec(L.off_materials, L.n_materials, np.uint8, (-1,)),
            "bounds": (tuple(L.bounds_min), tuple(L.bounds_max)),
        }

    @property
    def leaf_count(self) -> int:
        n = C.c_uint64()
        self.ctx.lib.dust_tree_leaf_count(self.h, C.byref(n))
        return int(n.value)

    @property
    def device_bytes(self) -> int:
        n = C.c_uint64()
        self.ctx.lib.dust_tree_device_bytes(self.h, C.byref(n))
        return int(n.value)

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.dust_tree_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        self.close()


def _ptr(buf):
    """Device or host pointer of a torch tensor / numpy array / None."""
    if buf is None:
        return None
    if isinstance(buf, np.ndarray):
        return buf.ctypes.data
    return buf.data_ptr()


class Scene:
    """`instances` = iterable of (obj2world[12] row-major 3x4, DeviceTree)."""

    def __init__(self, ctx: Context, instances):
        self.ctx = ctx
        instances = list(instances)
        self._trees = [t for _, t in instances]          # keep alive
        arr = (DustInstance * max(1, len(instances)))()
        for i, (m, t) in enumerate(instances):
            arr[i].obj2world[:] = [float(v) for v in np.asarray(m, dtype=np.float32).reshape(12)]
            arr[i].tree = t.h
        h = C.c_void_p()
        check(ctx.h, ctx.lib.dust_scene_create(ctx.h, arr, len(instances), C.byref(h)), "dust_scene_create")
        self.h = h
        self.n_instances = len(instances)

    def update_transforms(self, obj2world):
        """New object->world matrices [n_instances, 12] for the same trees (moving instances); the old ones become
        last frame's for DustGBuffer.motion."""
        a = np.ascontiguousarray(np.asarray(obj2world, dtype=np.float32).reshape(-1, 12))
        check(self.ctx.h, self.ctx.lib.dust_scene_update_transforms(self.ctx.h, self.h, a.ctypes.data, len(a)),
              "dust_scene_update_transforms")

    def set_previous_transforms(self, obj2world_prev):
        """Last frame's object->world matrices, [n_instances, 12] (DustGBuffer.motion, hit.rchit:90-94)."""
        a = np.ascontiguousarray(np.asarray(obj2world_prev, dtype=np.float32).reshape(-1, 12))
        check(self.ctx.h, self.ctx.lib.dust_scene_set_previous_transforms(self.ctx.h, self.h, a.ctypes.data, len(a)),
              "dust_scene_set_previous_transforms")

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.dust_scene_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        self.close()

    # ---- dust_trace ------------------------------------------------------------------------------
    def trace(self, camera, sky, width, height, spp=1, max_bounces=0, seed=0, frame_index=0,
              shard_index=0, shard_count=1, hits=False, stats=False, fmt=DUST_FMT_RGBA32F, out=None, pipeline=None, gbuffer=False,
              demodulate=False):
        """Host-buffer call: returns dict(radiance[H,W,4], hit_instance, hit_leaf, hit_voxel, hit_t).
        pipeline: None / "auto", "megakernel" or "wavefront" (DUST_TRACE_PIPELINE_*; same results)."""
        ctx, lib = self.ctx, self.ctx.lib
        cfg = DustTraceConfig(width, height, spp, max_bounces, seed, frame_index, shard_index, shard_count,
                              (DUST_TRACE_WRITE_HITS if hits else 0) | (DUST_TRACE_STATS if stats else 0) | PIPELINE_FLAGS[pipeline] |
                              (capi.DUST_TRACE_DEMODULATE if demodulate else 0))
        res = {}
        if shard_count == 1:
            n_out, shape = width * height, (height, width, 4)
        else:
            per = C.c_uint64()
            lib.dust_shard_layout(width, height, shard_count, None, C.byref(per))
            n_out, shape = int(per.value), (int(per.value), 4)
        dt = np.float32 if fmt == DUST_FMT_RGBA32F else np.float16
        rad = out if out is not None else np.empty(shape, dtype=dt)
        assert rad.size == n_out * 4 and rad.dtype == dt
        fr = DustFrame(_ptr(rad), None, None, None, None, DUST_MEM_HOST, fmt)
        if hits:
            res["hit_instance"] = np.empty((height, width), np.uint32)
            res["hit_leaf"] = np.empty((height, width), np.uint32)
            res["hit_voxel"] = np.empty((height, width), np.uint8)
            res["hit_t"] = np.empty((height, width), np.float32)
            fr.hit_instance, fr.hit_leaf = _ptr(res["hit_instance"]), _ptr(res["hit_leaf"])
            fr.hit_voxel, fr.hit_t = _ptr(res["hit_voxel"]), _ptr(res["hit_t"])
        if gbuffer:
            # DustGBuffer images (standard.rs:880-917); normal / voxel_id keep their initial value on miss pixels
            g = {"albedo": np.zeros((height, width), np.uint32), "depth": np.zeros((height, width), np.float32),
                 "normal": np.zeros((height, width), np.uint32), "motion": np.zeros((height, width, 4), np.float16),
                 "voxel_id": np.full((height, width), 0xFFFFFFFF, np.uint32)}
            gb = DustGBuffer(*[_ptr(g[k]) for k in ("albedo", "depth", "normal", "motion", "voxel_id")])
            fr.gbuffer = C.pointer(gb)
            res["gbuffer"] = g
        check(ctx.h, lib.dust_trace(ctx.h, self.h, C.byref(camera), C.byref(sky), C.byref(cfg), C.byref(fr)), "dust_trace")
        res["radiance"] = rad
        if stats:
            res["stats"] = ctx.stats()
        return res

    def trace_async(self, camera, sky, width, height, out, spp=1, max_bounces=0, seed=0, frame_index=0, shard_index=0, shard_count=1,
                    fmt=DUST_FMT_RGBA16F, pipeline=None) -> int:
        """dust_trace_async: queue the frame and its download into `out` (a pinned host array the caller keeps alive and does not
        read until frame_wait(ticket)); returns the ticket.  Up to 3 frames in flight."""
        ctx, lib = self.ctx, self.ctx.lib
        cfg = DustTraceConfig(width, height, spp, max_bounces, seed, frame_index, shard_index, shard_count, PIPELINE_FLAGS[pipeline])
        fr = DustFrame(_ptr(out), None, None, None, None, DUST_MEM_HOST, fmt)
        t = C.c_uint64()
        check(ctx.h, lib.dust_trace_async(ctx.h, self.h, C.byref(camera), C.byref(sky), C.byref(cfg), C.byref(fr), C.byref(t)), "dust_trace_async")
        return int(t.value)

    def frame_wait(self, ticket: int):
        check(self.ctx.h, self.ctx.lib.dust_frame_wait(self.ctx.h, ticket), "dust_frame_wait")

    def trace_device(self, camera, sky, width, height, radiance, spp=1, max_bounces=0, seed=0, frame_index=0,
                     shard_index=0, shard_count=1, stats=False, fmt=DUST_FMT_RGBA32F, hit_buffers=None, pipeline=None, timing=False):
        """Device-buffer call (no copies, no sync): `radiance` is a CUDA tensor the kernel writes."""
        ctx, lib = self.ctx, self.ctx.lib
        flags = (DUST_TRACE_STATS if stats else 0) | (DUST_TRACE_WRITE_HITS if hit_buffers else 0) | PIPELINE_FLAGS[pipeline] | \
                (capi.DUST_TRACE_TIMING if timing else 0)
        cfg = DustTraceConfig(width, height, spp, max_bounces, seed, frame_index, shard_index, shard_count, flags)
        fr = DustFrame(_ptr(radiance), None, None, None, None, DUST_MEM_DEVICE, fmt)
        if hit_buffers:
            fr.hit_instance, fr.hit_leaf, fr.hit_voxel, fr.hit_t = (_ptr(b) for b in hit_buffers)
        check(ctx.h, lib.dust_trace(ctx.h, self.h, C.byref(camera), C.byref(sky), C.byref(cfg), C.byref(fr)), "dust_trace")

    # ---- dust_trace_rays -------------------------------------------------------------------------
    def trace_rays(self, rays: np.ndarray, stats=False, out=None, rough=False) -> np.ndarray:
        """Host-buffer call.  Copies and kernels of consecutive 2 Mi-ray chunks overlap when `rays` / `out` are pinned."""
        rays = np.ascontiguousarray(rays)
        assert rays.dtype == RAY_DTYPE
        hits = out if out is not None else np.empty(len(rays), dtype=HIT_DTYPE)
        assert hits.dtype == HIT_DTYPE and len(hits) == len(rays) and hits.flags.c_contiguous
        ctx = self.ctx
        check(ctx.h, ctx.lib.dust_trace_rays(ctx.h, self.h, _ptr(rays), len(rays), _ptr(hits), DUST_MEM_HOST,
                                             (DUST_TRACE_STATS if stats else 0) | (capi.DUST_RAYS_ROUGH if rough else 0)), "dust_trace_rays")
        return hits

    def trace_rays_device(self, rays, hits, n, stats=False):
        ctx = self.ctx
        check(ctx.h, ctx.lib.dust_trace_rays(ctx.h, self.h, _ptr(rays), n, _ptr(hits), DUST_MEM_DEVICE,
                                             DUST_TRACE_STATS if stats else 0), "dust_trace_rays")


CACHE_ENTRY_DTYPE = np.dtype([("fingerprint", "<u4"), ("radiance", "<u4"), ("last_accessed_frame", "<u2"), ("sample_count", "<u2")])
SURFEL_DTYPE = np.dtype([("position", "<f4", 3), ("direction", "<u4")])


class RadianceCache:
    """The persistent state of the reference's GI on the device (dust_cache_*): the spatial-hash radiance cache and the surfel pool
    (standard.rs:334-358).  gi_frame = dust_trace_gi: one frame of primary + ambient-occlusion + final-gather + surfel passes."""

    def __init__(self, ctx: Context, capacity=32 * 1024 * 1024, surfel_pool_size=720 * 480):
        self.ctx, self.capacity, self.pool_size = ctx, int(capacity), int(surfel_pool_size)
        h = C.c_void_p()
        check(ctx.h, ctx.lib.dust_cache_create(ctx.h, self.capacity, self.pool_size, C.byref(h)), "dust_cache_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.dust_cache_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        self.close()

    def clear(self):
        check(self.ctx.h, self.ctx.lib.dust_cache_clear(self.ctx.h, self.h), "dust_cache_clear")

    def download(self):
        """(table[capacity + 2], pool[pool_size]) as structured numpy arrays"""
        t, p = np.zeros(self.capacity + 2, CACHE_ENTRY_DTYPE), np.zeros(self.pool_size, SURFEL_DTYPE)
        check(self.ctx.h, self.ctx.lib.dust_cache_download(self.ctx.h, self.h, t.ctypes.data, len(t), p.ctypes.data, len(p)), "dust_cache_download")
        return t, p

    def upload(self, table=None, pool=None):
        t = None if table is None else np.ascontiguousarray(table, CACHE_ENTRY_DTYPE)
        p = None if pool is None else np.ascontiguousarray(pool, SURFEL_DTYPE)
        check(self.ctx.h, self.ctx.lib.dust_cache_upload(self.ctx.h, self.h, None if t is None else t.ctypes.data, 0 if t is None else len(t),
                                                         None if p is None else p.ctypes.data, 0 if p is None else len(p)), "dust_cache_upload")

    def gi_frame(self, scene: "Scene", camera, sky, width, height, seed=0, frame_index=0, gbuffer=True, stats=True):
        """Host-buffer call.  Returns dict(illuminance[H, W, 4] float16 = YCoCg + normHitDist, gbuffer, stats)."""
        ctx = self.ctx
        cfg = DustTraceConfig(width, height, 1, 0, seed, frame_index, 0, 1, 0)
        img = np.zeros((height, width, 4), np.float16)
        fr = DustFrame(_ptr(img), None, None, None, None, DUST_MEM_HOST, capi.DUST_FMT_YCOCG16F)
        res = {"illuminance": img}
        if gbuffer:
            g = {"albedo": np.zeros((height, width), np.uint32), "depth": np.zeros((height, width), np.float32),
                 "normal": np.zeros((height, width), np.uint32), "motion": np.zeros((height, width, 4), np.float16),
                 "voxel_id": np.full((height, width), 0xFFFFFFFF, np.uint32)}
            gb = DustGBuffer(*[_ptr(g[k]) for k in ("albedo", "depth", "normal", "motion", "voxel_id")])
            fr.gbuffer = C.pointer(gb)
            res["gbuffer"] = g
        st = capi.DustGiStats()
        check(ctx.h, ctx.lib.dust_trace_gi(ctx.h, scene.h, self.h, C.byref(camera), C.byref(sky), C.byref(cfg), C.byref(fr),
                                           C.byref(st) if stats else None), "dust_trace_gi")
        if stats:
            res["stats"] = st.as_dict()
        return res

    def gi_frame_device(self, scene: "Scene", camera, sky, width, height, illuminance, seed=0, frame_index=0):
        """Device-buffer call (no copies, no sync): `illuminance` is a CUDA float16 tensor [H, W, 4]."""
        ctx = self.ctx
        cfg = DustTraceConfig(width, height, 1, 0, seed, frame_index, 0, 1, 0)
        fr = DustFrame(_ptr(illuminance), None, None, None, None, DUST_MEM_DEVICE, capi.DUST_FMT_YCOCG16F)
        check(ctx.h, ctx.lib.dust_trace_gi(ctx.h, scene.h, self.h, C.byref(camera), C.byref(sky), C.byref(cfg), C.byref(fr), None), "dust_trace_gi")


class Exposure:
    """The reference's Histogram buffer + the two post passes (dust_auto_exposure, dust_tone_map).
    Images are numpy arrays (host) or torch CUDA tensors (device): illuminance [H,W,4] float16 in
    DUST_FMT_YCOCG16F, albedo [H,W] uint32 (DustGBuffer.albedo)."""
    IDENTITY = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)

    def __init__(self, ctx: Context, min_log_luminance=-6.0, max_log_luminance=8.5, time_coefficient=0.2):
        self.ctx = ctx
        self.params = capi.DustExposureParams(min_log_luminance, max_log_luminance - min_log_luminance, time_coefficient)
        h = C.c_void_p()
        check(ctx.h, ctx.lib.dust_exposure_create(ctx.h, C.byref(h)), "dust_exposure_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.dust_exposure_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        self.close()

    @property
    def avg(self) -> float:
        v = C.c_float()
        check(self.ctx.h, self.ctx.lib.dust_exposure_get(self.ctx.h, self.h, C.byref(v)), "dust_exposure_get")
        return float(v.value)

    @avg.setter
    def avg(self, value: float):
        check(self.ctx.h, self.ctx.lib.dust_exposure_set(self.ctx.h, self.h, float(value)), "dust_exposure_set")

    @staticmethod
    def _mem(buf):
        return DUST_MEM_HOST if isinstance(buf, np.ndarray) else DUST_MEM_DEVICE

    def update(self, illuminance, width, height, read_back=True):
        """auto_exposure.comp + auto_exposure_avg.comp; returns the adapted luminance when read_back."""
        v = C.c_float()
        check(self.ctx.h, self.ctx.lib.dust_auto_exposure(self.ctx.h, self.h, _ptr(illuminance), width, height, self._mem(illuminance),
                                                          C.byref(self.params), C.byref(v) if read_back else None), "dust_auto_exposure")
        return float(v.value) if read_back else None

    def tone_map(self, illuminance, albedo, width, height, dst, csc=IDENTITY, oetf=1):
        fmt = DUST_FMT_RGBA32F if (dst.dtype == np.float32 if isinstance(dst, np.ndarray) else dst.element_size() == 4) else DUST_FMT_RGBA16F
        tp = capi.DustToneMapParams((C.c_float * 9)(*csc), oetf, fmt)
        check(self.ctx.h, self.ctx.lib.dust_tone_map(self.ctx.h, self.h, _ptr(illuminance), _ptr(albedo), width, height, self._mem(illuminance),
                                                     C.byref(tp), _ptr(dst)), "dust_tone_map")
        return dst


def shard_layout(width, height, shard_count):
    tiles, px = C.c_uint32(), C.c_uint64()
    capi.load().dust_shard_layout(width, height, shard_count, C.byref(tiles), C.byref(px))
    return int(tiles.value), int(px.value)


def resolve_shards(ctx: Context, width, height, shard_count, gathered, out):
    check(ctx.h, ctx.lib.dust_resolve_shards(ctx.h, width, height, shard_count, _ptr(gathered), _ptr(out)), "dust_resolve_shards")
