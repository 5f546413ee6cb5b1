"""ctypes binding of the C ABI in include/dust_b200.h (libdust_b200.so).

This is the stub a Python host would write against the drop-in boundary; the Rust equivalent is
in INTEGRATION.md / rust/dust_cuda_sys.rs.  The library is loaded from dust_b200/lib/ (built
in-tree by `make -C dust_b200/csrc` or `__graft_entry__.build()`).  There is no fallback: a
missing library raises ImportError-like RuntimeError, a missing GPU makes `Context()` raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.environ.get("DUST_B200_LIB") or os.path.join(LIB_DIR, "libdust_b200.so")   # override: tuning builds only

DUST_OK = 0
ERR_NAMES = {
    0: "DUST_OK", 1: "DUST_ERR_INVALID_ARGUMENT", 2: "DUST_ERR_NO_DEVICE", 3: "DUST_ERR_CUDA",
    4: "DUST_ERR_UNSUPPORTED", 5: "DUST_ERR_OUT_OF_MEMORY", 6: "DUST_ERR_COMM",
}
DUST_TRACE_WRITE_HITS = 1
DUST_TRACE_STATS = 2
DUST_TRACE_PIPELINE_MEGAKERNEL = 4
DUST_TRACE_PIPELINE_WAVEFRONT = 8
PIPELINE_FLAGS = {None: 0, "auto": 0, "megakernel": 4, "wavefront": 8}
DUST_MEM_HOST, DUST_MEM_DEVICE = 0, 1
DUST_FMT_RGBA32F, DUST_FMT_RGBA16F, DUST_FMT_YCOCG16F = 0, 1, 2
DUST_TRACE_DEMODULATE = 16
DUST_TRACE_TIMING = 32
DUST_RAYS_ROUGH = 64


class DustError(RuntimeError):
    def __init__(self, code: int, where: str, msg: str = ""):
        self.code = code
        super().__init__(f"{where}: {ERR_NAMES.get(code, code)} {msg}".strip())


class DustTreeDesc(C.Structure):
    _fields_ = [
        ("n_levels", C.c_uint32),
        ("fanout_log2", C.c_uint32 * 4),
        ("root", C.c_void_p),
        ("root_bytes", C.c_uint64),
        ("pool_chunks", C.POINTER(C.c_void_p) * 3),
        ("n_chunks", C.c_uint32 * 3),
        ("item_bytes", C.c_uint32 * 3),
        ("n_items", C.c_uint32 * 3),
        ("materials", C.c_void_p),
        ("n_materials", C.c_uint64),
        ("palette", C.c_void_p),
    ]


class DustInstance(C.Structure):
    _fields_ = [("obj2world", C.c_float * 12), ("tree", C.c_void_p)]


class DustCamera(C.Structure):
    _fields_ = [
        ("view_proj", C.c_float * 16),
        ("inverse_view_proj", C.c_float * 16),
        ("camera_view_col0", C.c_float * 3), ("position_x", C.c_float),
        ("camera_view_col1", C.c_float * 3), ("position_y", C.c_float),
        ("camera_view_col2", C.c_float * 3), ("position_z", C.c_float),
        ("tan_half_fov", C.c_float), ("far", C.c_float), ("near", C.c_float), ("_padding", C.c_float),
    ]


class DustSkyChannel(C.Structure):
    _fields_ = [
        ("config", C.c_float * 9), ("radiance", C.c_float),
        ("ld_coefficient0", C.c_float), ("ld_coefficient1", C.c_float), ("ld_coefficient2", C.c_float * 4),
    ]


class DustSky(C.Structure):
    _fields_ = [("r", DustSkyChannel), ("g", DustSkyChannel), ("b", DustSkyChannel),
                ("direction", C.c_float * 4), ("sunlight_intensity", C.c_float * 4)]


class DustTraceConfig(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in
                ("width", "height", "spp", "max_bounces", "seed", "frame_index", "shard_index", "shard_count", "flags")]


class DustGBuffer(C.Structure):
    _fields_ = [("albedo", C.c_void_p), ("depth", C.c_void_p), ("normal", C.c_void_p), ("motion", C.c_void_p), ("voxel_id", C.c_void_p)]


class DustFrame(C.Structure):
    _fields_ = [
        ("radiance", C.c_void_p), ("hit_instance", C.c_void_p), ("hit_leaf", C.c_void_p),
        ("hit_voxel", C.c_void_p), ("hit_t", C.c_void_p),
        ("memory", C.c_uint32), ("radiance_format", C.c_uint32),
        ("gbuffer", C.POINTER(DustGBuffer)),
    ]


class DustTreeLayout(C.Structure):
    _fields_ = [("n_levels", C.c_uint32), ("_pad", C.c_uint32)] + \
               [(n, C.c_uint64) for n in ("n_nodes16", "n_mids", "n_leaves", "n_materials", "off_nodes16", "off_blocks", "off_mids", "off_leaves",
                                          "off_palette_linear", "off_palette", "off_materials", "total_bytes")] + \
               [("bounds_min", C.c_int32 * 3), ("bounds_max", C.c_int32 * 3)]


class DustLaunchInfo(C.Structure):
    _fields_ = [("kernel", C.c_char * 24), ("ms", C.c_float), ("_pad", C.c_uint32)] + \
               [(n, C.c_uint64) for n in ("primary_rays", "bounce_rays", "shadow_rays", "mask_words", "child_ptrs", "hits_shaded", "mask_words_strict")]


class DustCacheEntry(C.Structure):
    _fields_ = [("fingerprint", C.c_uint32), ("radiance", C.c_uint32), ("last_accessed_frame", C.c_uint16), ("sample_count", C.c_uint16)]


class DustSurfel(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("direction", C.c_uint32)]


class DustGiStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("near_hits", "far_hits", "far_misses", "cache_found", "surfels_scheduled", "surfel_rays",
                                          "cache_inserts", "insert_segments")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class DustExposureParams(C.Structure):
    _fields_ = [("min_log_luminance", C.c_float), ("log_luminance_range", C.c_float), ("time_coefficient", C.c_float)]


class DustToneMapParams(C.Structure):
    _fields_ = [("color_space_conversion", C.c_float * 9), ("output_transfer_function", C.c_uint32), ("dst_format", C.c_uint32)]


class DustStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("primary_rays", "bounce_rays", "shadow_rays", "mask_words", "child_ptrs", "hits_shaded", "kernel_launches", "mask_words_strict")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


assert C.sizeof(DustCamera) == 192 and C.sizeof(DustSky) == 224

# name -> (restype, argtypes); every symbol include/dust_b200.h declares
SYMBOLS = {
    "dust_abi_version": (C.c_int32, []),
    "dust_ctx_create": (C.c_int32, [C.c_int32, C.POINTER(C.c_void_p)]),
    "dust_ctx_destroy": (C.c_int32, [C.c_void_p]),
    "dust_ctx_set_stream": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32]),
    "dust_ctx_synchronize": (C.c_int32, [C.c_void_p]),
    "dust_last_error_string": (C.c_char_p, [C.c_void_p]),
    "dust_tree_upload": (C.c_int32, [C.c_void_p, C.POINTER(DustTreeDesc), C.POINTER(C.c_void_p)]),
    "dust_tree_build_voxels": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32), C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]),
    "dust_tree_layout": (C.c_int32, [C.c_void_p, C.POINTER(DustTreeLayout)]),
    "dust_tree_download": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "dust_tree_destroy": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "dust_tree_leaf_count": (C.c_int32, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "dust_tree_device_bytes": (C.c_int32, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "dust_scene_create": (C.c_int32, [C.c_void_p, C.POINTER(DustInstance), C.c_uint32, C.POINTER(C.c_void_p)]),
    "dust_scene_destroy": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "dust_scene_update_transforms": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "dust_scene_set_previous_transforms": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "dust_trace": (C.c_int32, [C.c_void_p, C.c_void_p, C.POINTER(DustCamera), C.POINTER(DustSky),
                               C.POINTER(DustTraceConfig), C.POINTER(DustFrame)]),
    "dust_trace_async": (C.c_int32, [C.c_void_p, C.c_void_p, C.POINTER(DustCamera), C.POINTER(DustSky),
                                     C.POINTER(DustTraceConfig), C.POINTER(DustFrame), C.POINTER(C.c_uint64)]),
    "dust_frame_wait": (C.c_int32, [C.c_void_p, C.c_uint64]),
    "dust_trace_rays": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32]),
    "dust_shard_layout": (C.c_int32, [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    "dust_resolve_shards": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "dust_comm_unique_id": (C.c_int32, [C.c_void_p]),
    "dust_comm_init": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "dust_comm_destroy": (C.c_int32, [C.c_void_p]),
    "dust_allgather_framebuffer": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "dust_exposure_create": (C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "dust_exposure_destroy": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "dust_exposure_get": (C.c_int32, [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]),
    "dust_exposure_set": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_float]),
    "dust_auto_exposure": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.POINTER(DustExposureParams), C.POINTER(C.c_float)]),
    "dust_tone_map": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                  C.POINTER(DustToneMapParams), C.c_void_p]),
    "dust_cache_create": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "dust_cache_destroy": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "dust_cache_clear": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "dust_cache_download": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]),
    "dust_cache_upload": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]),
    "dust_trace_gi": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(DustCamera), C.POINTER(DustSky), C.POINTER(DustTraceConfig),
                                  C.POINTER(DustFrame), C.POINTER(DustGiStats)]),
    "dust_get_stats": (C.c_int32, [C.c_void_p, C.POINTER(DustStats)]),
    "dust_get_launch_info": (C.c_int32, [C.c_void_p, C.POINTER(DustLaunchInfo), C.c_uint32, C.POINTER(C.c_uint32)]),
}
DUST_COMM_ID_BYTES = 128

_lib = None


def load() -> C.CDLL:
    """dlopen libdust_b200.so and bind every declared symbol.  Fails loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C dust_b200/csrc` "
                           "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    if lib.dust_abi_version() != 4:
        raise RuntimeError("libdust_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(ctx, rc: int, where: str):
    if rc != DUST_OK:
        msg = ""
        if ctx:
            msg = (load().dust_last_error_string(ctx) or b"").decode()
        raise DustError(rc, where, msg)
