"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes loader for oracle/_build/liboracle.so (the plain-C CPU restatement of the reference's hot
path).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; nothing under dust_b200/ does.  The library is (re)built on demand with
`make -C oracle` (gcc), because the checker has to exist wherever the tests run.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_SRCS = ["oracle_vdb.c", "oracle_vox.c", "oracle_sky.c", "oracle_trace.c", "oracle_dense.c",
         "oracle_vdb.h", "oracle_vox.h", "oracle_sky.h", "oracle_trace.h", "Makefile"]

O_MISS = 0xFFFFFFFF


def build(force=False):
    stale = force or not os.path.exists(LIB_PATH)
    if not stale:
        t = os.path.getmtime(LIB_PATH)
        stale = any(os.path.getmtime(os.path.join(_HERE, s)) > t for s in _SRCS)
    if stale:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return LIB_PATH


class OSkyChannel(C.Structure):
    _fields_ = [("config", C.c_float * 9), ("radiance", C.c_float), ("ld_coefficient0", C.c_float),
                ("ld_coefficient1", C.c_float), ("ld_coefficient2", C.c_float * 4)]


class OSky(C.Structure):
    _fields_ = [("r", OSkyChannel), ("g", OSkyChannel), ("b", OSkyChannel),
                ("direction", C.c_float * 4), ("sunlight_intensity", C.c_float * 4)]


class OCamera(C.Structure):
    _fields_ = [("view_proj", C.c_float * 16), ("inverse_view_proj", C.c_float * 16),
                ("camera_view_col0", C.c_float * 3), ("position_x", C.c_float),
                ("camera_view_col1", C.c_float * 3), ("position_y", C.c_float),
                ("camera_view_col2", C.c_float * 3), ("position_z", C.c_float),
                ("tan_half_fov", C.c_float), ("far", C.c_float), ("near", C.c_float), ("_padding", C.c_float)]


class OTraceConfig(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("width", "height", "spp", "max_bounces", "seed", "frame_index", "y0", "y1", "flags")]
O_TRACE_DEMODULATE = 1


class OExposureParams(C.Structure):
    _fields_ = [("min_log_lum", C.c_float), ("log_lum_range", C.c_float), ("time_coeff", C.c_float)]


class OSpatialHashEntry(C.Structure):
    _fields_ = [("fingerprint", C.c_uint32), ("radiance", C.c_uint32), ("last_accessed_frame", C.c_uint16), ("sample_count", C.c_uint16)]


class OSpatialHashKey(C.Structure):
    _fields_ = [("position", C.c_int32 * 3), ("direction", C.c_uint32)]


class OHistogram(C.Structure):
    _fields_ = [("histogram", C.c_uint32 * 256), ("avg", C.c_float)]


class OGiState(C.Structure):
    _fields_ = [("table", C.c_void_p), ("capacity", C.c_uint32), ("pool", C.c_void_p), ("pool_size", C.c_uint32)]


class OGiStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("near_hits", "far_hits", "far_misses", "cache_found", "surfels_scheduled", "surfel_rays",
                                          "cache_inserts", "insert_segments")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


CACHE_ENTRY_DTYPE = np.dtype([("fingerprint", "<u4"), ("radiance", "<u4"), ("last_accessed_frame", "<u2"), ("sample_count", "<u2")])
SURFEL_DTYPE = np.dtype([("position", "<f4", 3), ("direction", "<u4")])


class GiState:
    """Persistent state of the reference's GI (standard.rs:334-358): the spatial-hash table (capacity + 2 entries, zeroed) and the
    surfel pool (filled with 0xFF = invalid), as numpy arrays the oracle works on in place."""

    def __init__(self, capacity, pool_size):
        self.capacity, self.pool_size = int(capacity), int(pool_size)
        self.table = np.zeros(self.capacity + 2, CACHE_ENTRY_DTYPE)
        self.pool = np.zeros(self.pool_size, SURFEL_DTYPE)
        self.pool.view(np.uint8)[:] = 0xFF

    def c(self):
        return OGiState(self.table.ctypes.data, self.capacity, self.pool.ctypes.data, self.pool_size)


class OStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("primary_rays", "bounce_rays", "shadow_rays", "mask_words", "child_ptrs", "hits_shaded")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class OGBuffer(C.Structure):
    _fields_ = [("albedo", C.c_void_p), ("depth", C.c_void_p), ("normal", C.c_void_p), ("motion", C.c_void_p), ("voxel_id", C.c_void_p)]


class OVoxInstance(C.Structure):
    _fields_ = [("model_id", C.c_uint32), ("obj2world", C.c_float * 12)]


class OVoxModelData(C.Structure):
    _fields_ = [("tree", C.c_void_p), ("size", C.c_uint8 * 3), ("n_materials", C.c_uint64), ("materials", C.c_void_p),
                ("n_leaves", C.c_uint32), ("nodes", C.c_void_p), ("aabbs", C.c_void_p)]


HIT_DTYPE = np.dtype([("t", "<f4"), ("instance", "<u4"), ("leaf", "<u4"), ("voxel", "<u4")])
RAY_DTYPE = np.dtype([("o", "<f4", 3), ("tmin", "<f4"), ("d", "<f4", 3), ("tmax", "<f4")])
GPU_VOX_NODE = np.dtype([("x", "<u2"), ("y", "<u2"), ("z", "<u2"), ("w", "<u2"), ("mask", "<u8"),
                         ("material_ptr", "<u4"), ("avg_albedo", "<u4")])

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    vp, u32, u64, i32, f32, f64 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32, C.c_float, C.c_double
    pu32, pf32 = C.POINTER(u32), C.POINTER(f32)
    sig = {
        "o_bitmask_get": (i32, [vp, u32]), "o_bitmask_set": (None, [vp, u32, i32]),
        "o_bitmask_count_ones": (u32, [vp, u32]), "o_bitmask_iter_set_bits": (u32, [vp, u32, pu32, u32]),
        "o_pool_init": (None, [vp, u32, u32]), "o_pool_destroy": (None, [vp]), "o_pool_alloc": (u32, [vp]),
        "o_pool_free": (None, [vp, u32]), "o_pool_num_chunks": (u32, [vp]),
        "o_tree_new": (vp, [u32, pu32]), "o_tree_destroy": (None, [vp]),
        "o_tree_view": (vp, [u32, pu32, vp, vp, pu32, pu32]), "o_tree_view_free": (None, [vp]),
        "o_tree_set_value": (i32, [vp, u32, u32, u32, i32]), "o_tree_get_value": (i32, [vp, u32, u32, u32]),
        "o_tree_iter": (u64, [vp, pu32, u64]), "o_tree_iter_leaf": (u64, [vp, pu32, pu32, u64]),
        "o_tree_num_leaves": (u64, [vp]), "o_tree_root": (vp, [vp]), "o_tree_node_bytes": (u32, [vp, u32]),
        "o_tree_pool_count": (u32, [vp, u32]), "o_tree_pool_top": (u32, [vp, u32]),
        "o_tree_pool_num_chunks": (u32, [vp, u32]), "o_tree_pool_chunk": (vp, [vp, u32, u32]),
        "o_meta_mask": (u32, [vp]), "o_lowest_common_ancestor_level": (u32, [pu32, pu32, u32, u32]),
        "o_accessor_init": (None, [vp, vp]), "o_accessor_get": (i32, [vp, u32, u32, u32]),
        "o_vox_parse": (vp, [vp, u64]), "o_vox_scene_free": (None, [vp]),
        "o_vox_instances": (u32, [vp, C.POINTER(OVoxInstance), u32]),
        "o_vox_load_model": (C.POINTER(OVoxModelData), [vp, u32]), "o_vox_model_free": (None, [C.POINTER(OVoxModelData)]),
        "o_vox_to_transform": (None, [C.POINTER(i32), C.c_uint8, pu32, pf32]),
        "o_sunlight_bake": (None, [f32, pf32, pf32, C.POINTER(OSky)]),
        "o_sky_radiance": (None, [C.POINTER(OSky), pf32, pf32]), "o_sun_radiance": (None, [C.POINTER(OSky), pf32, pf32]),
        "o_srgb_to_linear": (f32, [f32]),
        "o_scene_create": (vp, [u32, pf32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]), "o_scene_destroy": (None, [vp]),
        "o_scene_set_cull": (None, [vp, i32]),
        "o_render": (None, [vp, C.POINTER(OCamera), C.POINTER(OSky), C.POINTER(OTraceConfig), vp, vp, vp, vp, vp, C.POINTER(OStats), u32]),
        "o_render_g": (None, [vp, C.POINTER(OCamera), C.POINTER(OSky), C.POINTER(OTraceConfig), vp, u32, vp, vp, vp, vp, vp,
                              C.POINTER(OGBuffer), C.POINTER(OStats), u32]),
        "o_scene_set_previous": (None, [vp, vp]),
        "o_render_rows": (None, [vp, C.POINTER(OCamera), C.POINTER(OSky), C.POINTER(OTraceConfig), vp, u32, vp, vp, vp, vp, vp, C.POINTER(OStats), u32]),
        "o_trace_rays": (None, [vp, vp, u64, vp, C.POINTER(OStats), u32]),
        "o_set_rough_blocks": (None, [i32]),
        "o_debug_pixel": (u32, [vp, C.POINTER(OCamera), C.POINTER(OSky), C.POINTER(OTraceConfig), u32, u32, vp, vp, u32, pf32]),
        "o_leaf_dda": (i32, [pf32, pf32, u64, f32, pf32, pu32]),
        "o_camera_ray": (None, [C.POINTER(OCamera), u32, u32, u32, u32, f32, f32, pf32, pf32]),
        "o_pcg": (u32, [u32]), "o_cossin_2pi": (None, [f32, pf32, pf32]),
        "o_rotate_by_normal": (None, [pf32, pf32, pf32]), "o_world2obj": (None, [pf32, pf32]),
        "o_half_to_float": (f32, [C.c_uint16]), "o_float_to_half": (C.c_uint16, [f32]), "o_post_log2": (f32, [f32]),
        "o_color_to_bin": (u32, [pf32, C.POINTER(OExposureParams)]),
        "o_exposure_histogram": (None, [vp, u32, u32, C.POINTER(OExposureParams), C.POINTER(OHistogram)]),
        "o_exposure_average": (None, [C.POINTER(OHistogram), u32, u32, C.POINTER(OExposureParams)]),
        "o_tone_map": (None, [vp, vp, u32, u32, f32, pf32, u32, vp]),
        "o_xxhash32": (u32, [u32]),
        "o_cache_fingerprint": (u32, [C.POINTER(OSpatialHashKey)]), "o_cache_location": (u32, [C.POINTER(OSpatialHashKey), u32]),
        "o_logluv_encode": (u32, [pf32]), "o_logluv_decode": (None, [u32, pf32]),
        "o_normal_to_face_id": (u32, [pf32]), "o_face_id_to_normal": (None, [u32, pf32]),
        "o_cache_key_of": (None, [pf32, pf32, C.POINTER(OSpatialHashKey)]),
        "o_cache_insert": (None, [vp, u32, C.POINTER(OSpatialHashKey), pf32, u32]),
        "o_cache_get": (i32, [vp, u32, C.POINTER(OSpatialHashKey), u32, pf32, pu32]),
        "o_gi_frame": (None, [vp, C.POINTER(OCamera), C.POINTER(OSky), C.POINTER(OTraceConfig), C.POINTER(OGiState), vp, C.POINTER(OGBuffer),
                              C.POINTER(OGiStats), u32]),
        "o_gi_unpack_normal": (None, [u32, pf32]), "o_leaf_ao": (i32, [pf32, pf32, u64, f32, pf32, pu32]),
        "o_dense_from_tree": (vp, [vp]), "o_dense_free": (None, [vp]),
        "o_dense_march": (i32, [vp, C.POINTER(f64), C.POINTER(f64), f64, f64, C.POINTER(i32), C.POINTER(f64)]),
        "o_dense_march_batch": (None, [vp, vp, vp, u64, f64, f64, i32, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def _u32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


class Tree:
    """Oracle Tree (oracle_vdb.c).  value: None / False / True."""

    def __init__(self, hierarchy, _borrowed=None, _owner=None):
        self.hierarchy = tuple(int(h) for h in hierarchy)
        self._owner = _owner
        if _borrowed is not None:
            self._h, self._own = _borrowed, False
        else:
            arr = (C.c_uint32 * len(self.hierarchy))(*self.hierarchy)
            self._h, self._own = lib().o_tree_new(len(self.hierarchy), arr), True
            if not self._h:
                raise ValueError("unsupported hierarchy")

    def __del__(self):
        if getattr(self, "_own", False) and self._h:
            lib().o_tree_destroy(self._h)
            self._h = None

    @property
    def levels(self):
        return len(self.hierarchy)

    def set_value(self, xyz, value):
        v = -1 if value is None else int(bool(value))
        if lib().o_tree_set_value(self._h, int(xyz[0]), int(xyz[1]), int(xyz[2]), v) != 0:
            raise NotImplementedError("clearing through an InternalNode is todo!() in the reference")

    def get_value(self, xyz):
        r = lib().o_tree_get_value(self._h, int(xyz[0]), int(xyz[1]), int(xyz[2]))
        return None if r < 0 else bool(r)

    def iter(self):
        n = lib().o_tree_iter(self._h, None, 0)
        out = np.zeros((n, 3), np.uint32)
        if n:
            lib().o_tree_iter(self._h, _u32p(out), n)
        return out

    def iter_leaf(self):
        n = lib().o_tree_num_leaves(self._h)
        pos, ptr = np.zeros((n, 3), np.uint32), np.zeros(n, np.uint32)
        if n:
            lib().o_tree_iter_leaf(self._h, _u32p(pos), _u32p(ptr), n)
        return pos, ptr

    def node_bytes(self, level):
        return lib().o_tree_node_bytes(self._h, level)

    def pool_count(self, level):
        return lib().o_tree_pool_count(self._h, level)

    def pool_top(self, level):
        return lib().o_tree_pool_top(self._h, level)

    def root_bytes(self):
        return C.string_at(lib().o_tree_root(self._h), self.node_bytes(self.levels - 1))

    def pool_bytes(self, level):
        top, nb = self.pool_top(level), self.node_bytes(level)
        out = bytearray()
        for c in range(lib().o_tree_pool_num_chunks(self._h, level)):
            n = min(1024, top - c * 1024)
            if n <= 0:
                break
            out += C.string_at(lib().o_tree_pool_chunk(self._h, level, c), n * nb)
        return bytes(out)

    def meta_mask(self):
        return lib().o_meta_mask(self._h)

    def accessor_get_many(self, xyz):
        acc = C.create_string_buffer(64)
        lib().o_accessor_init(acc, self._h)
        out = []
        for x, y, z in np.asarray(xyz).reshape(-1, 3):
            r = lib().o_accessor_get(acc, int(x), int(y), int(z))
            out.append(None if r < 0 else bool(r))
        return out


class TreeView(Tree):
    """Oracle view over a tree described by a DustTreeDesc-shaped object (same bytes the product got)."""

    def __init__(self, desc):
        n = int(desc.n_levels)
        self.hierarchy = tuple(int(desc.fanout_log2[i]) for i in range(n))
        self._desc = desc                                   # keeps the described memory alive
        hier = (C.c_uint32 * n)(*self.hierarchy)
        chunks = (C.c_void_p * 3)(*[C.cast(desc.pool_chunks[i], C.c_void_p) for i in range(3)])
        nch = (C.c_uint32 * 3)(*[int(desc.n_chunks[i]) for i in range(3)])
        tops = (C.c_uint32 * 3)(*[int(desc.n_items[i]) for i in range(3)])
        self._h = lib().o_tree_view(n, hier, desc.root, chunks, nch, tops)
        self._own, self._owner = False, None
        if not self._h:
            raise ValueError("unsupported hierarchy")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().o_tree_view_free(self._h)
            self._h = None


def lowest_common_ancestor_level(a, b, mask, root_level):
    return lib().o_lowest_common_ancestor_level((C.c_uint32 * 3)(*a), (C.c_uint32 * 3)(*b), mask, root_level)


class VoxModel:
    def __init__(self, ptr):
        self._p = ptr
        d = ptr.contents
        self.tree = Tree((4, 2, 2), _borrowed=d.tree, _owner=self)
        self.size = tuple(d.size)
        self.materials = (np.frombuffer(C.string_at(d.materials, d.n_materials), np.uint8) if d.n_materials else np.zeros(0, np.uint8))
        self.nodes = (np.frombuffer(C.string_at(d.nodes, d.n_leaves * 24), GPU_VOX_NODE) if d.n_leaves else np.zeros(0, GPU_VOX_NODE))
        self.aabbs = (np.frombuffer(C.string_at(d.aabbs, d.n_leaves * 24), np.float32).reshape(-1, 6) if d.n_leaves else np.zeros((0, 6), np.float32))

    def __del__(self):
        if getattr(self, "_p", None):
            self.tree = None
            lib().o_vox_model_free(self._p)
            self._p = None


class VoxFile:
    def __init__(self, data: bytes):
        self._buf = bytes(data)
        self._h = lib().o_vox_parse(self._buf, len(self._buf))
        if not self._h:
            raise ValueError("oracle: cannot parse .vox")
        n = lib().o_vox_instances(self._h, None, 0)
        arr = (OVoxInstance * max(1, n))()
        lib().o_vox_instances(self._h, arr, n)
        self.instances = [(int(arr[i].model_id), np.array(arr[i].obj2world, np.float32)) for i in range(n)]
        # OVoxScene: n_models u32, pad, models*, palette[256][4] @16
        self.palette = np.frombuffer(C.string_at(self._h + 16, 1024), np.uint8).reshape(256, 4).copy()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().o_vox_scene_free(self._h)
            self._h = None

    def load_model(self, model_id) -> VoxModel:
        p = lib().o_vox_load_model(self._h, model_id)
        if not p:
            raise ValueError("oracle: load_model failed")
        return VoxModel(p)


def to_transform(translation, rotation, size):
    out = (C.c_float * 12)()
    lib().o_vox_to_transform((C.c_int32 * 3)(*translation), rotation, (C.c_uint32 * 3)(*size), out)
    return np.array(out, np.float32)


def sunlight_bake(turbidity=1.0, albedo=(0.2, 0.2, 0.2), direction=(0.0, 0.80114365, -0.5984721)) -> OSky:
    out = OSky()
    lib().o_sunlight_bake(float(turbidity), (C.c_float * 3)(*albedo), (C.c_float * 3)(*direction), C.byref(out))
    return out


def sky_radiance(sky: OSky, d):
    out = (C.c_float * 3)()
    lib().o_sky_radiance(C.byref(sky), (C.c_float * 3)(*d), out)
    return np.array(out, np.float32)


def sun_radiance(sky: OSky, d):
    out = (C.c_float * 3)()
    lib().o_sun_radiance(C.byref(sky), (C.c_float * 3)(*d), out)
    return np.array(out, np.float32)


def as_struct(cls, other):
    """Reinterpret a same-layout ctypes struct (e.g. dust_b200.capi.DustSky) as the oracle's type."""
    assert C.sizeof(cls) == C.sizeof(other)
    return cls.from_buffer_copy(bytes(other))


class Scene:
    """instances: list of (obj2world[12], tree_handle(void*), materials ndarray, palette ndarray[>=255,4]).
    `tree_handle` is an OTree* — either an oracle Tree._h, or ANY memory with the same layout."""

    def __init__(self, instances):
        n = len(instances)
        self._keep = instances
        o2w = np.ascontiguousarray(np.stack([np.asarray(i[0], np.float32).reshape(12) for i in instances]) if n else np.zeros((0, 12), np.float32))
        trees = (C.c_void_p * max(1, n))(*[i[1] for i in instances])
        self._mats = [np.ascontiguousarray(i[2], np.uint8) for i in instances]
        self._pals = [np.ascontiguousarray(i[3], np.uint8) for i in instances]
        mats = (C.c_void_p * max(1, n))(*[m.ctypes.data for m in self._mats])
        pals = (C.c_void_p * max(1, n))(*[p.ctypes.data for p in self._pals])
        self._h = lib().o_scene_create(n, o2w.ctypes.data_as(C.POINTER(C.c_float)), trees, mats, pals)
        if not self._h:
            raise ValueError("oracle: scene_create failed")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().o_scene_destroy(self._h)
            self._h = None

    def set_cull(self, enabled: bool):
        lib().o_scene_set_cull(self._h, int(bool(enabled)))

    def set_previous_transforms(self, obj2world_prev):
        a = np.ascontiguousarray(np.asarray(obj2world_prev, dtype=np.float32).reshape(-1, 12))
        lib().o_scene_set_previous(self._h, a.ctypes.data)

    def render(self, camera, sky, width, height, spp=1, max_bounces=0, seed=0, frame_index=0, y0=0, y1=None,
               hits=True, threads=None, rows=None, gbuffer=False, demodulate=False):
        """rows: optional list of rows rendered in ONE call (the other rows of the returned frame stay zero).
        gbuffer: also return the G-buffer images of the sample-0 primary hit (motion as float16, like the image)."""
        cam = camera if isinstance(camera, OCamera) else as_struct(OCamera, camera)
        sk = sky if isinstance(sky, OSky) else as_struct(OSky, sky)
        y1 = height if y1 is None else y1
        cfg = OTraceConfig(width, height, spp, max_bounces, seed, frame_index, y0, y1, O_TRACE_DEMODULATE if demodulate else 0)
        rad = np.zeros((height, width, 4), np.float32)
        res = {"radiance": rad}
        hp = [None] * 4
        if hits:
            res["hit_instance"] = np.full((height, width), O_MISS, np.uint32)
            res["hit_leaf"] = np.full((height, width), O_MISS, np.uint32)
            res["hit_voxel"] = np.full((height, width), 0xFF, np.uint8)
            res["hit_t"] = np.full((height, width), np.inf, np.float32)
            hp = [res[k].ctypes.data for k in ("hit_instance", "hit_leaf", "hit_voxel", "hit_t")]
        st = OStats()
        if gbuffer:
            g = {"albedo": np.zeros((height, width), np.uint32), "depth": np.zeros((height, width), np.float32),
                 "normal": np.zeros((height, width), np.uint32), "voxel_id": np.full((height, width), 0xFFFFFFFF, np.uint32)}
            motion = np.zeros((height, width, 4), np.float32)
            gb = OGBuffer(g["albedo"].ctypes.data, g["depth"].ctypes.data, g["normal"].ctypes.data, motion.ctypes.data, g["voxel_id"].ctypes.data)
            ra = None if rows is None else np.ascontiguousarray(rows, dtype=np.uint32)
            lib().o_render_g(self._h, C.byref(cam), C.byref(sk), C.byref(cfg), None if ra is None else ra.ctypes.data, 0 if ra is None else len(ra),
                             rad.ctypes.data, hp[0], hp[1], hp[2], hp[3], C.byref(gb), C.byref(st), threads or os.cpu_count() or 1)
            g["motion"] = motion.astype(np.float16)           # R16G16B16A16_SFLOAT (round to nearest even)
            res["gbuffer"] = g
        elif rows is not None:
            ra = np.ascontiguousarray(rows, dtype=np.uint32)
            lib().o_render_rows(self._h, C.byref(cam), C.byref(sk), C.byref(cfg), ra.ctypes.data, len(ra), rad.ctypes.data,
                                hp[0], hp[1], hp[2], hp[3], C.byref(st), threads or os.cpu_count() or 1)
        else:
            lib().o_render(self._h, C.byref(cam), C.byref(sk), C.byref(cfg), rad.ctypes.data, hp[0], hp[1], hp[2], hp[3],
                           C.byref(st), threads or os.cpu_count() or 1)
        res["stats"] = st.as_dict()
        return res

    def gi_frame(self, camera, sky, width, height, state: "GiState", seed=0, frame_index=0, threads=None):
        """One frame of the reference's own GI (oracle_gi.h).  Returns illuminance [H, W, 4] float16 (YCoCg + normHitDist), the
        G-buffer of the primary hit and the pass statistics; `state` (cache table + surfel pool) is updated in place."""
        cam = camera if isinstance(camera, OCamera) else as_struct(OCamera, camera)
        sk = sky if isinstance(sky, OSky) else as_struct(OSky, sky)
        cfg = OTraceConfig(width, height, 1, 0, seed, frame_index, 0, height, 0)
        img = np.zeros((height, width, 4), np.uint16)
        g = {"albedo": np.zeros((height, width), np.uint32), "depth": np.zeros((height, width), np.float32),
             "normal": np.zeros((height, width), np.uint32), "voxel_id": np.full((height, width), 0xFFFFFFFF, np.uint32)}
        motion = np.zeros((height, width, 4), np.float32)
        gb = OGBuffer(g["albedo"].ctypes.data, g["depth"].ctypes.data, g["normal"].ctypes.data, motion.ctypes.data, g["voxel_id"].ctypes.data)
        st = OGiStats()
        cs = state.c()
        lib().o_gi_frame(self._h, C.byref(cam), C.byref(sk), C.byref(cfg), C.byref(cs), img.ctypes.data, C.byref(gb), C.byref(st),
                         threads or os.cpu_count() or 1)
        g["motion"] = motion.astype(np.float16)
        return {"illuminance": img.view(np.float16), "gbuffer": g, "stats": st.as_dict()}

    def debug_pixel(self, camera, sky, width, height, px, py, spp=1, max_bounces=0, seed=0, frame_index=0, cap=256):
        """All rays of one pixel's paths and the oracle's answers (voxel bit 31 = any-hit shadow ray)."""
        cam = camera if isinstance(camera, OCamera) else as_struct(OCamera, camera)
        sk = sky if isinstance(sky, OSky) else as_struct(OSky, sky)
        cfg = OTraceConfig(width, height, spp, max_bounces, seed, frame_index, 0, height)
        rays, hits = np.zeros(cap, RAY_DTYPE), np.zeros(cap, HIT_DTYPE)
        rgba = (C.c_float * 4)()
        n = lib().o_debug_pixel(self._h, C.byref(cam), C.byref(sk), C.byref(cfg), px, py, rays.ctypes.data, hits.ctypes.data, cap, rgba)
        return rays[:n], hits[:n], np.array(rgba, np.float32)

    def trace_rays(self, rays: np.ndarray, threads=None, rough=False):
        """rough: block-granular leaves (final_gather/rough.rint): a leaf is hit where the ray enters its 4^3 box."""
        rays = np.ascontiguousarray(rays)
        assert rays.dtype.itemsize == 32
        hits = np.zeros(len(rays), HIT_DTYPE)
        st = OStats()
        lib().o_set_rough_blocks(int(bool(rough)))
        try:
            lib().o_trace_rays(self._h, rays.ctypes.data, len(rays), hits.ctypes.data, C.byref(st), threads or os.cpu_count() or 1)
        finally:
            lib().o_set_rough_blocks(0)
        return hits, st.as_dict()


# ---- radiance cache (oracle_cache.c): groundwork for SURVEY 8f rank 2, no CUDA counterpart yet ---------

class RadianceCache:
    """Sequential semantics of headers/spatial_hash.glsl's insert / get on a table of `capacity` SpatialHashEntry."""

    def __init__(self, capacity: int):
        self.capacity = int(capacity)
        self.table = (OSpatialHashEntry * (self.capacity + 2))()

    @staticmethod
    def key(position, direction) -> OSpatialHashKey:
        return OSpatialHashKey((C.c_int32 * 3)(*[int(v) for v in position]), int(direction))

    def insert(self, key, value, frame_index):
        lib().o_cache_insert(self.table, self.capacity, C.byref(key), (C.c_float * 3)(*[float(v) for v in value]), int(frame_index))

    def get(self, key, frame_index):
        out, n = (C.c_float * 3)(), C.c_uint32()
        found = lib().o_cache_get(self.table, self.capacity, C.byref(key), int(frame_index), out, C.byref(n))
        return bool(found), np.array(out[:], np.float32), int(n.value)

    def entries(self) -> np.ndarray:
        return np.frombuffer(self.table, dtype=np.dtype([("fingerprint", "<u4"), ("radiance", "<u4"), ("frame", "<u2"), ("count", "<u2")])).copy()


# ---- post chain (oracle_post.c) ----------------------------------------------------------------------

def pack_illuminance(radiance):
    """REBLUR_FrontEnd_PackRadianceAndNormHitDist of an ACEScg float32 frame -> [H,W,4] float16 (Y, Co, Cg, 1):
    the DUST_FMT_YCOCG16F image (headers/nrd.glsl:99-147; same float expression order as the kernel)."""
    r, g, b = (np.asarray(radiance[..., k], np.float32) for k in range(3))
    f = np.float32
    out = np.empty(radiance.shape[:-1] + (4,), np.float32)
    out[..., 0] = (r * f(0.25) + g * f(0.5)) + b * f(0.25)
    out[..., 1] = (r * f(0.5) + g * f(0.0)) + b * f(-0.5)
    out[..., 2] = (r * f(-0.25) + g * f(0.5)) + b * f(-0.25)
    out[..., 3] = 1.0
    with np.errstate(over="ignore"):
        return out.astype(np.float16)


def exposure_update(illuminance16, avg, min_log_lum=-6.0, log_lum_range=14.5, time_coeff=0.2):
    """auto_exposure.comp + auto_exposure_avg.comp on a [H,W,4] float16 image; returns (histogram[256], new avg)."""
    img = np.ascontiguousarray(illuminance16, dtype=np.float16)
    h, w = img.shape[:2]
    p = OExposureParams(min_log_lum, log_lum_range, time_coeff)
    hist = OHistogram()
    hist.avg = avg
    lib().o_exposure_histogram(img.ctypes.data, w, h, C.byref(p), C.byref(hist))
    bins = np.array(hist.histogram[:], np.uint32)
    lib().o_exposure_average(C.byref(hist), w, h, C.byref(p))
    return bins, float(hist.avg)


def tone_map(illuminance16, albedo, avg, csc=(1, 0, 0, 0, 1, 0, 0, 0, 1), oetf=1):
    img = np.ascontiguousarray(illuminance16, dtype=np.float16)
    alb = np.ascontiguousarray(albedo, dtype=np.uint32)
    h, w = img.shape[:2]
    dst = np.empty((h, w, 4), np.float32)
    m = (C.c_float * 9)(*[float(v) for v in csc])
    lib().o_tone_map(img.ctypes.data, alb.ctypes.data, w, h, avg, m, oetf, dst.ctypes.data)
    return dst


def leaf_dda(origin, direction, grid, tmin=0.0):
    t, v = C.c_float(), C.c_uint32()
    ok = lib().o_leaf_dda((C.c_float * 3)(*origin), (C.c_float * 3)(*direction), int(grid), float(tmin), C.byref(t), C.byref(v))
    return (float(t.value), int(v.value)) if ok else None


class Dense:
    def __init__(self, tree: Tree):
        self._h = lib().o_dense_from_tree(tree._h)
        if not self._h:
            raise MemoryError("dense grid of this tree's occupied bounds would exceed 2 GiB")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().o_dense_free(self._h)
            self._h = None

    def march(self, o, d, tmin, tmax):
        xyz, t = (C.c_int32 * 3)(), C.c_double()
        ok = lib().o_dense_march(self._h, (C.c_double * 3)(*[float(v) for v in o]), (C.c_double * 3)(*[float(v) for v in d]),
                                 float(tmin), float(tmax), xyz, C.byref(t))
        return (tuple(xyz), float(t.value)) if ok else None

    def march_batch(self, o, d, tmin, tmax, guard=False):
        """o, d: [n, 3] float64 object-space rays.  Returns (hit[n] bool, xyz[n, 3] int32, t[n] float64).
        guard: also apply hit.rint:107's leaf-exit rule (see oracle_dense.c)."""
        o = np.ascontiguousarray(o, np.float64); d = np.ascontiguousarray(d, np.float64)
        n = len(o)
        xyz, t, hit = np.zeros((n, 3), np.int32), np.zeros(n, np.float64), np.zeros(n, np.uint8)
        lib().o_dense_march_batch(self._h, o.ctypes.data, d.ctypes.data, n, float(tmin), float(tmax), int(bool(guard)),
                                  xyz.ctypes.data, t.ctypes.data, hit.ctypes.data)
        return hit.astype(bool), xyz, t
