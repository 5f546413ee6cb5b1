"""ctypes wrapper of libdust_host.so — the C++ host-side mirror of the reference's Rust crates
(`crates/vdb` Tree/Pool/Accessor, `crates/vox` loader/collector/geometry, `crates/render`
Sunlight::bake / CameraSettings / StandardPipeline::render).  Python only marshals arguments;
all host logic lives in dust_b200/csrc/host/*.hpp.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi

LIB_PATH = os.path.join(capi.LIB_DIR, "libdust_host.so")
_lib = None


def _u32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    capi.load()                                   # libdust_host links against libdust_b200
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C dust_b200/csrc`")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32, f32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32, C.c_float
    sig = {
        "dh_last_error": (C.c_char_p, []),
        "dh_tree_new": (vp, [u32, C.POINTER(u32)]),
        "dh_tree_free": (None, [vp]),
        "dh_tree_set": (i32, [vp, u32, u32, u32, i32]),
        "dh_tree_get": (i32, [vp, u32, u32, u32]),
        "dh_tree_set_many": (None, [vp, C.POINTER(u32), u64]),
        "dh_tree_or_leaves": (None, [vp, C.POINTER(u32), C.POINTER(u64), u64]),
        "dh_tree_iter": (u64, [vp, C.POINTER(u32), u64]),
        "dh_tree_iter_leaf": (u64, [vp, C.POINTER(u32), C.POINTER(u32), u64]),
        "dh_tree_node_bytes": (u32, [vp, u32]),
        "dh_tree_pool_count": (u32, [vp, u32]),
        "dh_tree_pool_top": (u32, [vp, u32]),
        "dh_tree_pool_chunks": (u32, [vp, u32]),
        "dh_tree_pool_chunk": (vp, [vp, u32, u32]),
        "dh_tree_root": (vp, [vp]),
        "dh_tree_meta_mask": (u32, [vp]),
        "dh_lca_level": (u32, [C.POINTER(u32), C.POINTER(u32), u32, u32]),
        "dh_tree_accessor_get_many": (None, [vp, C.POINTER(u32), u64, C.POINTER(i32)]),
        "dh_tree_describe": (None, [vp, vp, u64, vp, C.POINTER(capi.DustTreeDesc)]),
        "dh_pool_new": (vp, [u32, u32]),
        "dh_pool_free_pool": (None, [vp]),
        "dh_pool_alloc": (u32, [vp]),
        "dh_pool_free": (None, [vp, u32]),
        "dh_pool_num_chunks": (u32, [vp]),
        "dh_vox_parse": (vp, [vp, u64]),
        "dh_vox_free": (None, [vp]),
        "dh_vox_num_models": (u32, [vp]),
        "dh_vox_num_instances": (u32, [vp]),
        "dh_vox_instance": (None, [vp, u32, C.POINTER(u32), C.POINTER(f32)]),
        "dh_vox_palette": (vp, [vp]),
        "dh_vox_model_info": (None, [vp, u32, C.POINTER(u32), C.POINTER(u64)]),
        "dh_vox_model_voxels": (vp, [vp, u32]),
        "dh_to_transform": (None, [C.POINTER(i32), C.c_uint8, C.POINTER(u32), C.POINTER(f32)]),
        "dh_vox_load_model": (vp, [vp, u32]),
        "dh_model_free": (None, [vp]),
        "dh_model_tree": (vp, [vp]),
        "dh_model_num_materials": (u64, [vp]),
        "dh_model_materials": (vp, [vp]),
        "dh_model_num_leaves": (u64, [vp]),
        "dh_model_nodes": (vp, [vp]),
        "dh_model_aabbs": (vp, [vp]),
        "dh_model_size": (None, [vp, C.POINTER(C.c_uint8)]),
        "dh_model_describe": (None, [vp, C.POINTER(capi.DustTreeDesc)]),
        "dh_synth_new": (vp, [u32, C.c_double, u32]),
        "dh_synth_free": (None, [vp]),
        "dh_synth_tree": (vp, [vp]),
        "dh_synth_num_leaves": (u64, [vp]),
        "dh_synth_materials": (vp, [vp]),
        "dh_synth_num_materials": (u64, [vp]),
        "dh_synth_palette": (vp, [vp]),
        "dh_synth_describe": (None, [vp, C.POINTER(capi.DustTreeDesc)]),
        "dh_sunlight_bake": (None, [f32, C.POINTER(f32), C.POINTER(f32), C.POINTER(capi.DustSky)]),
        "dh_camera_look_at": (None, [C.POINTER(f32), C.POINTER(f32), C.POINTER(f32), f32, f32, f32, u32, u32,
                                     C.POINTER(capi.DustCamera)]),
        "dh_castle_sun_direction": (None, [f32, C.POINTER(f32)]),
        "dh_pipeline_render": (i32, [vp, vp, C.POINTER(capi.DustFrame), u32, u32, f32, C.POINTER(f32), C.POINTER(f32),
                                     C.POINTER(f32), C.POINTER(f32), C.POINTER(f32), f32, f32, f32,
                                     u32, u32, u32, u32, u32, u32, u32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def _err():
    return (lib().dh_last_error() or b"").decode()


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class Tree:
    """`dust_vdb::Tree<hierarchy!(..)>` (crates/vdb/src/tree.rs).  value: None / False / True."""

    def __init__(self, hierarchy, _borrowed=None, _owner=None):
        self.hierarchy = tuple(int(h) for h in hierarchy)
        self._owner = _owner
        if _borrowed is not None:
            self._h, self._own = _borrowed, False
        else:
            arr = (C.c_uint32 * len(self.hierarchy))(*self.hierarchy)
            self._h = lib().dh_tree_new(len(self.hierarchy), arr)
            self._own = True
            if not self._h:
                raise ValueError(_err())

    def __del__(self):
        if getattr(self, "_own", False) and self._h:
            lib().dh_tree_free(self._h)
            self._h = None

    @property
    def levels(self):
        return len(self.hierarchy)

    def set_value(self, xyz, value):
        v = -1 if value is None else int(bool(value))
        if lib().dh_tree_set(self._h, int(xyz[0]), int(xyz[1]), int(xyz[2]), v) != 0:
            raise NotImplementedError(_err())

    def get_value(self, xyz):
        r = lib().dh_tree_get(self._h, int(xyz[0]), int(xyz[1]), int(xyz[2]))
        return None if r < 0 else bool(r)

    def set_many(self, xyz: np.ndarray):
        xyz = np.ascontiguousarray(xyz, dtype=np.uint32).reshape(-1, 3)
        lib().dh_tree_set_many(self._h, _u32p(xyz), len(xyz))

    def or_leaves(self, leaf_xyz: np.ndarray, masks: np.ndarray):
        leaf_xyz = np.ascontiguousarray(leaf_xyz, dtype=np.uint32).reshape(-1, 3)
        masks = np.ascontiguousarray(masks, dtype=np.uint64)
        lib().dh_tree_or_leaves(self._h, _u32p(leaf_xyz), masks.ctypes.data_as(C.POINTER(C.c_uint64)), len(masks))

    def iter(self) -> np.ndarray:
        n = lib().dh_tree_iter(self._h, None, 0)
        out = np.zeros((n, 3), dtype=np.uint32)
        if n:
            lib().dh_tree_iter(self._h, _u32p(out), n)
        return out

    def iter_leaf(self):
        n = lib().dh_tree_iter_leaf(self._h, None, None, 0)
        pos = np.zeros((n, 3), dtype=np.uint32)
        ptr = np.zeros(n, dtype=np.uint32)
        if n:
            lib().dh_tree_iter_leaf(self._h, _u32p(pos), _u32p(ptr), n)
        return pos, ptr

    def node_bytes(self, level):
        return lib().dh_tree_node_bytes(self._h, level)

    def pool_count(self, level):
        return lib().dh_tree_pool_count(self._h, level)

    def pool_top(self, level):
        return lib().dh_tree_pool_top(self._h, level)

    def root_bytes(self) -> bytes:
        return C.string_at(lib().dh_tree_root(self._h), self.node_bytes(self.levels - 1))

    def pool_bytes(self, level) -> bytes:
        """Concatenated live range of pool[level] (slots 0..top)."""
        top, nb = self.pool_top(level), self.node_bytes(level)
        out = bytearray()
        for c in range(lib().dh_tree_pool_chunks(self._h, level)):
            n = min(1024, top - c * 1024)
            if n <= 0:
                break
            out += C.string_at(lib().dh_tree_pool_chunk(self._h, level, c), n * nb)
        return bytes(out)

    def meta_mask(self):
        return lib().dh_tree_meta_mask(self._h)

    @staticmethod
    def lowest_common_ancestor_level(a, b, mask, root_level):
        aa, bb = (C.c_uint32 * 3)(*a), (C.c_uint32 * 3)(*b)
        return lib().dh_lca_level(aa, bb, mask, root_level)

    def accessor_get_many(self, xyz: np.ndarray):
        xyz = np.ascontiguousarray(xyz, dtype=np.uint32).reshape(-1, 3)
        out = np.zeros(len(xyz), dtype=np.int32)
        lib().dh_tree_accessor_get_many(self._h, _u32p(xyz), len(xyz), out.ctypes.data_as(C.POINTER(C.c_int32)))
        return [None if v < 0 else bool(v) for v in out]

    def describe(self, materials: np.ndarray, palette: np.ndarray) -> capi.DustTreeDesc:
        """Fill the DustTreeDesc a Rust caller would build from &tree.root / tree.pool[i].chunks."""
        d = capi.DustTreeDesc()
        self._keep = (np.ascontiguousarray(materials, dtype=np.uint8), np.ascontiguousarray(palette, dtype=np.uint8))
        lib().dh_tree_describe(self._h, self._keep[0].ctypes.data, self._keep[0].size, self._keep[1].ctypes.data, C.byref(d))
        return d


class Pool:
    """crates/vdb/src/pool.rs"""

    def __init__(self, item_bytes, chunk_size_log2):
        self._h = lib().dh_pool_new(item_bytes, chunk_size_log2)

    def __del__(self):
        if self._h:
            lib().dh_pool_free_pool(self._h)
            self._h = None

    def alloc(self):
        return lib().dh_pool_alloc(self._h)

    def free(self, i):
        lib().dh_pool_free(self._h, i)

    def num_chunks(self):
        return lib().dh_pool_num_chunks(self._h)


GPU_VOX_NODE = np.dtype([("x", "<u2"), ("y", "<u2"), ("z", "<u2"), ("w", "<u2"), ("mask", "<u8"),
                         ("material_ptr", "<u4"), ("avg_albedo", "<u4")])
assert GPU_VOX_NODE.itemsize == 24


class VoxModel:
    """(VoxGeometry, PaletteMaterial) of one model after VoxLoader::load_model (loader.rs:238-308)."""

    def __init__(self, handle, vox):
        self._h, self._vox = handle, vox
        self.tree = Tree((4, 2, 2), _borrowed=lib().dh_model_tree(handle), _owner=self)
        n = lib().dh_model_num_materials(handle)
        self.materials = np.frombuffer(C.string_at(lib().dh_model_materials(handle), n), dtype=np.uint8) if n else np.zeros(0, np.uint8)
        nl = lib().dh_model_num_leaves(handle)
        self.nodes = np.frombuffer(C.string_at(lib().dh_model_nodes(handle), nl * 24), dtype=GPU_VOX_NODE) if nl else np.zeros(0, GPU_VOX_NODE)
        self.aabbs = np.frombuffer(C.string_at(lib().dh_model_aabbs(handle), nl * 24), dtype=np.float32).reshape(-1, 6) if nl else np.zeros((0, 6), np.float32)
        s = (C.c_uint8 * 3)()
        lib().dh_model_size(handle, s)
        self.size = tuple(s)

    def __del__(self):
        if getattr(self, "_h", None):
            self.tree = None
            lib().dh_model_free(self._h)
            self._h = None

    def describe(self) -> capi.DustTreeDesc:
        d = capi.DustTreeDesc()
        lib().dh_model_describe(self._h, C.byref(d))
        return d


class VoxFile:
    """A parsed .vox asset (dot_vox::DotVoxData + the loader's scene-graph walk)."""

    def __init__(self, data: bytes):
        self._buf = bytes(data)
        self._h = lib().dh_vox_parse(self._buf, len(self._buf))
        if not self._h:
            raise ValueError(_err())
        self.num_models = lib().dh_vox_num_models(self._h)
        self.palette = np.frombuffer(C.string_at(lib().dh_vox_palette(self._h), 1024), dtype=np.uint8).reshape(256, 4)
        self.instances = []
        for i in range(lib().dh_vox_num_instances(self._h)):
            mid = C.c_uint32()
            m = (C.c_float * 12)()
            lib().dh_vox_instance(self._h, i, C.byref(mid), m)
            self.instances.append((int(mid.value), np.array(m, dtype=np.float32)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().dh_vox_free(self._h)
            self._h = None

    def model_info(self, model_id):
        size = (C.c_uint32 * 3)()
        n = C.c_uint64()
        lib().dh_vox_model_info(self._h, model_id, size, C.byref(n))
        return tuple(size), int(n.value)

    def model_voxels(self, model_id) -> np.ndarray:
        """The model's dot_vox voxels as parsed: [n, 4] uint8 (x, y, z, i) — a view into the parsed file."""
        _, n = self.model_info(model_id)
        if not n:
            return np.zeros((0, 4), np.uint8)
        p = lib().dh_vox_model_voxels(self._h, model_id)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n, 4))

    def load_model(self, model_id) -> VoxModel:
        h = lib().dh_vox_load_model(self._h, model_id)
        if not h:
            raise ValueError(_err())
        return VoxModel(h, self)


def to_transform(translation, rotation, size) -> np.ndarray:
    t = (C.c_int32 * 3)(*translation)
    s = (C.c_uint32 * 3)(*size)
    out = (C.c_float * 12)()
    lib().dh_to_transform(t, rotation, s, out)
    return np.array(out, dtype=np.float32)


class SynthDomain:
    """hierarchy!(4,4,2,2) domain of BASELINE config 5 (dust_host_c.cpp: dh_synth_new)."""

    def __init__(self, seed=5, fill=0.01, domain_leaves=1024):
        self._h = lib().dh_synth_new(seed, fill, domain_leaves)
        if not self._h:
            raise ValueError(_err())
        self.tree = Tree((4, 4, 2, 2), _borrowed=lib().dh_synth_tree(self._h), _owner=self)
        self.num_leaves = lib().dh_synth_num_leaves(self._h)
        self.extent = domain_leaves * 4
        n = lib().dh_synth_num_materials(self._h)
        self._mat_ptr, self._mat_n = lib().dh_synth_materials(self._h), n
        self.palette = np.frombuffer(C.string_at(lib().dh_synth_palette(self._h), 1024), dtype=np.uint8).reshape(256, 4)

    @property
    def materials(self) -> np.ndarray:
        if not self._mat_n:
            return np.zeros(0, np.uint8)
        return np.ctypeslib.as_array(C.cast(self._mat_ptr, C.POINTER(C.c_uint8)), shape=(self._mat_n,))

    def __del__(self):
        if getattr(self, "_h", None):
            self.tree = None
            lib().dh_synth_free(self._h)
            self._h = None

    def describe(self) -> capi.DustTreeDesc:
        d = capi.DustTreeDesc()
        lib().dh_synth_describe(self._h, C.byref(d))
        return d


def sunlight_bake(turbidity=1.0, albedo=(0.2, 0.2, 0.2), direction=(0.0, 0.80114365, -0.5984721)) -> capi.DustSky:
    """Sunlight::bake (crates/render/src/pipeline/sky.rs:90-133); defaults = Sunlight::default() (:15-22)."""
    out = capi.DustSky()
    lib().dh_sunlight_bake(float(turbidity), _f3(albedo), _f3(direction), C.byref(out))
    return out


def camera_look_at(eye, target, up=(0.0, 1.0, 0.0), fov=np.float32(np.pi / 4), near=0.1, far=10000.0,
                   width=1920, height=1080) -> capi.DustCamera:
    """PinholeProjection::default (projection.rs:21-28) + looking_at + CameraSettings (standard.rs:277-302)."""
    out = capi.DustCamera()
    lib().dh_camera_look_at(_f3(eye), _f3(target), _f3(up), float(fov), float(near), float(far), int(width), int(height), C.byref(out))
    return out


def castle_sun_direction(scroll=0.0):
    """examples/castle.rs:260-265 at scroll state `current`."""
    out = (C.c_float * 3)()
    lib().dh_castle_sun_direction(float(scroll), out)
    return tuple(out)
