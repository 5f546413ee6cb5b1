"""VoxLoader::load analogue (crates/vox/src/loader.rs:322-415): .vox bytes -> per-model trees built on
the host (C++ mirror), uploaded once through dust_tree_upload, instanced through dust_scene_create."""
from __future__ import annotations

from . import host
from .pipeline import Context, Scene


class LoadedScene:
    """device_build=False: trees built on the host (C++ mirror of the Rust crates) and uploaded (dust_tree_upload).
    device_build=True: only the .vox parse stays on the host; every model's tree is built on the device from its
    voxel list (dust_tree_build_voxels) — same blobs, bytewise (tests/test_device_build.py)."""

    def __init__(self, ctx: Context, data: bytes, device_build: bool = False):
        from .pipeline import DeviceTree
        self.vox = host.VoxFile(data)
        self.instances = self.vox.instances                       # [(model_id, obj2world[12])]
        self.models = {}                                          # model_id -> host.VoxModel (host build only)
        self.trees = {}                                           # model_id -> DeviceTree
        self.sizes = {}
        for mid in sorted({m for m, _ in self.instances}):        # loader.rs:360-372 (distinct model ids)
            if device_build:
                size, _ = self.vox.model_info(mid)
                self.sizes[mid] = size
                self.trees[mid] = DeviceTree.build_voxels(ctx, self.vox.model_voxels(mid), size, self.vox.palette)
            else:
                self.models[mid] = self.vox.load_model(mid)
                self.trees[mid] = ctx.upload_tree(self.models[mid].describe())
        self.scene = Scene(ctx, [(m, self.trees[mid]) for mid, m in self.instances])

    @property
    def num_voxels(self):
        return sum(int(t.layout().n_materials) for t in self.trees.values())

    @property
    def num_leaves(self):
        return sum(t.leaf_count for t in self.trees.values())

    @property
    def device_bytes(self):
        return sum(t.device_bytes for t in self.trees.values())


def load_vox(ctx: Context, data: bytes, device_build: bool = False) -> LoadedScene:
    return LoadedScene(ctx, data, device_build)
