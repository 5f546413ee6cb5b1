"""VoxLoader::load analogue (crates/vox/src/loader.rs:322-415): .vox bytes -> per-model trees built on
the host (C++ mirror), uploaded once through dust_tree_upload, instanced through dust_scene_create."""
from __future__ import annotations

from . import host
from .pipeline import Context, Scene


class LoadedScene:
    def __init__(self, ctx: Context, data: bytes):
        self.vox = host.VoxFile(data)
        self.instances = self.vox.instances                       # [(model_id, obj2world[12])]
        self.models = {}                                          # model_id -> host.VoxModel
        self.trees = {}                                           # model_id -> DeviceTree
        for mid in sorted({m for m, _ in self.instances}):        # loader.rs:360-372 (distinct model ids)
            self.models[mid] = self.vox.load_model(mid)
            self.trees[mid] = ctx.upload_tree(self.models[mid].describe())
        self.scene = Scene(ctx, [(m, self.trees[mid]) for mid, m in self.instances])

    @property
    def num_voxels(self):
        return sum(len(m.materials) for m in self.models.values())

    @property
    def num_leaves(self):
        return sum(len(m.nodes) for m in self.models.values())

    @property
    def device_bytes(self):
        return sum(t.device_bytes for t in self.trees.values())


def load_vox(ctx: Context, data: bytes) -> LoadedScene:
    return LoadedScene(ctx, data)
