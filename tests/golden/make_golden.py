#!/usr/bin/env python3
"""Writes the golden fixtures tests/golden/*.npz.

Provenance (be precise about what these pin): the reference (dust-engine/dust @ 7ea772b) holds NO
golden vector, fixture or test for traversal, hit ids, t or radiance and cannot be
built or run in this environment, so these fixtures are outputs of the repo's own CPU oracle
(oracle/), frozen at the commit that added them.  They are "parity unpinned" with respect to the
reference; what they pin is (a) the oracle against silent drift and (b) the CUDA path against a
committed, machine-independent answer (the -m gpu tests compare against them too).  The only
reference-pinned vectors are the crates/vdb doc-test KATs in tests/test_vdb_kat.py.

Scenes are tiny and generated in code (no asset needed).  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _scene_two_models():
    """Two small models, three instances (one rotated, one mirrored), deterministic content."""
    from dust_b200 import scenes
    rng = np.random.default_rng(42)
    X, Y, Z = np.meshgrid(np.arange(24), np.arange(20), np.arange(16), indexing="ij")
    blob = ((X - 12) ** 2 / 100 + (Y - 10) ** 2 / 64 + (Z - 8) ** 2 / 36 < 1) & ~((X - 12) ** 2 + (Y - 10) ** 2 < 9)
    a = np.stack([*np.nonzero(blob), rng.integers(0, 255, int(blob.sum()))], axis=1)
    slab = (Z[:16, :16, :6] < 2) | ((X[:16, :16, :6] + Y[:16, :16, :6]) % 5 == 0)
    b = np.stack([*np.nonzero(slab), rng.integers(0, 255, int(slab.sum()))], axis=1)
    rots = scenes.all_rotation_bytes(False)
    inst = [(0, (0, 0, 8), 4), (1, (6, 4, -3), rots[9]), (0, (-20, 14, 12), rots[30])]
    return scenes.write_vox([((24, 20, 16), a), ((16, 16, 6), b)], inst, scenes.default_palette(1))


def _built(oracle):
    from conftest import BuiltScene
    return BuiltScene(_scene_two_models(), oracle)


def _camera(s, w, h):
    lo, hi = s.bounds()
    c = (lo + hi) / 2
    return s.host.camera_look_at(c + np.array((0.9, 0.7, 1.1)) * (hi - lo).max(), c, width=w, height=h)


def case_primary(oracle):
    s = _built(oracle)
    r = s.oscene.render(_camera(s, 48, 40), s.host.sunlight_bake(), 48, 40, spp=1, max_bounces=0, seed=9)
    return {k: r[k] for k in ("radiance", "hit_instance", "hit_leaf", "hit_voxel", "hit_t")}


def case_gi(oracle):
    s = _built(oracle)
    r = s.oscene.render(_camera(s, 32, 24), s.host.sunlight_bake(turbidity=2.5, direction=(0.3, 0.7, 0.648)), 32, 24,
                        spp=3, max_bounces=2, seed=21, frame_index=5)
    return {k: r[k] for k in ("radiance", "hit_instance", "hit_leaf", "hit_voxel", "hit_t")}


def case_sky(oracle):
    out = {}
    for i, (turb, alb, d) in enumerate([(1.0, (0.2, 0.2, 0.2), (0.0, 0.80114365, -0.5984721)),
                                        (3.7, (0.1, 0.5, 0.9), (0.5, 0.5, 0.70710678)),
                                        (10.0, (1.0, 1.0, 1.0), (0.0, 1.0, 0.0))]):
        sky = oracle.sunlight_bake(turb, alb, d)
        out[f"bake{i}"] = np.frombuffer(bytes(sky), np.float32).copy()
    return out


def _rays(s, n, seed):
    from dust_b200 import scenes
    lo, hi = s.bounds()
    rays = scenes.ray_batch(n, 1.0, seed=seed)
    rays["o"] = lo.astype(np.float32) - 6 + rays["o"] * (hi - lo + 12).astype(np.float32)
    return rays


def case_rays(oracle):
    """Explicit ray batch: fine hits (hit.rint) and block-granular far-field hits (final_gather/rough.rint, tmin = 8)."""
    s = _built(oracle)
    rays = _rays(s, 4000, 77)
    fine, _ = s.oscene.trace_rays(rays)
    far = rays.copy()
    far["tmin"] = 8.0
    rough, _ = s.oscene.trace_rays(far, rough=True)
    return {"hit_fine": fine.view(np.uint32).reshape(-1, 4), "hit_rough": rough.view(np.uint32).reshape(-1, 4)}


def case_post(oracle):
    """Demodulated YCoCg16F frame + G-buffer of the primary hit -> exposure histogram / adapted luminance -> tone map (sRGB OETF)."""
    s = _built(oracle)
    W, H = 40, 32
    r = s.oscene.render(_camera(s, W, H), s.host.sunlight_bake(), W, H, spp=2, max_bounces=1, seed=13, demodulate=True, gbuffer=True)
    illum = oracle.pack_illuminance(r["radiance"])
    bins, avg = oracle.exposure_update(illum, 0.0)
    ldr = oracle.tone_map(illum, r["gbuffer"]["albedo"], avg)
    g = r["gbuffer"]
    return {"illuminance": illum.view(np.uint16), "hit_albedo": g["albedo"], "hit_normal": g["normal"], "hit_voxel_id": g["voxel_id"],
            "hit_depth": g["depth"].view(np.uint32), "hit_histogram": bins, "adapted_luminance": np.float32(avg).reshape(1), "ldr": ldr}


def case_gi_cache(oracle):
    """Three frames of the reference's own GI (ambient-occlusion / final-gather / surfel passes over the radiance cache, oracle_gi.c) on a
    small cache: keys, sample counts and LRU stamps of the table, the surfel pool and the last frame's illuminance image."""
    s = _built(oracle)
    W, H = 48, 40
    st = oracle.GiState(4096, 512)
    cam, sky = _camera(s, W, H), s.host.sunlight_bake()
    for f in range(3):
        r = s.oscene.gi_frame(cam, sky, W, H, st, seed=31, frame_index=f)
    return {"illuminance": r["illuminance"].view(np.uint16), "hit_cache_fingerprint": st.table["fingerprint"].copy(),
            "hit_cache_sample_count": st.table["sample_count"].copy(), "hit_cache_frame": st.table["last_accessed_frame"].copy(),
            "hit_pool_direction": st.pool["direction"].copy(), "hit_gi_stats": np.array([r["stats"][k] for k in sorted(r["stats"])], np.uint64)}


CASES = {"two_models_gi_cache": case_gi_cache, "two_models_primary": case_primary, "two_models_gi": case_gi, "sky_bake": case_sky, "two_models_rays": case_rays,
         "two_models_post": case_post}


class _Npz(dict):
    @property
    def files(self):
        return list(self.keys())


def load_all():
    out = {}
    for name in CASES:
        with np.load(os.path.join(HERE, name + ".npz")) as z:
            out[name] = _Npz({k: z[k] for k in z.files})
    return out


def main():
    from oracle import pyoracle
    for name, fn in CASES.items():
        data = fn(pyoracle)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
        print(name, {k: v.shape for k, v in data.items()})


if __name__ == "__main__":
    main()
