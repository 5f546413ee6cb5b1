"""examples/castle.cpp — the reference's examples/castle.rs frame loop written against the C++ host mirror and the
C ABI only (no Python in the data path).  CPU: it builds, and refuses to run without a device.  GPU: host-built and
device-built scenes give identical frames, and frame 0 equals the same calls made through the ctypes binding."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "castle")


def _build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True, capture_output=True)
    assert os.path.exists(EXE)


def _scene_file(tmp_path):
    from dust_b200 import scenes
    p = tmp_path / "scene.vox"
    p.write_bytes(scenes.castle_like(tiles=2, tile=64, height=128, props=12))
    return str(p)


def test_example_builds_and_needs_a_device(tmp_path):
    import torch
    _build()
    if torch.cuda.is_available():
        pytest.skip("covered by the gpu test")
    r = subprocess.run([EXE, _scene_file(tmp_path), str(tmp_path / "out"), "1", "32", "32"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


def _read_ppm(path):
    data = open(path, "rb").read()
    header, rest = data.split(b"\n255\n", 1)
    w, h = (int(v) for v in header.split()[1:3])
    return np.frombuffer(rest, np.uint8).reshape(h, w, 3)


@pytest.mark.gpu
def test_example_frames(tmp_path, ctx, host):
    from dust_b200.capi import DUST_FMT_YCOCG16F
    from dust_b200.loader import load_vox
    from dust_b200.pipeline import Exposure
    _build()
    scene_file = _scene_file(tmp_path)
    W, H, frames = 160, 90, 3
    outs = {}
    for mode in ("host", "device"):
        prefix = str(tmp_path / mode)
        r = subprocess.run([EXE, scene_file, prefix, str(frames), str(W), str(H)] + (["--device-build"] if mode == "device" else []),
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert r.stdout.count("adapted_luminance") == frames
        outs[mode] = [_read_ppm(f"{prefix}_{f}.ppm") for f in range(frames)]
    for f in range(frames):
        assert np.array_equal(outs["host"][f], outs["device"][f])
    assert not np.array_equal(outs["host"][0], outs["host"][2])                 # the instance moved, exposure adapted
    # frame 0 through the ctypes binding
    loaded = load_vox(ctx, open(scene_file, "rb").read())
    cam = host.camera_look_at((122.0, 300.61, 54.45), (0, 0, 0), width=W, height=H)
    sky = host.sunlight_bake()
    g = loaded.scene.trace(cam, sky, W, H, spp=2, max_bounces=1, seed=0, frame_index=0, fmt=DUST_FMT_YCOCG16F, demodulate=True, gbuffer=True)
    ex = Exposure(ctx)
    ex.update(g["radiance"], W, H)
    ldr = ex.tone_map(g["radiance"], g["gbuffer"]["albedo"], W, H, np.empty((H, W, 4), np.float32))
    want = (np.clip(ldr[..., :3], 0.0, 1.0) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)
    assert np.array_equal(want, outs["host"][0])
    ex.close()
    # --gi: the same loop with the reference's own GI; frames equal the ctypes path (same cache sizes, seeds and frame indices)
    from dust_b200.pipeline import RadianceCache
    prefix = str(tmp_path / "gi")
    r = subprocess.run([EXE, scene_file, prefix, "2", str(W), str(H), "--gi"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.count("gi frame") == 2 and "cache hits 0" not in r.stdout.splitlines()[-2]
    cache = RadianceCache(ctx, 1 << 20, 16384)
    g = cache.gi_frame(loaded.scene, cam, sky, W, H, seed=0, frame_index=0)
    ex = Exposure(ctx)
    ex.update(g["illuminance"], W, H)
    ldr = ex.tone_map(g["illuminance"], g["gbuffer"]["albedo"], W, H, np.empty((H, W, 4), np.float32))
    want = (np.clip(ldr[..., :3], 0.0, 1.0) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)
    assert np.array_equal(want, _read_ppm(f"{prefix}_0.ppm"))
    ex.close(); cache.close()
