"""BASELINE.json's full sizes, checked through size-independent properties (the oracle needs minutes for these
frames, so it only sees evenly spread rows / a sample of the batch):
  config 2 (castle stand-in, 1920x1080, 4 spp, primary + 1 bounce):
    * megakernel == wavefront, bit for bit  (two independent schedulings of the same arithmetic)
    * the frame's sample-0 hit buffers == dust_trace_rays on the same camera rays  (two entry points)
    * tmax clipping: re-tracing a hit ray with tmax = t finds the same hit, with tmax = nextbelow(t) nothing
    * 8 shards, gathered and resolved == the whole frame
    * evenly spread rows: hits bit-exact, radiance within 1e-4 against the oracle
  config 5 (4096^3 domain, 1 % of the leaf slots, 10^7 of the 10^8 rays per GPU):
    * tmax clipping as above, and the first rays of the batch against the oracle, bit for bit
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H, SPP = 1920, 1080, 4


@pytest.fixture(scope="module")
def castle(ctx, host):
    from dust_b200 import scenes
    from dust_b200.loader import load_vox
    data, _ = scenes.load_asset("castle")
    loaded = load_vox(ctx, data)
    cam = host.camera_look_at(width=W, height=H, **dict(scenes.CASTLE_CAMERA))
    sky = host.sunlight_bake(turbidity=1.0, albedo=(0.2, 0.2, 0.2), direction=host.castle_sun_direction(0.0))
    return data, loaded, cam, sky


def camera_rays(cam, width, height):
    """primary.rgen:9-21 / camera.glsl:1-16 for the pixel centres (sample 0), in float32 with the kernel's operation order."""
    from dust_b200.pipeline import RAY_DTYPE
    f = np.float32
    px, py = np.meshgrid(np.arange(width, dtype=f), np.arange(height, dtype=f))
    cx = f(2.0) * ((px + f(0.5)) / f(width)) - f(1.0)
    cy = (f(2.0) * ((py + f(0.5)) / f(height)) - f(1.0)) * f(-1.0)
    cx = (cx * (f(width) / f(height))) * f(cam.tan_half_fov)
    cy = cy * f(cam.tan_half_fov)
    c0, c1, c2 = (np.array(getattr(cam, "camera_view_col%d" % k)[:], f) for k in range(3))
    rays = np.zeros(width * height, RAY_DTYPE)
    for a in range(3):
        rays["d"][:, a] = ((c0[a] * cx + c1[a] * cy) + c2[a] * f(-1.0)).ravel()
    rays["o"][:] = np.array([cam.position_x, cam.position_y, cam.position_z], f)
    rays["tmin"], rays["tmax"] = cam.near, cam.far
    return rays


def _clip_property(scene, rays, hits):
    hit = hits["instance"] != 0xFFFFFFFF
    sub = rays[hit].copy()
    sub["tmax"] = hits["t"][hit]
    again = scene.trace_rays(sub)
    assert np.array_equal(again.view(np.uint32), hits[hit].view(np.uint32)), "tmax = t must reproduce the hit (inclusive acceptance)"
    sub["tmax"] = np.nextafter(hits["t"][hit], np.float32(-np.inf))
    below = scene.trace_rays(sub)
    assert (below["instance"] == 0xFFFFFFFF).all(), "nothing may be closer than the closest hit"


def test_config2_full_frame_properties(castle, ctx, oracle):
    import torch
    from dust_b200.pipeline import resolve_shards, shard_layout
    data, loaded, cam, sky = castle
    kw = dict(spp=SPP, max_bounces=1, seed=2)
    wf = loaded.scene.trace(cam, sky, W, H, hits=True, pipeline="wavefront", **kw)
    mk = loaded.scene.trace(cam, sky, W, H, hits=True, pipeline="megakernel", **kw)
    assert np.array_equal(wf["radiance"].view(np.uint32), mk["radiance"].view(np.uint32))
    for k in ("hit_instance", "hit_leaf", "hit_voxel", "hit_t"):
        assert np.array_equal(wf[k].view(np.uint8), mk[k].view(np.uint8)), k
    assert (wf["hit_instance"] != 0xFFFFFFFF).mean() > 0.3        # (the castle camera looks down: every pixel hits)

    # the same primary rays through the ray-batch entry point
    rays = camera_rays(cam, W, H)
    hits = loaded.scene.trace_rays(rays)
    assert np.array_equal(hits["instance"].reshape(H, W), wf["hit_instance"])
    assert np.array_equal(hits["leaf"].reshape(H, W), wf["hit_leaf"])
    assert np.array_equal(hits["t"].view(np.uint32).reshape(H, W), wf["hit_t"].view(np.uint32))
    assert np.array_equal((hits["voxel"] & 0xFF).astype(np.uint8).reshape(H, W), wf["hit_voxel"])
    _clip_property(loaded.scene, rays, hits)

    # 8 shards -> all-gather layout -> resolve == whole frame
    shards = 8
    _, per = shard_layout(W, H, shards)
    gathered = torch.empty((shards, per, 4), dtype=torch.float32, device="cuda")
    ctx.use_torch_stream()
    for s in range(shards):
        loaded.scene.trace_device(cam, sky, W, H, gathered[s], shard_index=s, shard_count=shards, **kw)
    out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    resolve_shards(ctx, W, H, shards, gathered, out)
    torch.cuda.synchronize()
    ctx.set_stream(None)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), wf["radiance"].view(np.uint32))

    # evenly spread rows against the oracle
    ov = oracle.VoxFile(data)
    models = {m: ov.load_model(m) for m in sorted({m for m, _ in ov.instances})}
    osc = oracle.Scene([(mat, models[m].tree._h, models[m].materials, ov.palette) for m, mat in ov.instances])
    rows = [int((i + 0.5) * H / 24) for i in range(24)]           # (bench.py compares ALL 1080 rows with the oracle on every run: `parity`)
    o = osc.render(cam, sky, W, H, rows=rows, **kw)
    for k in ("hit_instance", "hit_leaf", "hit_voxel"):
        assert np.array_equal(wf[k][rows], o[k][rows]), k
    assert np.array_equal(wf["hit_t"][rows].view(np.uint32), o["hit_t"][rows].view(np.uint32))
    a, b = wf["radiance"][rows][..., :3].astype(np.float64), o["radiance"][rows][..., :3].astype(np.float64)
    scale = np.maximum(np.abs(b).max(axis=-1, keepdims=True), 1e-3 * np.abs(b).mean())
    assert (np.abs(a - b) / scale).max() <= 1e-4


def test_config3_4k_sample_batches_against_oracle(castle, ctx, host, oracle):
    """BASELINE.json configs[3]: 3840x2160, 2 bounces, 8 spp.  66 M paths do not fit one wavefront batch (16 Mi paths), so the frame runs
    in four sample batches of 2 spp, carried by the per-pixel accumulator: evenly spread rows against the oracle (ids / t bit-exact,
    radiance 1e-4), and two 32-row bands of the megakernel's frame (no batches at all) bit for bit."""
    from dust_b200 import scenes
    data, loaded, _, sky = castle
    W4, H4 = 3840, 2160
    cam = host.camera_look_at(width=W4, height=H4, **dict(scenes.CASTLE_CAMERA))
    kw = dict(spp=8, max_bounces=2, seed=2)
    launches0 = ctx.stats()["kernel_launches"]
    wf = loaded.scene.trace(cam, sky, W4, H4, hits=True, pipeline="wavefront", **kw)
    assert ctx.stats()["kernel_launches"] - launches0 >= 4 * 10           # four batches of >= 10 launches each
    ov = oracle.VoxFile(data)
    models = {m: ov.load_model(m) for m in sorted({m for m, _ in ov.instances})}
    osc = oracle.Scene([(mat, models[m].tree._h, models[m].materials, ov.palette) for m, mat in ov.instances])
    rows = [int((i + 0.5) * H4 / 8) for i in range(8)]
    o = osc.render(cam, sky, W4, H4, rows=rows, **kw)
    for k in ("hit_instance", "hit_leaf", "hit_voxel"):
        assert np.array_equal(wf[k][rows], o[k][rows]), k
    assert np.array_equal(wf["hit_t"][rows].view(np.uint32), o["hit_t"][rows].view(np.uint32))
    a, b = wf["radiance"][rows][..., :3].astype(np.float64), o["radiance"][rows][..., :3].astype(np.float64)
    scale = np.maximum(np.abs(b).max(axis=-1, keepdims=True), 1e-3 * np.abs(b).mean())
    assert (np.abs(a - b) / scale).max() <= 1e-4
    # one shard of a 64-way split through the megakernel (a different scheduling, no sample batches): its tiles equal the frame's
    from dust_b200.pipeline import shard_layout
    shards, idx = 64, 17
    tiles_per, per = shard_layout(W4, H4, shards)
    part = loaded.scene.trace(cam, sky, W4, H4, shard_index=idx, shard_count=shards, pipeline="megakernel", **kw)["radiance"].reshape(tiles_per, 32, 32, 4)
    tiles_x = (W4 + 31) // 32
    n_tiles = tiles_x * ((H4 + 31) // 32)
    checked = 0
    for lt in range(tiles_per):
        t = lt * shards + idx
        if t >= n_tiles:
            continue
        ty, tx = divmod(t, tiles_x)
        ref = wf["radiance"][ty * 32:ty * 32 + 32, tx * 32:tx * 32 + 32]
        got = part[lt][:ref.shape[0], :ref.shape[1]]
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (lt, t)
        checked += 1
    assert checked >= 120


def test_config5_ray_batch_properties(ctx, host, oracle):
    from dust_b200 import scenes
    from dust_b200.pipeline import Scene
    dom = host.SynthDomain(seed=5, fill=0.01, domain_leaves=1024)
    desc = dom.describe()
    dtree = ctx.upload_tree(desc)
    assert dtree.leaf_count == dom.num_leaves > 10_000_000
    ident = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
    scene = Scene(ctx, [(ident, dtree)])
    n = 10_000_000
    rays = scenes.ray_batch(n, dom.extent, seed=5)
    hits = scene.trace_rays(rays)
    frac = (hits["instance"] != 0xFFFFFFFF).mean()
    assert 0.1 < frac < 0.6
    _clip_property(scene, rays, hits)
    otree = oracle.TreeView(desc)
    osc = oracle.Scene([(ident, otree._h, dom.materials, dom.palette)])
    m = 2_000_000                      # includes rays that start on a leaf plane inside a set voxel (t = +0, canonical)
    o, _ = osc.trace_rays(rays[:m])
    assert np.array_equal(hits[:m].view(np.uint32), o.view(np.uint32))
