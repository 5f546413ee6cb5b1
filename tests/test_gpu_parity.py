"""GPU parity tests proper: the CUDA path, called through the C ABI (include/dust_b200.h), against
the CPU oracle on the same seeded inputs.  Bar: hit instance / leaf / voxel ids and t BIT-EXACT;
radiance within 1e-4 relative (BASELINE.json north_star) — the tolerance covers CUDA's
expf/powf/acosf vs glibc's in the sky model only; every ray path is reproduced exactly.
"""
import numpy as np
import pytest

from conftest import BuiltScene

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _cam_for(scene, host, width, height, eye_dir=(1.2, 0.9, 1.2)):
    lo, hi = scene.bounds()
    c = (lo + hi) / 2
    eye = c + np.array(eye_dir) * (hi - lo).max()
    return host.camera_look_at(eye, c, width=width, height=height)


def _assert_hits_equal(g, o):
    for k in ("hit_instance", "hit_leaf", "hit_voxel"):
        bad = np.argwhere(g[k] != o[k])
        detail = [(tuple(p), {q: (g[q][tuple(p)].item(), o[q][tuple(p)].item()) for q in ("hit_instance", "hit_leaf", "hit_voxel", "hit_t")})
                  for p in bad[:3]]
        assert len(bad) == 0, f"{k}: {len(bad)} pixels differ (gpu, oracle): {detail}"
    assert np.array_equal(g["hit_t"].view(np.uint32), o["hit_t"].view(np.uint32)), "hit_t not bit-exact"


def _assert_radiance_close(g, o, rtol=RTOL):
    a, b = g["radiance"][..., :3].astype(np.float64), o["radiance"][..., :3].astype(np.float64)
    # relative to the pixel's own magnitude (luminance floor 1e-3 of the frame mean avoids 0/0 in black pixels)
    scale = np.maximum(np.abs(b).max(axis=-1, keepdims=True), 1e-3 * np.abs(b).mean())
    err = np.abs(a - b) / scale
    assert err.max() <= rtol, f"radiance max rel err {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}"


@pytest.fixture(scope="module")
def teapot(oracle, ctx):
    from dust_b200 import scenes
    return BuiltScene(scenes.teapot_like(), oracle, ctx)


@pytest.fixture(scope="module")
def castle_small(oracle, ctx):
    from dust_b200 import scenes
    return BuiltScene(scenes.castle_like(tiles=2, tile=64, height=128, props=12), oracle, ctx)


def test_teapot_primary_hits_bit_exact(teapot, host):
    """BASELINE config 1 geometry (teapot stand-in, 256x256, primary only, 1 spp) on the GPU."""
    W = H = 256
    cam, sky = _cam_for(teapot, host, W, H), host.sunlight_bake()
    g = teapot.loaded.scene.trace(cam, sky, W, H, spp=1, max_bounces=0, hits=True)
    o = teapot.oscene.render(cam, sky, W, H, spp=1, max_bounces=0)
    assert (o["hit_instance"] != 0xFFFFFFFF).mean() > 0.05
    _assert_hits_equal(g, o)
    _assert_radiance_close(g, o)


PIPELINES = ["megakernel", "wavefront"]


@pytest.mark.parametrize("pipeline", PIPELINES)
@pytest.mark.parametrize("spp,bounces", [(1, 1), (4, 1), (2, 2), (3, 0)])
def test_teapot_bounces_radiance(teapot, host, spp, bounces, pipeline):
    W = H = 128
    cam, sky = _cam_for(teapot, host, W, H), host.sunlight_bake()
    g = teapot.loaded.scene.trace(cam, sky, W, H, spp=spp, max_bounces=bounces, seed=7, frame_index=3, hits=True, pipeline=pipeline)
    o = teapot.oscene.render(cam, sky, W, H, spp=spp, max_bounces=bounces, seed=7, frame_index=3)
    _assert_hits_equal(g, o)
    _assert_radiance_close(g, o)


def test_castle_small_instances_bit_exact(castle_small, host):
    """Many instances with all 24 rotations + mirrors; the castle.rs camera and zenith sun."""
    from dust_b200 import scenes
    W, H = 320, 180
    cam = host.camera_look_at(width=W, height=H, **{k: v for k, v in scenes.CASTLE_CAMERA.items()})
    sky = host.sunlight_bake(direction=host.castle_sun_direction(0.0))
    g = castle_small.loaded.scene.trace(cam, sky, W, H, spp=1, max_bounces=0, hits=True)
    o = castle_small.oscene.render(cam, sky, W, H, spp=1, max_bounces=0)
    assert (o["hit_instance"] != 0xFFFFFFFF).mean() > 0.2
    assert len(np.unique(o["hit_instance"])) > 3
    _assert_hits_equal(g, o)
    _assert_radiance_close(g, o)


@pytest.mark.parametrize("pipeline", PIPELINES)
def test_castle_small_gi(castle_small, host, pipeline):
    W, H = 160, 96
    cam = _cam_for(castle_small, host, W, H, eye_dir=(0.7, 0.5, 0.9))
    sky = host.sunlight_bake()
    g = castle_small.loaded.scene.trace(cam, sky, W, H, spp=2, max_bounces=2, seed=11, hits=True, stats=True, pipeline=pipeline)
    o = castle_small.oscene.render(cam, sky, W, H, spp=2, max_bounces=2, seed=11)
    _assert_hits_equal(g, o)
    _assert_radiance_close(g, o)
    # ray counts are order-independent and must agree exactly
    for k in ("primary_rays", "bounce_rays", "shadow_rays", "hits_shaded"):
        assert g["stats"][k] == o["stats"][k], k


@pytest.mark.parametrize("pipeline", PIPELINES)
def test_gbuffer_bit_exact(castle_small, host, pipeline):
    """DustGBuffer (hit.rchit:57-94, miss.rmiss:13-16): albedo / normal (A2R10G10B10), depth, motion (RGBA16F, with
    last-frame transforms that differ from this frame's) and voxel_id, bit for bit against the oracle."""
    from dust_b200 import scenes
    W, H = 256, 144
    cam = host.camera_look_at(width=W, height=H, **{k: v for k, v in scenes.CASTLE_CAMERA.items()})
    sky = host.sunlight_bake(direction=host.castle_sun_direction(0.0))
    cur = np.stack([m for _, m in castle_small.ovox.instances]).astype(np.float32).reshape(-1, 12)
    prev = cur.copy()
    rng = np.random.default_rng(3)
    prev[:, 3::4] += rng.normal(scale=0.7, size=(len(prev), 3)).astype(np.float32)
    castle_small.oscene.set_previous_transforms(prev)
    castle_small.loaded.scene.set_previous_transforms(prev)
    try:
        g = castle_small.loaded.scene.trace(cam, sky, W, H, spp=2, max_bounces=1, seed=6, hits=True, gbuffer=True, pipeline=pipeline)
        o = castle_small.oscene.render(cam, sky, W, H, spp=2, max_bounces=1, seed=6, gbuffer=True)
    finally:
        castle_small.oscene.set_previous_transforms(cur)
        castle_small.loaded.scene.set_previous_transforms(cur)
    _assert_hits_equal(g, o)
    assert (o["hit_instance"] != 0xFFFFFFFF).mean() > 0.2
    for k in ("albedo", "normal", "voxel_id"):
        bad = np.argwhere(g["gbuffer"][k] != o["gbuffer"][k])
        assert len(bad) == 0, (k, len(bad), [(tuple(p), hex(g["gbuffer"][k][tuple(p)]), hex(o["gbuffer"][k][tuple(p)])) for p in bad[:3]])
    assert np.array_equal(g["gbuffer"]["depth"].view(np.uint32), o["gbuffer"]["depth"].view(np.uint32))
    assert np.array_equal(g["gbuffer"]["motion"].view(np.uint16), o["gbuffer"]["motion"].view(np.uint16))
    assert np.abs(o["gbuffer"]["motion"].astype(np.float32)).max() > 0.3


def test_moving_instances_update_transforms(castle_small, host, oracle):
    """dust_scene_update_transforms (the moving teapot of examples/castle.rs:287-292): instance records and the
    instance BVH are rebuilt for new matrices; hits equal an oracle scene built with those matrices, and the motion
    vectors point back to the transforms that were in force before the update."""
    from dust_b200 import scenes
    W, H = 192, 108
    cam = host.camera_look_at(width=W, height=H, **{k: v for k, v in scenes.CASTLE_CAMERA.items()})
    sky = host.sunlight_bake(direction=host.castle_sun_direction(0.0))
    old = np.stack([m for _, m in castle_small.ovox.instances]).astype(np.float32).reshape(-1, 12)
    new = old.copy()
    rng = np.random.default_rng(11)
    new[:, 3::4] += np.round(rng.normal(scale=6.0, size=(len(new), 3))).astype(np.float32)      # every instance moves a few voxels
    oscene = oracle.Scene([(new[i], castle_small.omodels[m].tree._h, castle_small.omodels[m].materials, castle_small.ovox.palette)
                           for i, (m, _) in enumerate(castle_small.ovox.instances)])
    oscene.set_previous_transforms(old)
    scene = castle_small.loaded.scene
    scene.update_transforms(new)
    try:
        g = scene.trace(cam, sky, W, H, spp=2, max_bounces=1, seed=4, hits=True, gbuffer=True)
        o = oscene.render(cam, sky, W, H, spp=2, max_bounces=1, seed=4, gbuffer=True)
    finally:
        scene.update_transforms(old)
        scene.set_previous_transforms(old)
    _assert_hits_equal(g, o)
    _assert_radiance_close(g, o)
    assert np.array_equal(g["gbuffer"]["motion"].view(np.uint16), o["gbuffer"]["motion"].view(np.uint16))
    assert np.abs(o["gbuffer"]["motion"].astype(np.float32)).max() > 1.0
    # and the scene is back where it was
    g0 = scene.trace(cam, sky, W, H, hits=True)
    o0 = castle_small.oscene.render(cam, sky, W, H)
    _assert_hits_equal(g0, o0)


def test_visit_counters_against_oracle(teapot, host):
    """Ray / hit counts are properties of the paths and must match the oracle exactly.  The visit
    counters W, P are properties of the ALGORITHM: the kernel culls with the padded
    occupied-bounds box before entering an instance, the oracle walks the whole 256^3 box, so the
    kernel may only read fewer words, never more; the leaf-level work (P) is nearly identical."""
    W = H = 96
    cam, sky = _cam_for(teapot, host, W, H), host.sunlight_bake()
    g = teapot.loaded.scene.trace(cam, sky, W, H, spp=1, max_bounces=1, seed=3, stats=True)
    o = teapot.oscene.render(cam, sky, W, H, spp=1, max_bounces=1, seed=3)
    for k in ("primary_rays", "bounce_rays", "shadow_rays", "hits_shaded"):
        assert g["stats"][k] == o["stats"][k], (k, g["stats"][k], o["stats"][k])
    assert 0 < g["stats"]["mask_words"] <= o["stats"]["mask_words"]
    assert 0 < g["stats"]["child_ptrs"] <= o["stats"]["child_ptrs"]
    assert g["stats"]["child_ptrs"] > 0.9 * o["stats"]["child_ptrs"]


def test_fp16_output_and_device_buffers(teapot, host, ctx):
    import torch
    from dust_b200.capi import DUST_FMT_RGBA16F
    W = H = 64
    cam, sky = _cam_for(teapot, host, W, H), host.sunlight_bake()
    ref = teapot.loaded.scene.trace(cam, sky, W, H, spp=1, max_bounces=1)["radiance"]
    h16 = teapot.loaded.scene.trace(cam, sky, W, H, spp=1, max_bounces=1, fmt=DUST_FMT_RGBA16F)["radiance"]
    assert np.array_equal(h16, ref.astype(np.float16))
    dev = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    ctx.use_torch_stream()
    teapot.loaded.scene.trace_device(cam, sky, W, H, dev, spp=1, max_bounces=1)
    torch.cuda.synchronize()
    ctx.set_stream(None)
    assert np.array_equal(dev.cpu().numpy(), ref)


def test_wavefront_sample_batches_equal_megakernel(castle_small, host, monkeypatch):
    """The wavefront pipeline splits the samples of a frame into batches when the path state would not
    fit its budget; pixel sums must still be accumulated in sample order: bit-identical frames."""
    W, H = 200, 120                                   # ragged tiles: 7 x 4 tiles of 32x32
    cam = _cam_for(castle_small, host, W, H, eye_dir=(0.6, 0.8, 0.7))
    sky = host.sunlight_bake()
    kw = dict(spp=5, max_bounces=2, seed=13, frame_index=2)
    mega = castle_small.loaded.scene.trace(cam, sky, W, H, pipeline="megakernel", **kw)["radiance"]
    one = castle_small.loaded.scene.trace(cam, sky, W, H, pipeline="wavefront", **kw)["radiance"]
    monkeypatch.setenv("DUST_WF_MAX_PATHS", str(28 * 1024 * 2))      # 2 samples per batch -> batches of 2, 2, 1
    many = castle_small.loaded.scene.trace(cam, sky, W, H, pipeline="wavefront", **kw)["radiance"]
    assert np.array_equal(one, mega)
    assert np.array_equal(many, mega)


@pytest.mark.parametrize("pipeline", PIPELINES)
@pytest.mark.parametrize("shards", [2, 3, 8])
def test_sharded_frame_equals_whole_frame(castle_small, host, ctx, shards, pipeline):
    """Tile-sharded render + resolve == single-GPU frame, bit for bit (ragged 200x120 frame)."""
    import torch
    from dust_b200.pipeline import resolve_shards, shard_layout
    W, H = 200, 120
    cam = _cam_for(castle_small, host, W, H, eye_dir=(0.6, 0.8, 0.7))
    sky = host.sunlight_bake()
    whole = castle_small.loaded.scene.trace(cam, sky, W, H, spp=2, max_bounces=1, seed=5)["radiance"]
    _, per = shard_layout(W, H, shards)
    gathered = torch.zeros((shards, per, 4), dtype=torch.float32, device="cuda")
    for s in range(shards):
        castle_small.loaded.scene.trace_device(cam, sky, W, H, gathered[s], spp=2, max_bounces=1, seed=5,
                                               shard_index=s, shard_count=shards, pipeline=pipeline)
    out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    resolve_shards(ctx, W, H, shards, gathered, out)
    ctx.synchronize()
    assert np.array_equal(out.cpu().numpy(), whole)


def test_comm_allgather_single_rank(ctx):
    """dust_comm_* / dust_allgather_framebuffer with a one-rank communicator: exercises the run-time NCCL binding
    (dlopen, unique id, ncclCommInitRank, ncclAllGather on the context's stream) on one GPU."""
    import torch
    from dust_b200.pipeline import Context, shard_layout
    uid = Context.comm_unique_id()
    assert len(uid) == 128 and any(uid)
    ctx.comm_init(uid, 0, 1)
    try:
        _, per = shard_layout(200, 120, 1)
        ctx.use_torch_stream()
        shard = torch.rand((per, 4), dtype=torch.float32, device="cuda")
        gathered = torch.zeros((1, per, 4), dtype=torch.float32, device="cuda")
        ctx.allgather_framebuffer(shard, gathered)
        torch.cuda.synchronize()
        assert torch.equal(gathered[0], shard)
        with pytest.raises(Exception):
            ctx.comm_init(uid, 0, 1)                        # one communicator per context
    finally:
        ctx.set_stream(None)
        ctx.comm_destroy()
    ctx.comm_destroy()                                      # idempotent


def test_ray_batch_small_domain(oracle, host, ctx):
    """BASELINE config 5 in miniature: hierarchy!(4,4,2,2), sphere-union leaf fill, random rays."""
    from dust_b200 import scenes
    from dust_b200.pipeline import Scene
    dom = host.SynthDomain(seed=5, fill=0.02, domain_leaves=96)
    assert dom.num_leaves > 1000
    desc = dom.describe()
    dtree = ctx.upload_tree(desc)
    assert dtree.leaf_count == dom.num_leaves
    ident = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
    scene = Scene(ctx, [(ident, dtree)])
    rays = scenes.ray_batch(200_000, dom.extent, seed=5)
    g = scene.trace_rays(rays)
    otree = oracle.TreeView(desc)
    oscene = oracle.Scene([(ident, otree._h, dom.materials, dom.palette)])
    o, _ = oscene.trace_rays(rays)
    assert (o["instance"] != 0xFFFFFFFF).mean() > 0.03
    assert np.array_equal(g.view(np.uint32), o.view(np.uint32))


def test_ray_batch_origins_on_lattice_planes(oracle, host, ctx):
    """Rays that start ON voxel / leaf / node planes (integer origins, tmin = 0), many of them inside set voxels:
    t = 0 hits must come out as the canonical +0.0 on both sides (max(-0, +0) is not defined by IEEE), and
    plane-time ties at the very first step must not change a result."""
    from dust_b200 import scenes
    from dust_b200.pipeline import Scene
    dom = host.SynthDomain(seed=7, fill=0.05, domain_leaves=64)
    desc = dom.describe()
    scene = Scene(ctx, [(np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32), ctx.upload_tree(desc))])
    rays = scenes.ray_batch(400_000, dom.extent, seed=9)
    rays["o"] = np.round(rays["o"])                                   # every origin on the voxel lattice
    rays["o"][::2] = np.round(rays["o"][::2] / 4) * 4                 # every other one on the leaf lattice
    rays["d"][::5, 1] = 0.0                                           # and some axis-parallel / zero components
    g = scene.trace_rays(rays)
    otree = oracle.TreeView(desc)
    o, _ = oracle.Scene([(np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32), otree._h, dom.materials, dom.palette)]).trace_rays(rays)
    zero = (o["instance"] != 0xFFFFFFFF) & (o["t"] == 0.0)
    assert zero.sum() > 1000
    assert (o["t"][zero].view(np.uint32) == 0).all()                  # +0.0, never -0.0
    assert np.array_equal(g.view(np.uint32), o.view(np.uint32))


def test_ray_batch_many_instances(castle_small, host, monkeypatch):
    """dust_trace_rays on a multi-instance scene: pre-culled candidate lists + persistent lanes (and the
    one-thread-per-ray kernel, DUST_RAYS_PLAIN) against the oracle's linear instance loop, bit for bit."""
    from dust_b200 import scenes
    lo, hi = castle_small.bounds()
    rays = scenes.ray_batch(150_000, 1.0, seed=8)
    rays["o"] = lo.astype(np.float32) - 20 + rays["o"] * (hi - lo + 40).astype(np.float32)
    rays["tmax"][::3] = 150.0                               # mix of bounded and unbounded rays
    o, _ = castle_small.oscene.trace_rays(rays)
    assert 0.1 < (o["instance"] != 0xFFFFFFFF).mean() < 0.95
    g = castle_small.loaded.scene.trace_rays(rays)
    assert np.array_equal(g.view(np.uint32), o.view(np.uint32))
    monkeypatch.setenv("DUST_RAYS_PLAIN", "1")
    g2 = castle_small.loaded.scene.trace_rays(rays)
    assert np.array_equal(g2.view(np.uint32), o.view(np.uint32))
    monkeypatch.delenv("DUST_RAYS_PLAIN")
    monkeypatch.setenv("DUST_RAYS_CHUNK", "20000")           # host path: 8 chunks ping-pong over the two side streams (ragged last chunk)
    g3 = castle_small.loaded.scene.trace_rays(rays)
    assert np.array_equal(g3.view(np.uint32), o.view(np.uint32))


def test_empty_scene_and_errors(host, ctx):
    from dust_b200.capi import DustError
    from dust_b200.pipeline import Scene
    W = H = 16
    cam = host.camera_look_at((0, 0, 10), (0, 0, 0), width=W, height=H)
    sky = host.sunlight_bake()
    empty = Scene(ctx, [])
    r = empty.trace(cam, sky, W, H, hits=True)
    assert (r["hit_instance"] == 0xFFFFFFFF).all() and np.isinf(r["hit_t"]).all()
    assert np.isfinite(r["radiance"]).all()
    with pytest.raises(DustError):
        empty.trace(cam, sky, 0, 16)
    with pytest.raises(DustError):
        empty.trace(cam, sky, W, H, shard_index=2, shard_count=2)


def test_cuda_matches_committed_golden_fixtures(oracle, host, ctx):
    """The CUDA path against tests/golden/*.npz (frozen oracle outputs): ids and t bit-exact, radiance 1e-4."""
    import os
    import sys
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    if gdir not in sys.path:
        sys.path.insert(0, gdir)
    import make_golden
    from dust_b200.loader import load_vox
    want = make_golden.load_all()
    loaded = load_vox(ctx, make_golden._scene_two_models())
    s = BuiltScene(make_golden._scene_two_models(), oracle)
    g = loaded.scene.trace(make_golden._camera(s, 48, 40), host.sunlight_bake(), 48, 40, spp=1, max_bounces=0, seed=9, hits=True)
    _assert_hits_equal(g, want["two_models_primary"])
    _assert_radiance_close(g, want["two_models_primary"])
    g = loaded.scene.trace(make_golden._camera(s, 32, 24), host.sunlight_bake(turbidity=2.5, direction=(0.3, 0.7, 0.648)), 32, 24,
                           spp=3, max_bounces=2, seed=21, frame_index=5, hits=True)
    _assert_hits_equal(g, want["two_models_gi"])
    _assert_radiance_close(g, want["two_models_gi"])
    # ray batch: fine and block-granular hits, bit for bit
    rays = make_golden._rays(s, 4000, 77)
    assert np.array_equal(loaded.scene.trace_rays(rays).view(np.uint32).reshape(-1, 4), want["two_models_rays"]["hit_fine"])
    far = rays.copy()
    far["tmin"] = 8.0
    assert np.array_equal(loaded.scene.trace_rays(far, rough=True).view(np.uint32).reshape(-1, 4), want["two_models_rays"]["hit_rough"])
    # demodulated frame + G-buffer -> exposure -> tone map
    from dust_b200.capi import DUST_FMT_YCOCG16F
    from dust_b200.pipeline import Exposure
    wp = want["two_models_post"]
    W, H = 40, 32
    g = loaded.scene.trace(make_golden._camera(s, W, H), host.sunlight_bake(), W, H, spp=2, max_bounces=1, seed=13, fmt=DUST_FMT_YCOCG16F,
                           demodulate=True, gbuffer=True)
    for k, gk in (("hit_albedo", "albedo"), ("hit_normal", "normal"), ("hit_voxel_id", "voxel_id")):
        assert np.array_equal(g["gbuffer"][gk], wp[k]), k
    assert np.array_equal(g["gbuffer"]["depth"].view(np.uint32), wp["hit_depth"])
    a, b = g["radiance"].astype(np.float32), wp["illuminance"].view(np.float16).astype(np.float32)
    assert np.allclose(a, b, rtol=1.5e-3, atol=1e-6)
    ex = Exposure(ctx)
    avg = ex.update(wp["illuminance"].view(np.float16), W, H)                  # the golden image itself: bins are then exact
    assert abs(avg - float(wp["adapted_luminance"][0])) <= 1e-6 * avg
    ldr = ex.tone_map(wp["illuminance"].view(np.float16), wp["hit_albedo"], W, H, np.empty((H, W, 4), np.float32))
    assert np.abs(ldr - wp["ldr"]).max() <= 1e-4
    ex.close()
    # the reference's own GI: three frames over a small radiance cache
    from dust_b200.pipeline import RadianceCache
    wc = want["two_models_gi_cache"]
    cache = RadianceCache(ctx, 4096, 512)
    W, H = 48, 40
    for f in range(3):
        g = cache.gi_frame(loaded.scene, make_golden._camera(s, W, H), host.sunlight_bake(), W, H, seed=31, frame_index=f)
    t, p = cache.download()
    assert np.array_equal(t["fingerprint"], wc["hit_cache_fingerprint"]) and np.array_equal(t["sample_count"], wc["hit_cache_sample_count"])
    assert np.array_equal(t["last_accessed_frame"], wc["hit_cache_frame"]) and np.array_equal(p["direction"], wc["hit_pool_direction"])
    a, b = g["illuminance"].astype(np.float32), wc["illuminance"].view(np.float16).astype(np.float32)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin) and np.allclose(a[fin], b[fin], rtol=1.5e-3, atol=1e-6)
    cache.close()


def test_launch_info_timing_and_counter_deltas(castle_small, host):
    """DUST_TRACE_TIMING: one record per launch of the call; with STATS the per-launch counter deltas add up to the call's totals."""
    import torch
    from dust_b200 import scenes
    W, H = 256, 144
    cam = host.camera_look_at(width=W, height=H, **{k: v for k, v in scenes.CASTLE_CAMERA.items()})
    sky = host.sunlight_bake(direction=host.castle_sun_direction(0.0))
    ctx = castle_small.loaded.scene.ctx
    out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    ctx.use_torch_stream()
    try:
        for pipeline, names in (("wavefront", {"wf_tile_cull", "wf_sun_grid", "wf_trace", "wf_shade", "wf_cull", "wf_trace_pt", "wf_resolve", "wf_trace_slow"}),
                                ("megakernel", {"wf_tile_cull", "frame_kernel"})):
            castle_small.loaded.scene.trace_device(cam, sky, W, H, out, spp=2, max_bounces=1, seed=1, stats=True, timing=True, pipeline=pipeline)
            info = ctx.launch_info()
            tot = ctx.stats()
            assert {r["kernel"] for r in info} == names and all(r["ms"] > 0 for r in info)
            assert len(info) == (13 if pipeline == "wavefront" else 2)          # + wf_tile_cull, wf_sun_grid (the candidate lists of camera / sun rays)
            for k in ("primary_rays", "bounce_rays", "shadow_rays", "mask_words", "child_ptrs", "hits_shaded"):
                assert sum(r[k] for r in info) == tot[k], k
        castle_small.loaded.scene.trace_device(cam, sky, W, H, out, spp=2, max_bounces=1, seed=1)     # no flag: the records of the last timed call stay
        assert len(ctx.launch_info()) == 2
    finally:
        ctx.set_stream(None)


def _mixed_scene(castle_small, host, oracle, ctx, extra=()):
    """castle_small's 3-level instances + a 4-level synthetic domain + `extra` (matrix, model id) instances, built twice."""
    from dust_b200.pipeline import Scene
    dom = host.SynthDomain(seed=11, fill=0.03, domain_leaves=48)
    desc = dom.describe()
    dtree = ctx.upload_tree(desc)
    otree = oracle.TreeView(desc)
    dom_m = np.array([1, 0, 0, -96, 0, 1, 0, 40, 0, 0, 1, -96], np.float32)                   # above the castle
    ginst = [(m, castle_small.loaded.trees[mid]) for mid, m in castle_small.loaded.instances] + [(dom_m, dtree)]
    oinst = [(m, castle_small.omodels[mid].tree._h, castle_small.omodels[mid].materials, castle_small.ovox.palette)
             for mid, m in castle_small.ovox.instances] + [(dom_m, otree._h, dom.materials, dom.palette)]
    for m, mid in extra:
        ginst.append((m, castle_small.loaded.trees[mid]))
        oinst.append((m, castle_small.omodels[mid].tree._h, castle_small.omodels[mid].materials, castle_small.ovox.palette))
    keep = (dom, desc, dtree, otree)
    return Scene(ctx, ginst), oracle.Scene(oinst), keep


@pytest.mark.parametrize("pipeline", PIPELINES)
def test_mixed_three_and_four_level_instances(castle_small, host, oracle, ctx, pipeline):
    """A scene whose instances mix hierarchy!(4,2,2) and hierarchy!(4,4,2,2) trees runs the NL = 0 kernel instantiations
    (levels decided per instance), and a sheared / non-uniformly scaled instance exercises the general 3x4 inverse."""
    from dust_b200 import scenes
    mid0 = castle_small.ovox.instances[0][0]
    skew = np.array([1.5, 0.25, 0, 60, 0, 0.75, 0.1, 90, 0.2, 0, 1.25, -40], np.float32)
    gscene, oscene, keep = _mixed_scene(castle_small, host, oracle, ctx, extra=[(skew, mid0)])
    W, H = 224, 126
    cam = host.camera_look_at(width=W, height=H, **{k: v for k, v in scenes.CASTLE_CAMERA.items()})
    sky = host.sunlight_bake(direction=host.castle_sun_direction(0.0))
    g = gscene.trace(cam, sky, W, H, spp=2, max_bounces=2, seed=9, hits=True, pipeline=pipeline)
    o = oscene.render(cam, sky, W, H, spp=2, max_bounces=2, seed=9)
    _assert_hits_equal(g, o)
    _assert_radiance_close(g, o)
    n_inst = len(castle_small.ovox.instances)
    assert (o["hit_instance"] == n_inst).sum() > 50 and (o["hit_instance"] == n_inst + 1).sum() > 50     # the 4-level and the skewed instance are seen
    lo, hi = castle_small.bounds()
    rays = scenes.ray_batch(120_000, 1.0, seed=21)
    rays["o"] = (lo - 30).astype(np.float32) + rays["o"] * (hi - lo + 60).astype(np.float32)
    og, _ = oscene.trace_rays(rays)
    assert np.array_equal(gscene.trace_rays(rays).view(np.uint32), og.view(np.uint32))


def test_two_thousand_instances(teapot, host, oracle, ctx):
    """2 000 instances of one tree on a jittered grid with the 24 axis rotations: 11 instance-id bits, a deep SAH BVH,
    candidate lists that overflow the pre-cull record (rays along the grid rows touch dozens of boxes)."""
    from dust_b200 import scenes
    from dust_b200.pipeline import Scene
    rng = np.random.default_rng(5)
    rots = []
    for perm in ((0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)):
        for signs in ((1, 1, 1), (1, -1, -1), (-1, 1, -1), (-1, -1, 1)):
            R = np.zeros((3, 3), np.float32)
            for r in range(3):
                R[r, perm[r]] = signs[r]
            if np.linalg.det(R) < 0:
                R[2] = -R[2]
            rots.append(R)
    mats = []
    for k in range(2000):
        gx, gy, gz = k % 20, (k // 20) % 10, k // 200
        t = np.array([gx * 90, gy * 70, gz * 80], np.float32) + rng.integers(-6, 7, 3).astype(np.float32)
        M = np.concatenate([rots[k % 24], t[:, None]], axis=1).astype(np.float32).reshape(12)
        mats.append(M)
    mid = teapot.ovox.instances[0][0]
    gscene = Scene(ctx, [(m, teapot.loaded.trees[mid]) for m in mats])
    om = teapot.omodels[mid]
    oscene = oracle.Scene([(m, om.tree._h, om.materials, teapot.ovox.palette) for m in mats])
    rays = scenes.ray_batch(150_000, 1.0, seed=31)
    rays["o"] = np.array([-100, -100, -100], np.float32) + rays["o"] * np.array([2000, 900, 1000], np.float32)
    rays["d"][::4, 1:] *= 0.02                                   # a quarter of the rays run along the grid rows
    o, _ = oscene.trace_rays(rays)
    assert 0.2 < (o["instance"] != 0xFFFFFFFF).mean() < 0.99 and o["instance"][o["instance"] != 0xFFFFFFFF].max() > 1500
    assert np.array_equal(gscene.trace_rays(rays).view(np.uint32), o.view(np.uint32))
    W, H = 160, 90
    cam = host.camera_look_at((-150, 500, -200), (900, 300, 400), width=W, height=H)
    sky = host.sunlight_bake()
    for pipeline in PIPELINES:
        g = gscene.trace(cam, sky, W, H, spp=2, max_bounces=1, seed=2, hits=True, pipeline=pipeline)
        of = oscene.render(cam, sky, W, H, spp=2, max_bounces=1, seed=2)
        _assert_hits_equal(g, of)
        _assert_radiance_close(g, of)


def test_wide_bvh_far_from_origin_and_mixed_scales(teapot, host, oracle, ctx):
    """The instance BVH's 16-bit quantisation grid (device_layout.h "WIDE BVH") under stress: a scene 65 000 units away from the
    origin (the grid step is then bounded by the floats' own spacing, not by the extent), instances scaled 0.05x .. 20x and sheared,
    boxes that touch and nest, rays from inside, from far outside, exactly axis-parallel ones, and rays that graze box faces.
    The oracle visits every instance in a linear loop, so any candidate the quantised boxes lose shows up as a missing hit."""
    from dust_b200 import scenes
    from dust_b200.pipeline import Scene
    rng = np.random.default_rng(77)
    base = np.array([65000.0, -30000.0, 12000.0], np.float32)
    mats = []
    for k in range(300):
        s = np.float32([0.05, 0.3, 1.0, 1.0, 4.0, 20.0][k % 6])
        A = np.eye(3, dtype=np.float32) * s
        if k % 5 == 0:
            A[0, 1] = 0.3 * s; A[2, 0] = -0.2 * s                # shear
        if k % 7 == 0:
            A = A[[1, 2, 0]]                                     # an axis permutation (rotation)
        t = base + rng.integers(-900, 900, 3).astype(np.float32) * np.float32(0.5 if k % 2 else 3.0)
        if k % 11 == 0 and mats:
            t = mats[-1].reshape(3, 4)[:, 3] + np.float32([256.0 * float(s), 0.0, 0.0])   # touching / nested neighbours
        mats.append(np.concatenate([A, t[:, None]], axis=1).astype(np.float32).reshape(12))
    mid = teapot.ovox.instances[0][0]
    gscene = Scene(ctx, [(m, teapot.loaded.trees[mid]) for m in mats])
    om = teapot.omodels[mid]
    oscene = oracle.Scene([(m, om.tree._h, om.materials, teapot.ovox.palette) for m in mats])
    n = 120_000
    rays = scenes.ray_batch(n, 1.0, seed=78)
    span = np.float32([6000.0, 6000.0, 6000.0])
    rays["o"] = base - span / 2 + rays["o"] * span
    far = slice(0, n // 4)                                       # a quarter start far outside and aim into the scene
    rays["o"][far] = base + (rays["o"][far] - base) * np.float32(40.0)
    aim = base + (rng.random((n // 4, 3)).astype(np.float32) - 0.5) * np.float32(3000.0)
    d = aim - rays["o"][far]
    rays["d"][far] = d / np.linalg.norm(d, axis=1, keepdims=True)
    ax = slice(n // 4, n // 4 + n // 8)                          # exactly axis-parallel rays (two zero components)
    dax = np.zeros((n // 8, 3), np.float32)
    dax[np.arange(n // 8), rng.integers(0, 3, n // 8)] = rng.choice(np.float32([-1.0, 1.0]), n // 8)
    rays["d"][ax] = dax
    graze = slice(n // 2, n // 2 + n // 8)                       # rays in the plane of a box face of some instance
    for j, i in enumerate(range(graze.start, graze.stop)):
        M = mats[j % len(mats)].reshape(3, 4)
        rays["o"][i] = M[:, 3] + np.float32([-50.0, 0.0, 7.0])
        rays["d"][i] = np.float32([1.0, 0.0, 1e-3 * (j % 5)])
    o, _ = oscene.trace_rays(rays)
    hit = o["instance"] != 0xFFFFFFFF
    assert 0.05 < hit.mean() < 0.95 and hit[far].any() and hit[ax].any()
    assert np.array_equal(gscene.trace_rays(rays).view(np.uint32), o.view(np.uint32))
    W, H = 128, 72
    cam = host.camera_look_at(tuple(base + np.float32([2500.0, 1800.0, 2200.0])), tuple(base), width=W, height=H)
    sky = host.sunlight_bake()
    for pipeline in PIPELINES:
        g = gscene.trace(cam, sky, W, H, spp=2, max_bounces=1, seed=3, hits=True, pipeline=pipeline)
        of = oscene.render(cam, sky, W, H, spp=2, max_bounces=1, seed=3)
        _assert_hits_equal(g, of)
        _assert_radiance_close(g, of)


def test_ray_batch_rough_block_hits(castle_small, teapot, host, oracle, ctx):
    """DUST_RAYS_ROUGH (final_gather/rough.rint): block-granular leaves, far-field ray interval [8, far], against the oracle bit for
    bit — multi-instance scene (pre-cull + persistent lanes + slow path) and a single-instance one (no cull pass)."""
    from dust_b200 import scenes
    lo, hi = castle_small.bounds()
    rays = scenes.ray_batch(150_000, 1.0, seed=41)
    rays["o"] = lo.astype(np.float32) - 20 + rays["o"] * (hi - lo + 40).astype(np.float32)
    rays["tmin"] = 8.0
    rays["tmax"][::3] = 200.0
    o, _ = castle_small.oscene.trace_rays(rays, rough=True)
    fine, _ = castle_small.oscene.trace_rays(rays)
    assert 0.1 < (o["instance"] != 0xFFFFFFFF).mean() < 0.95 and not np.array_equal(o.view(np.uint32), fine.view(np.uint32))
    g = castle_small.loaded.scene.trace_rays(rays, rough=True)
    assert np.array_equal(g.view(np.uint32), o.view(np.uint32))
    assert np.array_equal(castle_small.loaded.scene.trace_rays(rays).view(np.uint32), fine.view(np.uint32))     # the flag does not stick
    tl, th = teapot.bounds()
    r2 = scenes.ray_batch(100_000, 1.0, seed=42)
    r2["o"] = tl.astype(np.float32) - 30 + r2["o"] * (th - tl + 60).astype(np.float32)
    o2, _ = teapot.oscene.trace_rays(r2, rough=True)
    assert (o2["instance"] != 0xFFFFFFFF).mean() > 0.02
    assert np.array_equal(teapot.loaded.scene.trace_rays(r2, rough=True).view(np.uint32), o2.view(np.uint32))


def test_frames_in_flight_equal_synchronous_frames(castle_small, host, ctx):
    """dust_trace_async / dust_frame_wait (the reference keeps up to 3 frames in flight, rhyolite_bevy/src/lib.rs:58): five frames
    queued back to back, three in flight at most, each bitwise equal to the synchronous call's frame."""
    import torch
    from dust_b200.capi import DUST_FMT_RGBA16F, DustError
    W, H = 192, 108
    cam, sky = _cam_for(castle_small, host, W, H), host.sunlight_bake()
    scene = castle_small.loaded.scene
    kw = dict(spp=2, max_bounces=1, seed=9, fmt=DUST_FMT_RGBA16F)
    want = [scene.trace(cam, sky, W, H, frame_index=k, **kw)["radiance"].copy() for k in range(5)]
    assert not np.array_equal(want[0], want[1])                      # frame_index reseeds the paths
    bufs = [torch.empty((H, W, 4), dtype=torch.float16).pin_memory() for _ in range(5)]
    tickets = [scene.trace_async(cam, sky, W, H, bufs[k].numpy(), frame_index=k, **kw) for k in range(5)]
    assert tickets == sorted(tickets) and len(set(tickets)) == 5
    for k in (4, 0, 2, 1, 3):                                        # any order; waiting twice is harmless
        scene.frame_wait(tickets[k])
        assert np.array_equal(bufs[k].numpy(), want[k]), k
    scene.frame_wait(tickets[0])
    with pytest.raises(DustError):
        scene.frame_wait(tickets[-1] + 1)                            # a ticket that was never issued


def test_bad_transforms_and_misaligned_device_buffers_are_rejected(teapot, host, ctx):
    import torch
    from dust_b200.capi import DustError
    from dust_b200.pipeline import Scene
    from dust_b200.pipeline import RAY_DTYPE
    tree = next(iter(teapot.loaded.trees.values()))
    ident = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
    singular = ident.copy(); singular[10] = 0.0                      # a flattened instance has no world -> object map
    nan = ident.copy(); nan[3] = np.nan
    for m in (singular, nan):
        with pytest.raises(DustError, match="singular|finite"):
            Scene(ctx, [(m, tree)])
    good = Scene(ctx, [(ident, tree)])
    with pytest.raises(DustError, match="singular|finite"):
        good.update_transforms(singular.reshape(1, 12))
    # the scene keeps its last valid state
    rays = np.zeros(4, dtype=RAY_DTYPE)
    rays["o"] = (-5, 10, 10); rays["d"] = (1, 0.01, 0.02); rays["tmax"] = np.inf
    before = good.trace_rays(rays).copy()
    # 128-bit loads: a device pointer at an 8-byte offset is refused, not faulted on
    d = torch.zeros(4 * 8 + 2, dtype=torch.float32, device="cuda")
    d[2:].view(4, 8).copy_(torch.from_numpy(rays.view(np.float32).reshape(4, 8)))
    hits = torch.empty((4, 4), dtype=torch.int32, device="cuda")
    with pytest.raises(DustError, match="aligned"):
        good.trace_rays_device(d[2:], hits, 4)
    assert np.array_equal(good.trace_rays(rays), before)


def test_tile_candidate_lists_do_not_change_frames(castle_small, teapot, host, ctx, oracle):
    """wf_tile_cull: per-tile frustum candidate lists replace the camera rays' BVH walk.  Boxes only cull, so frames must be bit-identical
    to the oracle for cameras that see the scene from outside, from INSIDE an instance box, with a narrow and a very wide field of view,
    looking away from the scene, and for a 2 000-instance scene where most tiles overflow their list and fall back to the BVH."""
    s = castle_small
    lo, hi = s.bounds()
    c = (lo + hi) / 2
    W, H = 200, 120                                                  # 7 x 4 tiles, the last column / row partial
    cams = [host.camera_look_at(c + np.array((0.9, 0.5, 0.7)) * (hi - lo).max(), c, width=W, height=H),
            host.camera_look_at(c + np.array((0.05, 0.02, 0.03)) * (hi - lo).max(), hi, width=W, height=H),           # inside the scene
            host.camera_look_at(c + np.array((1.5, 1.0, 1.2)) * (hi - lo).max(), c, fov=np.float32(0.08), width=W, height=H),
            host.camera_look_at(c + np.array((0.4, 0.3, 0.2)) * (hi - lo).max(), c, fov=np.float32(2.9), width=W, height=H),
            host.camera_look_at(hi + 20.0, hi + 40.0, width=W, height=H)]                                              # looking away
    sky = host.sunlight_bake()
    for k, cam in enumerate(cams):
        for pipeline in ("megakernel", "wavefront"):
            g = s.loaded.scene.trace(cam, sky, W, H, spp=2, max_bounces=1, seed=k, hits=True, pipeline=pipeline)
            o = s.oscene.render(cam, sky, W, H, spp=2, max_bounces=1, seed=k)
            _assert_hits_equal(g, o)
            assert np.array_equal(g["radiance"].view(np.uint32), o["radiance"].view(np.uint32)), (k, pipeline)
