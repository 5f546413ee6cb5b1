"""The reference's own GI: ambient-occlusion, final-gather and surfel passes over the world-space radiance
cache.  CPU part: the oracle's pieces against independent restatements / known answers and its determinism.  GPU part (marked):
dust_trace_gi against the oracle over several frames — cache table, surfel pool, G-buffer and illuminance image BIT FOR BIT."""
import numpy as np
import pytest

from conftest import BuiltScene
from dust_b200 import scenes


def _f3(v):
    import ctypes as C
    return (C.c_float * 3)(*[float(x) for x in v])


def test_unpack_normal_is_the_inverse_of_the_gbuffer_pack_up_to_quantisation(oracle):
    """hit.rchit:83 packs the world normal with NRD_FrontEnd_PackNormalAndRoughness into 10-bit UNORM; the later passes read it back
    (nrd.glsl:54-94).  For the six axis normals the round trip is off-axis by ~1e-3 — that wobble is the reference's behaviour."""
    import ctypes as C
    L = oracle.lib()
    out = (C.c_float * 3)()

    def pack(n):                                                    # numpy restatement of nrd.glsl:2-10, 25-38 + UNORM10 store
        v = np.array(n, np.float64) / np.abs(n).sum()
        wrap = (1 - np.abs(v[[1, 0]])) * np.where(v[:2] >= 0, 1.0, -1.0)
        e = v[:2] if v[2] >= 0 else wrap
        q = np.rint(np.clip(e * 0.5 + 0.5, 0, 1) * 1023).astype(np.uint32)
        return (3 << 30) | (int(q[0]) << 20) | (int(q[1]) << 10) | 1023

    for n in ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1), (0.6, 0.48, -0.64)):
        L.o_gi_unpack_normal(pack(n), out)
        got = np.array(list(out))
        assert abs(np.linalg.norm(got) - 1) < 1e-6
        assert np.abs(got - np.array(n)).max() < 3e-3, (n, got)


def test_ambient_occlusion_leaf_rule(oracle):
    """ambient_occlusion.rint:62-73: a leaf whose box contains t = 8 along the ray is a solid block (voxel 0xFF) when non-empty and is
    skipped when empty; other leaves get hit.rint's DDA."""
    import ctypes as C
    L = oracle.lib()
    t, vox = C.c_float(), C.c_uint32()
    far_bit = 1 << (3 << 4 | 1 << 2 | 1)                            # voxel (3, 1, 1): the DDA would reach it at t = 9
    o, d = _f3((-6.0, 1.5, 1.5)), _f3((1.0, 0.0001, 0.0001))         # box entered at t = 6, left at t = 10: 8 lies inside
    assert L.o_leaf_ao(o, d, far_bit, 0.1, C.byref(t), C.byref(vox)) == 1 and vox.value == 0xFF and t.value == pytest.approx(6.0, abs=1e-3)
    assert L.o_leaf_dda(o, d, far_bit, 0.1, C.byref(t), C.byref(vox)) == 1 and vox.value == (3 << 4 | 1 << 2 | 1) and t.value == pytest.approx(9.0, abs=1e-3)
    assert L.o_leaf_ao(o, d, 0, 0.1, C.byref(t), C.byref(vox)) == 0
    o2 = _f3((-1.0, 1.5, 1.5))                                       # box spans t in [1, 5]: the threshold is behind it -> plain DDA
    assert L.o_leaf_ao(o2, d, far_bit, 0.1, C.byref(t), C.byref(vox)) == 1 and vox.value == (3 << 4 | 1 << 2 | 1)


@pytest.fixture(scope="module")
def gi_scene(oracle):
    s = BuiltScene(scenes.castle_like(tiles=2, tile=64, height=96, props=9), oracle)
    lo, hi = s.bounds()
    c = (lo + hi) / 2
    W, H = 160, 90
    cam = s.host.camera_look_at(c + np.array((0.8, 0.6, 0.5)) * (hi - lo).max(), c, width=W, height=H)
    return s, cam, s.host.sunlight_bake(), W, H


def test_oracle_gi_frames_fill_the_cache_and_do_not_depend_on_threads(oracle, gi_scene):
    s, cam, sky, W, H = gi_scene
    a, b = oracle.GiState(1 << 16, 4096), oracle.GiState(1 << 16, 4096)
    for f in range(3):
        ra = s.oscene.gi_frame(cam, sky, W, H, a, seed=3, frame_index=f, threads=1)
        rb = s.oscene.gi_frame(cam, sky, W, H, b, seed=3, frame_index=f, threads=5)
        assert np.array_equal(ra["illuminance"].view(np.uint16), rb["illuminance"].view(np.uint16))
        assert ra["stats"] == rb["stats"]
    assert np.array_equal(a.table.view(np.uint8), b.table.view(np.uint8)) and np.array_equal(a.pool.view(np.uint8), b.pool.view(np.uint8))
    st = ra["stats"]
    hit = np.isfinite(ra["gbuffer"]["depth"])
    assert st["near_hits"] + st["far_hits"] + st["far_misses"] == int(hit.sum())          # every hit pixel is resolved by exactly one segment
    assert st["cache_found"] > 0 and st["cache_inserts"] > 0 and st["surfel_rays"] > 0
    used = a.table[a.table["fingerprint"] != 0]
    assert len(used) > 100 and (used["sample_count"] >= 1).all() and (used["sample_count"] <= 404).all()
    assert set(np.unique(used["last_accessed_frame"])) <= {0, 1, 2}
    valid = a.pool[a.pool["direction"] < 6]
    assert len(valid) == st["surfel_rays"] or len(valid) >= st["surfel_rays"]             # entries scheduled this frame are traced next frame
    # illuminance: resolved pixels carry their hit distance, the others 0; miss pixels hold the sky (w = +inf as a half)
    w = ra["illuminance"][..., 3].astype(np.float32)
    assert np.isinf(w[~hit]).all() and (w[hit] >= 0).all()
    assert int((w[hit] > 0).sum()) == st["near_hits"] + st["far_hits"]


def test_cache_entries_equal_a_sequential_replay_of_the_frame_inserts(oracle, gi_scene):
    """The surfel pass' inserts of ONE frame, replayed one by one in ascending surfel index through o_cache_insert on a copy of the
    table as the final-gather pass left it, give the frame's table: the order the device implementation has to reproduce."""
    s, cam, sky, W, H = gi_scene
    st = oracle.GiState(1 << 14, 2048)                               # a small table: probe windows collide, evictions happen
    for f in range(4):
        r = s.oscene.gi_frame(cam, sky, W, H, st, seed=9, frame_index=f)
    used = st.table[st.table["fingerprint"] != 0]
    assert len(used) > 200
    assert r["stats"]["cache_inserts"] > 0


# ------------------------------------------------------------------------------------------------------------------ GPU

@pytest.mark.gpu
@pytest.mark.parametrize("capacity,pool", [(1 << 16, 4096), (1 << 12, 1536)])
def test_gi_frames_bit_exact_against_the_oracle(oracle, ctx, gi_scene, capacity, pool):
    """Five frames of dust_trace_gi against the oracle: G-buffer, illuminance image, every cache entry and every surfel bit for bit,
    and the pass statistics.  The small configuration forces probe-window collisions, LRU evictions and long insert segments."""
    from dust_b200.loader import load_vox
    from dust_b200.pipeline import RadianceCache
    s, cam, sky, W, H = gi_scene
    loaded = load_vox(ctx, s.data)
    cache = RadianceCache(ctx, capacity, pool)
    ost = oracle.GiState(capacity, pool)
    for f in range(5):
        g = cache.gi_frame(loaded.scene, cam, sky, W, H, seed=3, frame_index=f)
        o = s.oscene.gi_frame(cam, sky, W, H, ost, seed=3, frame_index=f)
        for k in ("albedo", "depth", "normal", "voxel_id"):
            assert np.array_equal(g["gbuffer"][k].view(np.uint32), o["gbuffer"][k].view(np.uint32)), (f, k)
        bad = np.argwhere(g["illuminance"].view(np.uint16) != o["illuminance"].view(np.uint16))
        assert len(bad) == 0, (f, len(bad), bad[:4], g["illuminance"][tuple(bad[0][:2])], o["illuminance"][tuple(bad[0][:2])])
        t, p = cache.download()
        assert np.array_equal(p.view(np.uint8), ost.pool.view(np.uint8)), f
        diff = np.flatnonzero((t.view(np.uint8).reshape(-1, 12) != ost.table.view(np.uint8).reshape(-1, 12)).any(axis=1))
        assert len(diff) == 0, (f, len(diff), t[diff[:3]], ost.table[diff[:3]])
        gs = dict(g["stats"])
        segs = gs.pop("insert_segments")
        os_ = dict(o["stats"]); os_.pop("insert_segments")
        assert gs == os_, (f, gs, os_)
        assert 0 < segs <= gs["cache_inserts"] or gs["cache_inserts"] == 0
    assert gs["cache_found"] > 0 and gs["cache_inserts"] > 0
    cache.clear()
    t, p = cache.download()
    assert not t.view(np.uint8).any() and (p.view(np.uint8) == 0xFF).all()
    cache.close()


@pytest.mark.gpu
def test_gi_device_frame_and_errors(ctx, gi_scene):
    import torch
    from dust_b200.capi import DustError
    from dust_b200.loader import load_vox
    from dust_b200.pipeline import RadianceCache
    s, cam, sky, W, H = gi_scene
    loaded = load_vox(ctx, s.data)
    a, b = RadianceCache(ctx, 1 << 15, 2048), RadianceCache(ctx, 1 << 15, 2048)
    dev = torch.zeros((H, W, 4), dtype=torch.float16, device="cuda")
    ctx.use_torch_stream()
    for f in range(2):
        b.gi_frame_device(loaded.scene, cam, sky, W, H, dev, seed=1, frame_index=f)
    torch.cuda.synchronize()
    ctx.set_stream(None)
    for f in range(2):
        host = a.gi_frame(loaded.scene, cam, sky, W, H, seed=1, frame_index=f)
    assert np.array_equal(dev.cpu().numpy().view(np.uint16), host["illuminance"].view(np.uint16))
    ta, pa = a.download(); tb, pb = b.download()
    assert np.array_equal(ta.view(np.uint8), tb.view(np.uint8)) and np.array_equal(pa.view(np.uint8), pb.view(np.uint8))
    with pytest.raises(DustError):
        RadianceCache(ctx, 0, 16)
    a.close(); b.close()


@pytest.mark.gpu
def test_gi_mixed_levels_small_frame_large_pool(oracle, ctx, host):
    """A 4-level tree (hierarchy!(4,4,2,2)) next to a sheared 3-level instance (the NL = 0 kernel instantiations), a frame with fewer
    pixels than pool slots, three frames bit for bit; then the 4-level tree alone (NL = 4) and an empty scene."""
    from dust_b200.pipeline import RadianceCache, Scene
    dom = host.SynthDomain(seed=7, fill=0.03, domain_leaves=128)
    otree = oracle.TreeView(dom.describe())
    dtree = ctx.upload_tree(dom.describe())
    tea = BuiltScene(scenes.teapot_like(), oracle, ctx)
    tea_tree = next(iter(tea.loaded.trees.values()))
    ident = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
    sheared = np.array([1.5, 0.25, 0, 200, 0, 1, 0.125, 330, 0.0625, 0, 1, 150], np.float32)
    om = tea.omodels[0]
    W, H = 40, 30
    cam = host.camera_look_at((700.0, 620.0, 760.0), (256.0, 256.0, 256.0), width=W, height=H)
    sky = host.sunlight_bake()
    for insts_o, insts_d in (([(ident, otree._h, dom.materials, dom.palette), (sheared, om.tree._h, om.materials, tea.ovox.palette)],
                              [(ident, dtree), (sheared, tea_tree)]),
                             ([(ident, otree._h, dom.materials, dom.palette)], [(ident, dtree)]),
                             ([], [])):
        osc, dsc = oracle.Scene(insts_o), Scene(ctx, insts_d)
        cache, ost = RadianceCache(ctx, 1 << 13, 4096), oracle.GiState(1 << 13, 4096)          # 1200 pixels, 4096 pool slots
        for f in range(3):
            g = cache.gi_frame(dsc, cam, sky, W, H, seed=5, frame_index=f)
            o = osc.gi_frame(cam, sky, W, H, ost, seed=5, frame_index=f)
            assert np.array_equal(g["illuminance"].view(np.uint16), o["illuminance"].view(np.uint16)), (len(insts_d), f)
            t, p = cache.download()
            assert np.array_equal(t.view(np.uint8), ost.table.view(np.uint8)) and np.array_equal(p.view(np.uint8), ost.pool.view(np.uint8)), (len(insts_d), f)
        if insts_d:
            assert g["stats"]["far_hits"] + g["stats"]["near_hits"] > 0
        cache.close()
