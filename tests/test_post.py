"""Post chain: auto-exposure histogram / adapted luminance / tone map.
CPU: the checker (oracle/oracle_post.c) against independent numpy float64 formulas of the shaders.
GPU: dust_auto_exposure / dust_tone_map / DUST_FMT_YCOCG16F through the C ABI against the checker."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as po


def _rand_illuminance(h, w, seed, dark_frac=0.1):
    rng = np.random.default_rng(seed)
    rgb = np.exp(rng.uniform(np.log(1e-3), np.log(300.0), (h, w, 3))).astype(np.float32)
    rgb[rng.random((h, w)) < dark_frac] = 0.0                       # black pixels -> bin 0
    rgba = np.concatenate([rgb, np.ones((h, w, 1), np.float32)], -1)
    return po.pack_illuminance(rgba)


def _rand_albedo(h, w, seed):
    rng = np.random.default_rng(seed)
    r, g, b = (rng.integers(0, 1024, (h, w), dtype=np.uint32) for _ in range(3))
    return (np.uint32(3) << 30) | (r << 20) | (g << 10) | b


def _unpack64(img16):
    c = img16.astype(np.float64)
    t = c[..., 0] - c[..., 2]
    return np.maximum(np.stack([t + c[..., 1], c[..., 0] + c[..., 2], t - c[..., 1]], -1), 0.0)


def test_half_conversions_match_numpy():
    L = po.lib()
    bits = np.arange(65536, dtype=np.uint16)
    ref = bits.view(np.float16).astype(np.float32)
    got = np.array([L.o_half_to_float(int(b)) for b in bits], np.float32)
    ok = np.isnan(ref) | (ref.view(np.uint32) == got.view(np.uint32))
    assert ok.all()
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.standard_normal(20000).astype(np.float32) * np.float32(10.0) ** rng.integers(-9, 6, 20000).astype(np.float32),
                         np.array([0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e9, 5.96e-8, 2.98e-8, 2.9802322e-8, 6.1e-5, np.inf, -np.inf], np.float32),
                         bits.view(np.float16).astype(np.float32)[~np.isnan(ref)]])
    with np.errstate(over="ignore"):
        want = xs.astype(np.float16).view(np.uint16)
    got = np.array([L.o_float_to_half(float(x)) for x in xs], np.uint16)
    assert np.array_equal(want, got)


def test_log2_is_accurate():
    L = po.lib()
    rng = np.random.default_rng(2)
    xs = np.exp(rng.uniform(np.log(0.005), np.log(6e4), 20000)).astype(np.float32)
    got = np.array([L.o_post_log2(float(x)) for x in xs], np.float64)
    assert np.abs(got - np.log2(xs.astype(np.float64))).max() < 2e-6
    for k in range(-7, 16):
        assert L.o_post_log2(float(2.0 ** k)) == float(k)


def test_histogram_and_average_against_numpy():
    h, w = 48, 80
    img = _rand_illuminance(h, w, 3)
    bins, avg = po.exposure_update(img, 0.25)
    assert bins.sum() == h * w
    rgb = _unpack64(img)
    lum = rgb @ np.array([0.299, 0.587, 0.114])
    with np.errstate(divide="ignore"):
        ref = np.where(lum < 0.005, 0, np.floor(np.clip((np.log2(lum) + 6.0) / 14.5, 0, 1) * 254.0 + 1.0)).astype(np.int64)
    ref_bins = np.bincount(ref.ravel(), minlength=256)
    # float32 vs float64 can move a pixel sitting on a bin edge by one bin, nothing else
    assert np.abs(np.cumsum(bins.astype(np.int64) - ref_bins)).max() <= 3
    assert bins[0] == ref_bins[0]
    s = int((bins.astype(np.uint64) * np.arange(256, dtype=np.uint64)).sum())
    want = 0.25 + (2.0 ** (((s / (h * w) - 1.0) / 254.0) * 14.5 - 6.0) - 0.25) * 0.2
    assert abs(avg - want) <= 2e-6 * want


def test_average_known_answers():
    p = po.OExposureParams(-6.0, 14.5, 1.0)
    hist = po.OHistogram()
    hist.histogram[128] = 100                        # every pixel in bin 128
    po.lib().o_exposure_average(C.byref(hist), 10, 10, C.byref(p))
    assert abs(hist.avg - 2.0 ** ((127.0 / 254.0) * 14.5 - 6.0)) < 1e-5
    assert not any(hist.histogram[:])                # reset, auto_exposure_avg.comp:26
    hist.histogram[0] = 100                          # all black: weighted average -1 -> below the minimum
    po.lib().o_exposure_average(C.byref(hist), 10, 10, C.byref(p))
    assert abs(hist.avg - 2.0 ** ((-1.0 / 254.0) * 14.5 - 6.0)) < 1e-6    # float32: last + (lum - last) with last = 11.4


ACESCG2SRGB = np.array([[1.7312546, -0.6040432, -0.08010775], [-0.131619, 1.1348418, -0.008679431], [-0.024568284, -0.12575036, 1.0656371]])
SRGB2ACESCG = np.array([[0.6031065, 0.32633433, 0.047995567], [0.07011794, 0.9199162, 0.012763573], [0.022178888, 0.11607823, 0.94101846]])


def _oetf64(c, f):
    srgb = lambda v: np.where(v <= 0.0031308, 12.92 * v, 1.055 * np.power(np.maximum(v, 0), 1 / 2.4) - 0.055)
    if f == 1:
        return srgb(c)
    if f == 2:
        return np.where(c <= -0.0031308, -1.055 * np.power(np.maximum(-c, 0), 1 / 2.4) + 0.055, srgb(c))
    if f == 3:
        return np.power(c / 52.37, 1 / 2.6)
    if f == 4:
        return np.where(c < 0.0030186, 12.92 * c, 1.055 * np.power(np.maximum(c, 0), 1 / 2.4) - 0.055)
    if f == 5:
        return np.where(c < 0.0181, 4.5 * c, 1.0993 * np.power(np.maximum(c, 0), 0.45) - 0.0993)
    if f == 6:
        m1, m2, c2, c3 = 2610 / 16384, 2523 / 4096 * 128, 2413 / 4096 * 32, 2392 / 4096 * 32
        Lm = np.power(c, m1)
        return np.power((c3 - c2 + 1 + c2 * Lm) / (1 + c3 * Lm), m2)
    if f == 7:
        a = 0.17883277
        with np.errstate(invalid="ignore"):                      # sqrt of a negative ACES output is NaN in the shader too
            return np.where(c < 1 / 12, np.sqrt(3 * c), a * np.log(np.maximum(12 * c - (1 - 4 * a), 1e-30)) + 0.55991073)
    if f == 8:
        return np.power(c, 256 / 563)
    return c


def tone_map_numpy(img16, albedo, avg, csc, oetf):
    """tone_map.comp:170-219 in float64, written from the shader independently of oracle_post.c."""
    L = _unpack64(img16)
    al = np.stack([(albedo >> 20) & 1023, (albedo >> 10) & 1023, albedo & 1023], -1).astype(np.float64) / 1023.0
    al = np.where(al < 0.04045, al / 12.92, np.power(np.abs(al + 0.055) / 1.055, 2.4))
    exposure = 1.0 / (9.6 * avg) * 9.6
    m = ((L @ ACESCG2SRGB.T) * al) @ SRGB2ACESCG.T * exposure
    M = np.array(csc, np.float64).reshape(3, 3).T                   # mat3(c0..c8): columns
    v = m @ M.T
    rrt = np.array([[0.59719, 0.35458, 0.04823], [0.07600, 0.90834, 0.01566], [0.02840, 0.13383, 0.83777]])
    odt = np.array([[1.60475, -0.53108, -0.07367], [-0.10208, 1.10813, -0.00605], [-0.00327, -0.07276, 1.07602]])
    r = v @ rrt.T                                                    # v * mat3(cols) : component j = dot(v, col j)
    r = (r * (r + 0.0245786) - 0.000090537) / (r * (0.983729 * r + 0.4329510) + 0.238081)
    o = r @ odt.T
    return np.concatenate([_oetf64(o, oetf), np.ones(o.shape[:-1] + (1,))], -1)


@pytest.mark.parametrize("oetf", [0, 1, 2, 4, 5, 7])
def test_tone_map_against_numpy(oetf):
    h, w = 32, 48
    img, alb = _rand_illuminance(h, w, 5, dark_frac=0.05), _rand_albedo(h, w, 6)
    csc = (0.9, 0.05, 0.0, 0.1, 0.9, 0.02, 0.0, 0.05, 0.98)
    got = po.tone_map(img, alb, 0.37, csc, oetf).astype(np.float64)
    want = tone_map_numpy(img, alb, 0.37, csc, oetf)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    got, want = np.nan_to_num(got), np.nan_to_num(want)
    # (sqrt near zero in HLG amplifies the float32 cancellation of the ACES fit: looser bound there)
    assert np.abs(got - want).max() <= (1e-4 if oetf == 7 else 2e-5) * max(1.0, np.abs(want).max())
    assert np.array_equal(got[..., 3], np.ones((h, w)))


# ------------------------------------------------------------------------------------------------ GPU

@pytest.fixture(scope="module")
def ctx():
    from dust_b200.pipeline import Context
    c = Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1080, 1920), (37, 53), (1, 1)])
def test_gpu_auto_exposure_matches_checker(ctx, shape):
    from dust_b200.pipeline import Exposure
    h, w = shape
    img = _rand_illuminance(h, w, 7)
    ex = Exposure(ctx)
    assert ex.avg == 0.0                                             # zero-filled buffer, auto_exposure.rs:116-118
    avg_o = 0.0
    for _ in range(3):                                               # temporal adaptation over three frames
        got = ex.update(img, w, h)
        _, avg_o = po.exposure_update(img, avg_o)
        assert abs(got - avg_o) <= 1e-6 * avg_o                      # equal bin sums; exp2f differs by an ulp at most
    assert abs(ex.avg - avg_o) <= 1e-6 * avg_o
    ex.close()


@pytest.mark.gpu
def test_gpu_auto_exposure_device_buffers_and_bins(ctx):
    """One pixel moved to a neighbouring bin changes the average by more than the tolerance used above:
    demonstrate it, so the avg comparison is a real check of the bins."""
    import torch
    from dust_b200.pipeline import Exposure
    h, w = 64, 64
    img = _rand_illuminance(h, w, 8, dark_frac=0.0)
    ex = Exposure(ctx, time_coefficient=1.0)
    ctx.use_torch_stream()
    d = torch.from_numpy(img.view(np.int16)).cuda()
    got = ex.update(d, w, h)
    ctx.set_stream(None)
    bins, want = po.exposure_update(img, 0.0, time_coeff=1.0)
    assert abs(got - want) <= 1e-6 * want
    s = float((bins.astype(np.float64) * np.arange(256)).sum())
    moved = 2.0 ** ((((s + 1) / (h * w) - 1.0) / 254.0) * 14.5 - 6.0)
    assert abs(moved - want) > 5e-6 * want
    ex.close()


@pytest.mark.gpu
@pytest.mark.parametrize("oetf,fmt", [(1, np.float32), (0, np.float32), (5, np.float16), (7, np.float32)])
def test_gpu_tone_map_matches_checker(ctx, oetf, fmt):
    from dust_b200.pipeline import Exposure
    h, w = 270, 481
    img, alb = _rand_illuminance(h, w, 9, dark_frac=0.05), _rand_albedo(h, w, 10)
    ex = Exposure(ctx)
    ex.avg = 0.42
    csc = (0.9, 0.05, 0.0, 0.1, 0.9, 0.02, 0.0, 0.05, 0.98)
    dst = ex.tone_map(img, alb, w, h, np.empty((h, w, 4), fmt), csc=csc, oetf=oetf)
    want = po.tone_map(img, alb, 0.42, csc, oetf)
    # float32: CUDA's powf / logf against glibc's (<= 2 ulp each); the ACES fit cancels near zero, hence the absolute term
    rel, ab = (1e-5, 1e-4) if fmt == np.float32 else (1.0 / 1024, 1e-4)   # 1e-4 of an LDR value in [0, 1]: the OETF is steep near 0
    assert np.array_equal(np.isnan(dst), np.isnan(want))
    a, b = np.nan_to_num(dst.astype(np.float64)), np.nan_to_num(want.astype(np.float64))
    excess = np.abs(a - b) - (rel * np.abs(b) + ab)
    assert excess.max() <= 0, excess.max()
    ex.close()


@pytest.mark.gpu
@pytest.mark.parametrize("pipeline", ["megakernel", "wavefront"])
def test_gpu_demodulated_ycocg_frame_feeds_the_post_chain(ctx, pipeline):
    """dust_trace(DUST_TRACE_DEMODULATE, DUST_FMT_YCOCG16F) + G-buffer albedo -> dust_auto_exposure -> dust_tone_map
    equals the checker's chain; and modulating the demodulated frame gives back the default radiance."""
    from dust_b200 import host, scenes
    from dust_b200.capi import DUST_FMT_YCOCG16F
    from dust_b200.loader import load_vox
    from dust_b200.pipeline import Exposure
    data = scenes.teapot_like()
    loaded = load_vox(ctx, data)
    W, H = 96, 64
    ext = np.array(loaded.models[0].size, np.float32)
    cam = host.camera_look_at(ext / 2 + np.array([1.2, 0.9, 1.2], np.float32) * ext.max(), ext / 2, width=W, height=H)
    sky = host.sunlight_bake()
    kw = dict(spp=2, max_bounces=1, seed=3)
    g = loaded.scene.trace(cam, sky, W, H, fmt=DUST_FMT_YCOCG16F, demodulate=True, gbuffer=True, pipeline=pipeline, **kw)
    ov = po.VoxFile(data)
    om = ov.load_model(0)
    osc = po.Scene([(ov.instances[0][1], om.tree._h, om.materials, ov.palette)])
    o = osc.render(cam, sky, W, H, demodulate=True, gbuffer=True, **kw)
    want16 = po.pack_illuminance(o["radiance"])
    a, b = g["radiance"].astype(np.float64)[..., :3], want16.astype(np.float64)[..., :3]
    scale = np.maximum(np.abs(b).max(axis=-1, keepdims=True), 1e-3 * np.abs(b).mean())
    assert (np.abs(a - b) / scale).max() <= 1e-4 + 2.0 ** -10           # 1e-4 radiance tolerance + one half ulp
    assert np.array_equal(g["gbuffer"]["albedo"], o["gbuffer"]["albedo"])
    ex = Exposure(ctx)
    avg = ex.update(g["radiance"], W, H)
    _, avg_o = po.exposure_update(want16, 0.0)
    assert abs(avg - avg_o) <= 2e-3 * avg_o                              # a 1e-4 radiance difference may move a pixel across a bin edge
    ldr = ex.tone_map(g["radiance"], g["gbuffer"]["albedo"], W, H, np.empty((H, W, 4), np.float32))
    ldr_o = po.tone_map(want16, o["gbuffer"]["albedo"], avg_o)
    assert np.abs(ldr - ldr_o).max() <= 4e-3
    ex.close()
