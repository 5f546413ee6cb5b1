"""The instance BVH's 16-bit boxes (device_layout.h "WIDE BVH") must never lose a candidate: whatever ray touches a padded instance
box must touch its quantised box, and the entry time computed from the quantised box must not lie behind the real one (the tracer
stops at the first candidate whose entry lies behind its best hit).

This file restates the two halves of that arithmetic — the host's grid + outward quantisation (dust_abi.cu bvh_grid / build_wide) in
float64 and the device's decode (trace_core.cuh cull_inverse / cull_ray / cull_plane) in float32 — and checks the invariant on random
scenes, including the ones that stress it: coordinates 65 000 units from the origin (the grid step is then bounded by the floats' own
spacing), sub-unit boxes, ray origins 40 scene sizes away, exactly axis-parallel rays.  The GPU half of the evidence is
tests/test_gpu_parity.py::test_wide_bvh_far_from_origin_and_mixed_scales (every hit bit-identical to the oracle's linear instance loop)."""
import math

import numpy as np

f32 = np.float32


def bvh_grid(root_lo, root_hi):
    """dust_abi.cu bvh_grid: per axis a power-of-two step (<= 65 000 positions over the scene box, >= 4 ulp of its coordinates)."""
    org, step = [], []
    for a in range(3):
        maxabs = max(abs(float(root_lo[a])), abs(float(root_hi[a])), 1e-30)
        pad = 4e-7 * maxabs
        lo, hi = float(root_lo[a]) - pad, float(root_hi[a]) + pad
        e = math.ceil(math.log2(max(hi - lo, 1e-300) / 65000.0))
        e = min(max(e, math.frexp(maxabs)[1] - 1 - 21, -100), 100)          # ilogb(x) = frexp exponent - 1
        s = math.ldexp(1.0, e)
        o = f32(lo - 2.0 * s)
        while float(o) > lo - s:
            o = np.nextafter(o, f32(-np.inf))
        org.append(o)
        step.append(s)
    return np.array(org, f32), np.array(step, np.float64)


def quantise(box_lo, box_hi, org, step):
    """dust_abi.cu build_wide: outwards, widened by one step; returns (qlo, qhi) per axis."""
    q = []
    for a in range(3):
        maxabs = max(abs(float(box_lo[a])), abs(float(box_hi[a])))
        pad = 4e-7 * maxabs
        ql = math.floor((float(box_lo[a]) - pad - float(org[a])) / step[a]) - 1.0
        qh = math.ceil((float(box_hi[a]) + pad - float(org[a])) / step[a]) + 1.0
        assert 0.0 <= ql and qh <= 65535.0, "box does not fit the grid"
        q.append((int(ql), int(qh)))
    return q


def fma32(a, b, c):
    """float32 fma (the product of two f32 is exact in f64; the sum is then rounded once more — a double rounding that can only
    move the result by an amount far below what the test tolerates)"""
    return f32(np.float64(a) * np.float64(b) + np.float64(c))


def cull_inverse(d):
    """trace_core.cuh cull_inverse"""
    return np.array([f32(math.copysign(1e24, float(x))) if abs(float(x)) < 1e-24 else f32(1.0) / f32(x) for x in d], f32)


def device_interval(q, org, step, o, winv, tmin, tmax):
    """trace_core.cuh cull_ray + cull_plane + the slab test of cull_expand, in float32"""
    t0, t1 = f32(tmin), f32(tmax)
    for a in range(3):
        s = f32(step[a]) * winv[a]
        A = fma32(f32(-8388608.0), s, (org[a] - o[a]) * winv[a])
        near, far = (q[a][0], q[a][1]) if winv[a] > 0 else (q[a][1], q[a][0])
        tn = fma32(f32(8388608.0 + near), s, A)
        tf = fma32(f32(8388608.0 + far), s, A)
        t0, t1 = max(t0, tn), min(t1, tf)
    return t0, t1


def exact_interval(lo, hi, o, d, tmin, tmax):
    """the ray against the (padded) float box in float64; axis-parallel components: inside the slab or not"""
    t0, t1 = float(tmin), float(tmax)
    for a in range(3):
        oa, da = float(o[a]), float(d[a])
        if abs(da) < 1e-24:
            if not (float(lo[a]) <= oa <= float(hi[a])):
                return 1.0, 0.0
            continue
        ta, tb = (float(lo[a]) - oa) / da, (float(hi[a]) - oa) / da
        t0, t1 = max(t0, min(ta, tb)), min(t1, max(ta, tb))
    return t0, t1


def _scene(rng, centre, extent, n, smallest):
    los = centre + (rng.random((n, 3)) - 0.5) * extent
    sizes = smallest + rng.random((n, 3)) ** 3 * extent * 0.2
    lo, hi = los.astype(f32), (los + sizes).astype(f32)
    # the pad of dust_abi.cu scene_arrays (0.02 + 1e-5 * the box's largest coordinate): the boxes the BVH is built over
    big = np.maximum(np.abs(lo), np.abs(hi)).max(axis=1, keepdims=True)
    pad = (f32(0.02) + f32(1e-5) * big).astype(f32)
    return lo - pad, hi + pad


def _check(rng, centre, extent, smallest, n_boxes=40, n_rays=300):
    lo, hi = _scene(rng, np.asarray(centre, np.float64), extent, n_boxes, smallest)
    org, step = bvh_grid(lo.min(axis=0), hi.max(axis=0))
    qs = [quantise(lo[k], hi[k], org, step) for k in range(n_boxes)]
    for r in range(n_rays):
        kind = r % 4
        o = (np.asarray(centre) + (rng.random(3) - 0.5) * extent * (80.0 if kind == 1 else 1.2)).astype(f32)
        if kind == 2:                                            # exactly axis-parallel, through some box
            k = rng.integers(n_boxes)
            o = ((lo[k] + hi[k]) / 2 + (rng.random(3) - 0.5) * (hi[k] - lo[k]) * 1.2).astype(f32)
            d = np.zeros(3, f32)
            ax = rng.integers(3)
            d[ax] = rng.choice([-1.0, 1.0])
            o[ax] -= f32(extent) * d[ax]
        else:
            target = lo[rng.integers(n_boxes)] + rng.random(3) * 2.0     # aim near a box corner: grazing hits
            d = (target - o).astype(np.float64)
            d = (d / np.linalg.norm(d)).astype(f32)
            if kind == 3:
                d[rng.integers(3)] = f32(1e-30)                  # below cull_inverse's threshold
        winv = cull_inverse(d)
        tmin, tmax = f32(0.1), f32(1e30)
        for k in range(n_boxes):
            e0, e1 = exact_interval(lo[k], hi[k], o, d, tmin, tmax)
            if e0 > e1:
                continue
            t0, t1 = device_interval(qs[k], org, step, o, winv, tmin, tmax)
            assert t0 <= t1, f"candidate lost: box {k}, ray kind {kind}, exact [{e0}, {e1}], quantised [{t0}, {t1}]"
            # a lower bound of the real entry (up to float32 rounding of the real entry itself)
            assert float(t0) <= e0 + 4e-7 * abs(e0) + 1e-6, f"entry time behind the real one: {t0} > {e0}"


def test_grid_is_conservative_around_the_origin():
    _check(np.random.default_rng(1), (0.0, 0.0, 0.0), 2000.0, 4.0)


def test_grid_is_conservative_far_from_the_origin():
    _check(np.random.default_rng(2), (65000.0, -30000.0, 12000.0), 6000.0, 2.0)


def test_grid_is_conservative_for_tiny_scenes_and_huge_scenes():
    _check(np.random.default_rng(3), (3.0, 1.0, -2.0), 1.5, 0.05, n_boxes=12)
    _check(np.random.default_rng(4), (1.0e6, 2.0e5, -4.0e5), 3.0e5, 100.0, n_boxes=20)


def test_grid_steps_are_powers_of_two_and_boxes_fit():
    rng = np.random.default_rng(5)
    for centre, extent in (((0, 0, 0), 10.0), ((1e5, 1e5, 1e5), 50.0), ((-7e3, 2.0, 9e4), 3e4)):
        lo, hi = _scene(rng, np.asarray(centre, np.float64), extent, 30, 0.5)
        org, step = bvh_grid(lo.min(axis=0), hi.max(axis=0))
        for a in range(3):
            m, _ = math.frexp(step[a])
            assert m == 0.5                                      # a power of two: step * inv is an exact scaling on the device
            # the float origin's rounding stays below the step it anchors
            assert step[a] >= 4 * float(np.spacing(f32(max(abs(lo[:, a]).max(), abs(hi[:, a]).max()))))
        for k in range(30):
            quantise(lo[k], hi[k], org, step)                    # asserts 0 <= q <= 65535
