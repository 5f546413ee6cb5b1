"""The sun grid (kernels.cuh wf_sun_grid / sun_list_of, dust_abi.cu dust_trace): a sun shadow ray takes its candidate instances from the
cell its ORIGIN projects into along the sun direction.  Invariant: every padded instance box a ray along the sun direction touches is in
the list of the ray's cell (or the ray falls outside the grid and walks the BVH).

The host half (basis, grid bounds, margin: float64) and the device half (projection of boxes and origins: float32) are restated here and
checked on random scenes, suns and origins, including origins ON box faces and cell borders.  GPU evidence: every frame test of
tests/test_gpu_parity.py runs through the grid and stays bit-identical to the oracle's linear instance loop."""
import numpy as np

f32 = np.float32
G = 128                                                              # DUST_SUN_GRID


def sun_basis_and_grid(sun_dir, root_lo, root_hi):
    """dust_abi.cu dust_trace, "sun shadow rays": returns the FrameParams members (u, v, u0, v0, du, dv, inv_du, inv_dv, eps) as float32"""
    sx, sy, sz = (float(x) for x in sun_dir)
    ax = [1.0, 0.0, 0.0]
    if abs(sx) > abs(sy) or abs(sx) > abs(sz):
        ax = [0.0, 1.0, 0.0] if abs(sy) <= abs(sz) else [0.0, 0.0, 1.0]
    u = np.array([sy * ax[2] - sz * ax[1], sz * ax[0] - sx * ax[2], sx * ax[1] - sy * ax[0]])
    u /= np.linalg.norm(u)
    v = np.array([sy * u[2] - sz * u[1], sz * u[0] - sx * u[2], sx * u[1] - sy * u[0]])
    corners = np.array([[(root_hi if c & 1 else root_lo)[0], (root_hi if c & 2 else root_lo)[1], (root_hi if c & 4 else root_lo)[2]] for c in range(8)], np.float64)
    pu, pv = corners @ u, corners @ v
    umin, umax, vmin, vmax = pu.min(), pu.max(), pv.min(), pv.max()
    big = np.abs(corners).max()
    eu, ev = 1e-3 * (umax - umin) + 1e-6, 1e-3 * (vmax - vmin) + 1e-6
    umin, umax, vmin, vmax = umin - eu, umax + eu, vmin - ev, vmax + ev
    du, dv = (umax - umin) / G, (vmax - vmin) / G
    eps = 1e-3 * max(du, dv) + 2e-5 * (big + (umax - umin) + (vmax - vmin))
    return dict(u=u.astype(f32), v=v.astype(f32), u0=f32(umin), v0=f32(vmin), du=f32(du), dv=f32(dv), inv_du=f32(1.0 / du), inv_dv=f32(1.0 / dv), eps=f32(eps))


def dot32(a, b):
    return (f32(a[0]) * f32(b[0]) + f32(a[1]) * f32(b[1])) + f32(a[2]) * f32(b[2])


def cell_of(g, o):
    """kernels.cuh sun_list_of: (cu, cv) or None (outside: the ray walks the BVH)"""
    fu = (dot32(o, g["u"]) - g["u0"]) * g["inv_du"]
    fv = (dot32(o, g["v"]) - g["v0"]) * g["inv_dv"]
    if not (fu >= 0 and fu < G and fv >= 0 and fv < G):
        return None
    return int(fu), int(fv)


def box_in_cell(g, lo, hi, cu, cv):
    """kernels.cuh wf_sun_grid, float32"""
    ulo, uhi = g["u0"] + f32(cu) * g["du"] - g["eps"], g["u0"] + f32(cu + 1) * g["du"] + g["eps"]
    vlo, vhi = g["v0"] + f32(cv) * g["dv"] - g["eps"], g["v0"] + f32(cv + 1) * g["dv"] + g["eps"]
    c = [f32(0.5) * (lo[k] + hi[k]) for k in range(3)]
    h = [f32(0.5) * (hi[k] - lo[k]) for k in range(3)]
    pu, ru = dot32(c, g["u"]), dot32(h, np.abs(g["u"]))
    pv, rv = dot32(c, g["v"]), dot32(h, np.abs(g["v"]))
    return pu - ru <= uhi and pu + ru >= ulo and pv - rv <= vhi and pv + rv >= vlo


def ray_touches(lo, hi, o, s):
    """a ray from o along s against the box, float64 (t >= 0)"""
    t0, t1 = 0.0, np.inf
    for a in range(3):
        if abs(s[a]) < 1e-300:
            if not (lo[a] <= o[a] <= hi[a]):
                return False
            continue
        ta, tb = (lo[a] - o[a]) / s[a], (hi[a] - o[a]) / s[a]
        t0, t1 = max(t0, min(ta, tb)), min(t1, max(ta, tb))
    return t0 <= t1


def _run(seed, centre, extent, sun):
    rng = np.random.default_rng(seed)
    n = 60
    los = np.asarray(centre) + (rng.random((n, 3)) - 0.5) * extent
    lo = los.astype(f32)
    hi = (los + 1.0 + rng.random((n, 3)) ** 2 * extent * 0.15).astype(f32)
    s = np.asarray(sun, np.float64)
    s = (s / np.linalg.norm(s)).astype(f32)                          # the kernels' sun_dir is float32
    g = sun_basis_and_grid(s, lo.min(axis=0).astype(np.float64), hi.max(axis=0).astype(np.float64))
    checked = 0
    for r in range(1500):
        k = rng.integers(n)
        kind = r % 4
        if kind == 3:
            # the worst case for the margin: the origin is the box vertex that sticks out farthest along +-u (or +-v), and the box is
            # moved so that this vertex projects exactly onto a cell border — the ray touches the box at t = 0, and its cell touches
            # the box's projected rectangle in one point
            axis, sign = (g["u"], g["v"])[rng.integers(2)].astype(np.float64), rng.choice([-1.0, 1.0])
            vert = np.where(sign * axis > 0, hi[k], lo[k]).astype(np.float64)
            pa = float(vert @ axis)
            a0, da = (float(g["u0"]), float(g["du"])) if axis is not None and np.allclose(axis, g["u"]) else (float(g["v0"]), float(g["dv"]))
            border = a0 + round((pa - a0) / da) * da
            shift = (axis * (border - pa)).astype(f32)
            lo[k] = lo[k] + shift; hi[k] = hi[k] + shift
            o = np.where(sign * axis > 0, hi[k], lo[k]).astype(np.float64)
        elif kind == 0:                                              # anywhere in the scene
            o = np.asarray(centre) + (rng.random(3) - 0.5) * extent * 1.1
        elif kind == 1:                                              # below a box, offset so the ray grazes its side faces
            p = lo[k] + rng.random(3) * (hi[k] - lo[k])
            edge = rng.integers(3)
            p[edge] = (lo[k] if rng.random() < 0.5 else hi[k])[edge]
            o = p - s.astype(np.float64) * rng.random() * extent * 0.5
        else:                                                        # on a cell border of the grid
            cu, cv = rng.integers(G), rng.integers(G)
            pu, pv = float(g["u0"]) + cu * float(g["du"]), float(g["v0"]) + cv * float(g["dv"])
            o = pu * g["u"].astype(np.float64) + pv * g["v"].astype(np.float64) + s.astype(np.float64) * (rng.random() - 0.5) * extent
        o = o.astype(f32)
        cell = cell_of(g, o)
        if cell is None:
            continue
        for j in range(n):
            if ray_touches(lo[j].astype(np.float64), hi[j].astype(np.float64), o.astype(np.float64), s.astype(np.float64)):
                assert box_in_cell(g, lo[j], hi[j], *cell), f"box {j} touched by a sun ray from {o} is missing from cell {cell}"
                checked += 1
    assert checked > 300


def test_sun_grid_lists_every_box_a_sun_ray_touches():
    _run(1, (0.0, 0.0, 0.0), 2000.0, (0.1987, 0.9801, 0.0))          # the bench's sun (castle.rs:260-265)
    _run(2, (0.0, 0.0, 0.0), 2000.0, (0.0, 1.0, 0.0))                # straight up: two zero components
    _run(3, (500.0, 100.0, -300.0), 800.0, (-0.6, 0.3, 0.74))


def test_sun_grid_far_from_the_origin_and_flat_suns():
    _run(4, (65000.0, -30000.0, 12000.0), 6000.0, (0.3, 0.05, -0.95))
    _run(5, (0.0, 0.0, 0.0), 50.0, (1.0, 1e-4, 0.0))                 # sun on the horizon
