"""Sunlight::bake (crates/render/src/pipeline/sky.rs), sky/sun evaluation (headers/sky.glsl) and camera
rays (headers/camera.glsl, standard.rs:277-302): product host mirror vs oracle, plus sanity properties."""
import ctypes as C

import numpy as np
import pytest


@pytest.mark.parametrize("turb,alb,d", [
    (1.0, (0.2, 0.2, 0.2), (0.0, 0.80114365, -0.5984721)),          # Sunlight::default (sky.rs:15-22)
    (2.5, (0.0, 0.3, 1.0), (0.6, 0.64, 0.48)),
    (9.99, (0.5, 0.5, 0.5), (0.0, 0.05, 0.99875)),
    (10.0, (1.0, 0.0, 0.4), (0.0, 1.0, 0.0)),
])
def test_bake_host_equals_oracle(oracle, host, turb, alb, d):
    a, b = oracle.sunlight_bake(turb, alb, d), host.sunlight_bake(turb, alb, d)
    assert bytes(a) == bytes(b)
    v = np.frombuffer(bytes(a), np.float32)
    assert np.isfinite(v).all()
    assert v[48:51].tolist() == [np.float32(x) for x in d] and v[51] == 0.0     # direction
    assert v[55] == np.float32((np.float32(0.51) * (np.float32(np.pi) / np.float32(180.0))) / np.float32(2.0))   # solar radius (sky.rs:129)
    assert (v[52:55] > 0).all()                                                   # solar radiance XYZ


def test_castle_sun_direction_is_zenith(host):
    """examples/castle.rs:260-265 at scroll 0: R_y(0.2) * R_x(-pi/2) * (0,0,1) = straight up (to f32 rounding)"""
    d = host.castle_sun_direction(0.0)
    assert d[1] == 1.0 and abs(d[0]) < 1e-7 and abs(d[2]) < 1e-7


def test_sky_and_sun_evaluation(oracle):
    sky = oracle.sunlight_bake()
    sun = np.array([0.0, 0.80114365, -0.5984721], np.float32)
    up = oracle.sky_radiance(sky, (0, 1, 0))
    horizon = oracle.sky_radiance(sky, (1, 0, 0))
    assert (up > 0).all() and (horizon > 0).all() and np.isfinite(up).all()
    # the disc: non-zero only within the solar radius of the sun direction, brightest at its centre
    centre = oracle.sun_radiance(sky, sun)
    assert (centre > 1e3 * up).all()
    # literal restatement quirk: `singamma` is 1 - cos^2 = sin^2(gamma) and is squared again (sky.glsl:90-91),
    # so the disc ends at sin^2(gamma) = sin(radius), i.e. gamma ~ 0.067 rad, not at the 0.00445 rad radius
    off = sun + np.array([0.02, 0, 0], np.float32)
    assert (oracle.sun_radiance(sky, off / np.linalg.norm(off)) > 0).all()
    off = sun + np.array([0.1, 0, 0], np.float32)
    assert (oracle.sun_radiance(sky, off / np.linalg.norm(off)) == 0).all()
    assert (oracle.sun_radiance(sky, (0, -1, 0)) == 0).all()
    # sun below the horizon: the shader returns black sky (sky.glsl:23-26)
    night = oracle.sunlight_bake(direction=(0.0, -0.2, 0.98))
    assert (oracle.sky_radiance(night, (0, 1, 0)) == 0).all()
    # SRGBToLinear (color.glsl:1-5)
    f = oracle.lib().o_srgb_to_linear
    assert f(0.0) == 0.0 and f(1.0) == pytest.approx(1.0, abs=1e-6) and f(0.04) == pytest.approx(0.04 / 12.92)
    assert f(0.5) == pytest.approx(((0.5 + 0.055) / 1.055) ** 2.4, rel=1e-6)


def test_camera_settings_and_rays(oracle, host):
    W, H = 64, 48
    cam = host.camera_look_at((122.0, 300.61, 54.45), (0, 0, 0), width=W, height=H)      # examples/castle.rs:126-128
    assert C.sizeof(cam) == 192
    assert cam.tan_half_fov == pytest.approx(np.tan(np.pi / 8)) and cam.near == np.float32(0.1) and cam.far == 10000.0
    R = np.array([list(cam.camera_view_col0), list(cam.camera_view_col1), list(cam.camera_view_col2)]).T
    assert np.allclose(R.T @ R, np.eye(3), atol=1e-6) and np.linalg.det(R) == pytest.approx(1.0, abs=1e-6)
    fwd = -np.array([122.0, 300.61, 54.45]) / np.linalg.norm([122.0, 300.61, 54.45])
    assert np.allclose(-R[:, 2], fwd, atol=1e-6)
    vp = np.array(list(cam.view_proj)).reshape(4, 4).T
    ivp = np.array(list(cam.inverse_view_proj)).reshape(4, 4).T
    assert np.allclose(vp @ ivp, np.eye(4), atol=1e-3)
    ocam = oracle.as_struct(oracle.OCamera, cam)
    o, d = (C.c_float * 3)(), (C.c_float * 3)()
    L = oracle.lib()
    # camera.glsl:4-15 re-evaluated in numpy float32
    for px, py in ((0, 0), (W - 1, H - 1), (W // 2, H // 2), (5, 40)):
        L.o_camera_ray(C.byref(ocam), W, H, px, py, 0.5, 0.5, o, d)
        f = np.float32
        cx = f(2) * ((f(px) + f(0.5)) / f(W)) - f(1)
        cy = (f(2) * ((f(py) + f(0.5)) / f(H)) - f(1)) * f(-1)
        cx = cx * (f(W) / f(H)) * f(cam.tan_half_fov)
        cy = cy * f(cam.tan_half_fov)
        want = R.astype(np.float32) @ np.array([cx, cy, -1], np.float32)
        assert np.allclose(list(d), want, rtol=1e-6, atol=1e-7)
        assert list(o) == [np.float32(122.0), np.float32(300.61), np.float32(54.45)]
    # the centre of an even-sized image looks (almost exactly) along the forward axis; d is NOT normalised
    L.o_camera_ray(C.byref(ocam), 2, 2, 0, 0, 1.0, 1.0, o, d)
    assert np.allclose(list(d), fwd, atol=1e-6)


def test_world2obj_inverse(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(2)
    for _ in range(20):
        A = rng.normal(size=(3, 3)).astype(np.float32) + 2 * np.eye(3, dtype=np.float32)
        t = rng.normal(size=3).astype(np.float32) * 50
        m = np.concatenate([A, t[:, None]], axis=1).astype(np.float32).ravel()
        out = (C.c_float * 12)()
        L.o_world2obj(m.ctypes.data_as(C.POINTER(C.c_float)), out)
        W = np.array(out, np.float64).reshape(3, 4)
        assert np.allclose(W[:, :3] @ A.astype(np.float64), np.eye(3), atol=1e-5)
        assert np.allclose(W[:, :3] @ t + W[:, 3], 0, atol=1e-3)
