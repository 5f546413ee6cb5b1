"""Procedural stand-in scenes and a MagicaVoxel `.vox` writer.

The reference's assets/castle.vox and assets/teapot.vox are Git-LFS pointers (88 233 039 B,
sha256 cbc09a7c...61a2 and 143 083 B, sha256 68cdab79...52ff) and there is no network, so every
run uses these seeded stand-ins of matching scale, written as REAL .vox files and ingested through
the same loader path as a genuine asset.  `load_asset()` picks the genuine file when one with the
right SHA-256 is supplied.  Every report must say "stand-in" (bench.py does).
"""
from __future__ import annotations

import hashlib
import os
import struct

import numpy as np

CASTLE_SHA256 = "cbc09a7c84fc44d5f669c0c0801715fb281c2c2c09ae78853c9ee1ac7a4161a2"
TEAPOT_SHA256 = "68cdab795e00401866ae46584d31608da2a3137065bfdfc21ab8fdd9d67b52ff"

# examples/castle.rs:126-128, crates/render/src/projection.rs:21-28
CASTLE_CAMERA = dict(eye=(122.0, 300.61, 54.45), target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0),
                     fov=np.float32(np.pi / 4), near=0.1, far=10000.0)


# ------------------------------------------------------------------------------------------ writer

def _chunk(cid: bytes, content: bytes) -> bytes:
    return cid + struct.pack("<ii", len(content), 0) + content


def _dict(d: dict) -> bytes:
    out = struct.pack("<i", len(d))
    for k, v in d.items():
        kb, vb = k.encode(), str(v).encode()
        out += struct.pack("<i", len(kb)) + kb + struct.pack("<i", len(vb)) + vb
    return out


def rotation_byte(perm, signs) -> int:
    """`_r` byte: row r of the matrix has its non-zero in column perm[r] with sign signs[r]."""
    b = perm[0] | (perm[1] << 2)
    for r in range(3):
        if signs[r] < 0:
            b |= 1 << (4 + r)
    return b


def all_rotation_bytes(proper_only=True):
    out = []
    for p in ((0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)):
        parity = 1 if p in ((0, 1, 2), (1, 2, 0), (2, 0, 1)) else -1
        for s in range(8):
            sg = (1 - 2 * (s & 1), 1 - 2 * ((s >> 1) & 1), 1 - 2 * ((s >> 2) & 1))
            if not proper_only or parity * sg[0] * sg[1] * sg[2] == 1:
                out.append(rotation_byte(p, sg))
    return out


def write_vox(models, instances, palette: np.ndarray) -> bytes:
    """models: list of (size(sx,sy,sz), xyzi uint8[n,4] with i = 0-based palette index);
    instances: list of (model_id, translation(ix,iy,iz), rotation_byte) or None for a file
    without a scene graph (single model); palette: uint8[256,4] (entry 255 unused)."""
    body = b""
    for size, xyzi in models:
        xyzi = np.ascontiguousarray(xyzi, dtype=np.uint8).reshape(-1, 4).copy()
        xyzi[:, 3] += 1                                     # stored colour index is 1-based
        body += _chunk(b"SIZE", struct.pack("<iii", *[int(s) for s in size]))
        body += _chunk(b"XYZI", struct.pack("<i", len(xyzi)) + xyzi.tobytes())
    if instances is not None:
        n = len(instances)
        body += _chunk(b"nTRN", struct.pack("<i", 0) + _dict({}) + struct.pack("<iiii", 1, -1, -1, 1) + _dict({}))
        body += _chunk(b"nGRP", struct.pack("<i", 1) + _dict({}) + struct.pack("<i", n) +
                       b"".join(struct.pack("<i", 2 + 2 * k) for k in range(n)))
        for k, (mid, t, r) in enumerate(instances):
            frame = {"_t": f"{int(t[0])} {int(t[1])} {int(t[2])}"}
            if r != 4:
                frame["_r"] = int(r)
            body += _chunk(b"nTRN", struct.pack("<i", 2 + 2 * k) + _dict({}) +
                           struct.pack("<iiii", 3 + 2 * k, -1, 0, 1) + _dict(frame))
            body += _chunk(b"nSHP", struct.pack("<i", 3 + 2 * k) + _dict({}) + struct.pack("<ii", 1, int(mid)) + _dict({}))
    body += _chunk(b"RGBA", np.ascontiguousarray(palette, dtype=np.uint8).reshape(256, 4).tobytes())
    return b"VOX " + struct.pack("<i", 200) + b"MAIN" + struct.pack("<ii", 0, len(body)) + body


# ----------------------------------------------------------------------------------------- palette

def default_palette(seed=0) -> np.ndarray:
    rng = np.random.default_rng(seed + 77)
    pal = np.zeros((256, 4), np.uint8)
    pal[:, 3] = 255

    def ramp(lo, hi, a, b):
        n = hi - lo
        t = np.linspace(0, 1, n)[:, None]
        pal[lo:hi, :3] = np.clip((1 - t) * np.array(a) + t * np.array(b) + rng.integers(-6, 7, (n, 3)), 0, 255)
    ramp(0, 32, (54, 110, 40), (120, 170, 70))        # grass
    ramp(32, 64, (92, 64, 38), (140, 105, 70))        # dirt
    ramp(64, 112, (96, 96, 100), (190, 188, 180))     # stone
    ramp(112, 144, (150, 60, 45), (205, 95, 70))      # roof
    ramp(144, 176, (180, 160, 120), (235, 220, 180))  # plaster / sand
    ramp(176, 208, (60, 40, 25), (120, 85, 50))       # wood
    ramp(208, 255, (30, 30, 30), (250, 250, 250))     # greys / ceramic
    return pal


def _hash01(ix, iy, seed):
    h = (ix.astype(np.uint32) * np.uint32(374761393) + iy.astype(np.uint32) * np.uint32(668265263)
         + np.uint32((int(seed) * 2246822519) & 0xFFFFFFFF))
    h = (h ^ (h >> np.uint32(13))) * np.uint32(1274126177)
    h = h ^ (h >> np.uint32(16))
    return (h & np.uint32(0xFFFFFF)).astype(np.float32) / np.float32(0x1000000)


def _value_noise(x, y, cell, seed):
    gx, gy = np.floor_divide(x, cell), np.floor_divide(y, cell)
    fx, fy = (x - gx * cell) / cell, (y - gy * cell) / cell
    fx, fy = fx * fx * (3 - 2 * fx), fy * fy * (3 - 2 * fy)
    a, b = _hash01(gx, gy, seed), _hash01(gx + 1, gy, seed)
    c, d = _hash01(gx, gy + 1, seed), _hash01(gx + 1, gy + 1, seed)
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def _xyzi_from(solid: np.ndarray, color: np.ndarray) -> np.ndarray:
    x, y, z = np.nonzero(solid)
    return np.stack([x, y, z, color[x, y, z]], axis=1).astype(np.uint8)


# ------------------------------------------------------------------------------------------ teapot

def teapot_like(seed=1) -> bytes:
    """Config 1 stand-in: one model, ~35 k voxels, solid of revolution + spout + handle + lid."""
    sx, sy, sz = 72, 48, 44
    X, Y, Z = np.meshgrid(np.arange(sx), np.arange(sy), np.arange(sz), indexing="ij")
    cx, cy = 34.0, 24.0
    r = np.sqrt((X - cx) ** 2 + (Y - cy) ** 2)
    zc = Z + 0.5
    prof = 20.0 * np.sqrt(np.clip(1.0 - ((zc - 15.0) / 17.0) ** 2, 0, 1)) * (zc < 30)
    body = (r < prof) & ~((r < prof - 2.5) & (zc > 4) & (zc < 27))            # hollow shell, thick bottom
    lid = (r < 9.0 - 0.8 * (zc - 29)) & (zc >= 29) & (zc < 34)
    knob = ((X - cx) ** 2 + (Y - cy) ** 2 + (zc - 36.5) ** 2) < 3.2 ** 2
    t = (X - 52.0) / 16.0                                                      # spout along +x, rising
    spout = (np.abs(Y - cy) ** 2 + (zc - (12.0 + 16.0 * t)) ** 2 < (4.2 - 1.8 * t) ** 2) & (t > -0.2) & (t < 1.05)
    hx, hz = X - 13.0, zc - 17.0                                               # handle: torus segment on -x
    handle = (np.abs(np.sqrt(hx ** 2 + hz ** 2) - 9.0) < 1.8) & (np.abs(Y - cy) < 1.8) & (X < 16)
    solid = body | lid | knob | spout | handle
    rng = np.random.default_rng(seed)
    color = (216 + rng.integers(0, 30, solid.shape)).astype(np.uint8)
    color[knob | lid] = 120
    return write_vox([((sx, sy, sz), _xyzi_from(solid, color))], None, default_palette(seed))


# ------------------------------------------------------------------------------------------ castle

def castle_like(seed=2, tiles=8, tile=128, height=224, props=40) -> bytes:
    """Config 2-4 stand-in: `tiles`x`tiles` ground models of `tile`^2 footprint (solid terrain +
    curtain wall + keep), plus tower / house / rock models instanced `props` times with integer
    translations and the 24 axis rotations (some mirrored).  Defaults: ~21 M voxels, 64 + 3 models,
    ~105 instances; world spans +-512 around the origin the castle.rs camera looks at."""
    assert tile % 4 == 0 and tile <= 256 and height <= 256 and height % 2 == 0
    rng = np.random.default_rng(seed)
    pal = default_palette(seed)
    half = tiles * tile // 2
    scale = half / 512.0                                   # feature scale follows the world size
    models, instances = [], []

    def terrain_h(x, y):
        n = 0.55 * _value_noise(x, y, 96, seed) + 0.3 * _value_noise(x, y, 37, seed + 1) + 0.15 * _value_noise(x, y, 11, seed + 2)
        rr = np.sqrt(x.astype(np.float32) ** 2 + y.astype(np.float32) ** 2)
        plateau = np.clip(1.0 - rr / (300.0 * scale), 0, 1)
        return (10.0 + 34.0 * n * (1 - 0.8 * plateau) + 14.0 * plateau).astype(np.float32)

    wall_r, wall_t, wall_h = 230.0 * scale, 5.0, 46.0
    keep_r, keep_h = 46.0 * scale, 120.0

    zz = np.arange(height, dtype=np.float32)[None, None, :] + 0.5
    for ty in range(tiles):
        for tx in range(tiles):
            x0, y0 = tx * tile - half, ty * tile - half
            xs = np.arange(x0, x0 + tile)
            ys = np.arange(y0, y0 + tile)
            X2, Y2 = np.meshgrid(xs, ys, indexing="ij")
            h = terrain_h(X2, Y2)
            solid = zz < h[:, :, None]
            depth = h[:, :, None] - zz
            color = np.where(depth < 1.5, 4 + (12 * _value_noise(X2, Y2, 5, seed + 9)).astype(np.int32)[:, :, None],
                             np.where(depth < 6, 40, 72)).astype(np.uint8) * np.ones((1, 1, height), np.uint8)
            cheb = np.maximum(np.abs(X2 + 0.5), np.abs(Y2 + 0.5))
            # curtain wall: square ring with crenellations and a gate
            ring = (np.abs(cheb - wall_r) < wall_t)
            gate = (np.abs(X2) < 10) & (Y2 < 0)
            cren = ((X2 + Y2) // 3) % 2 == 0
            top = h + wall_h + np.where(cren & (np.abs(cheb - wall_r) > wall_t - 2), 4.0, 0.0)
            wall = ring[:, :, None] & (zz >= h[:, :, None]) & (zz < top[:, :, None]) & ~(gate[:, :, None] & (zz < (h + 22)[:, :, None]))
            # keep: hollow box with window slits and a pyramid roof
            inner = cheb < keep_r - 4
            keep = ((cheb < keep_r)[:, :, None] & (zz >= h[:, :, None]) & (zz < 14 + keep_h)
                    & ~(inner[:, :, None] & (zz > 20) & (zz < 12 + keep_h)))
            slit = ((np.abs(X2) % 16 < 2) | (np.abs(Y2) % 16 < 2))[:, :, None] & ((zz % 24) > 12) & ((zz % 24) < 20)
            keep &= ~(slit & (cheb > keep_r - 5)[:, :, None] & (zz > 24))
            roof = (zz >= 14 + keep_h) & (zz < 14 + keep_h + (keep_r + 4 - cheb)[:, :, None]) & (cheb < keep_r + 4)[:, :, None]
            solid |= wall | keep | roof
            color = np.where(wall | keep, (80 + (24 * _value_noise(X2 * 3 + 7, Y2 * 3 + 5, 9, seed + 4)).astype(np.int32))[:, :, None]
                             + ((zz.astype(np.int32)) % 3), color).astype(np.uint8)
            color = np.where(roof & ~(wall | keep), 120 + (zz.astype(np.int32) % 8), color).astype(np.uint8)
            if solid.any():
                models.append(((tile, tile, height), _xyzi_from(solid, color)))
                instances.append((len(models) - 1, (x0 + tile // 2, y0 + tile // 2, height // 2), 4))

    # instanced props ----------------------------------------------------------------------------
    def add_model(size, fn):
        sx, sy, sz = size
        X, Y, Z = np.meshgrid(np.arange(sx), np.arange(sy), np.arange(sz), indexing="ij")
        solid, color = fn(X + 0.5, Y + 0.5, Z + 0.5)
        models.append((size, _xyzi_from(solid, color.astype(np.uint8))))
        return len(models) - 1

    def tower(X, Y, Z):
        r = np.sqrt((X - 20) ** 2 + (Y - 20) ** 2)
        shell = (r < 18) & ~((r < 14) & (Z > 6) & (Z < 96)) & (Z < 104)
        shell &= ~((np.abs(X - 20) < 2) & (Z % 26 > 14) & (Z % 26 < 22) & (r > 13))
        cone = (Z >= 104) & (r < (19 - (Z - 104) * 0.45))
        return shell | cone, np.where(cone, 116 + (Z.astype(np.int32) % 10), 84 + (Z.astype(np.int32) % 5) + ((X + Y).astype(np.int32) % 3))

    def house(X, Y, Z):
        box = (Z < 22) & ~((X > 3) & (X < 33) & (Y > 3) & (Y < 25) & (Z > 2) & (Z < 20))
        box &= ~((np.abs(X - 18) < 3) & (Y < 4) & (Z < 13))
        roof = (Z >= 22) & (np.abs(Y - 14) < (14.5 - (Z - 22)))
        return box | roof, np.where(roof, 124 + (Z.astype(np.int32) % 6), 150 + ((X + Z).astype(np.int32) % 12))

    def rock(X, Y, Z):
        n = _value_noise((X * 4).astype(np.int64), (Y * 4 + Z * 7).astype(np.int64), 24, seed + 21)
        r = np.sqrt(((X - 16) / 15.0) ** 2 + ((Y - 12) / 11.0) ** 2 + ((Z - 10) / 9.0) ** 2)
        return r < 0.75 + 0.35 * n, 66 + (30 * n).astype(np.int32)

    m_tower = add_model((40, 40, 144), tower)
    m_house = add_model((36, 28, 38), house)
    m_rock = add_model((32, 24, 20), rock)
    # four corner towers + gate towers, upright (rotations about z only)
    zrot = [rotation_byte((0, 1, 2), (1, 1, 1)), rotation_byte((1, 0, 2), (-1, 1, 1)),
            rotation_byte((0, 1, 2), (-1, -1, 1)), rotation_byte((1, 0, 2), (1, -1, 1))]
    wr = int(wall_r)
    for k, (px, py) in enumerate([(-wr, -wr), (wr, -wr), (wr, wr), (-wr, wr), (-24, -wr), (24, -wr)]):
        if abs(px) < half - 24 and abs(py) < half - 24:
            gz = int(terrain_h(np.array([[px]]), np.array([[py]]))[0, 0])
            instances.append((m_tower, (px, py, gz + 72 - 4), zrot[k % 4]))
    rots = all_rotation_bytes(True)
    mirrored = [r for r in all_rotation_bytes(False) if r not in rots]
    for k in range(props):
        px, py = int(rng.integers(-half + 40, half - 40)), int(rng.integers(-half + 40, half - 40))
        gz = int(terrain_h(np.array([[px]]), np.array([[py]]))[0, 0])
        kind = k % 3
        if kind == 0:
            instances.append((m_house, (px, py, gz + 19 - 2), zrot[int(rng.integers(0, 4))]))
        elif kind == 1:   # rocks in any of the 24 orientations, half-buried or floating
            instances.append((m_rock, (px, py, gz + int(rng.integers(-4, 60))), rots[int(rng.integers(0, len(rots)))]))
        else:             # fallen / mirrored towers
            r = mirrored[int(rng.integers(0, len(mirrored)))] if k % 2 else rots[int(rng.integers(0, len(rots)))]
            instances.append((m_tower, (px, py, gz + 30), r))
    return write_vox(models, instances, pal)


# -------------------------------------------------------------------------------------- cache / IO

def cached(name: str, maker, cache_dir=None) -> bytes:
    cache_dir = cache_dir or os.environ.get("DUST_SCENE_CACHE", "/tmp/dust_b200_scenes")
    os.makedirs(cache_dir, exist_ok=True)
    path = os.path.join(cache_dir, name + ".vox")
    if os.path.exists(path):
        with open(path, "rb") as f:
            return f.read()
    data = maker()
    tmp = path + f".{os.getpid()}.tmp"
    with open(tmp, "wb") as f:
        f.write(data)
    os.replace(tmp, path)
    return data


def load_asset(kind: str, search=("assets", "/root/assets")):
    """Returns (bytes, label): the genuine asset if a file with the reference's SHA-256 is found,
    else the seeded stand-in."""
    want = {"castle": CASTLE_SHA256, "teapot": TEAPOT_SHA256}[kind]
    for d in search:
        p = os.path.join(d, kind + ".vox")
        if os.path.exists(p) and os.path.getsize(p) > 1024:
            data = open(p, "rb").read()
            if hashlib.sha256(data).hexdigest() == want:
                return data, f"{kind}.vox (genuine, sha256 verified)"
    if kind == "castle":
        return cached("castle_like_s2_t8", castle_like), "castle.vox STAND-IN (procedural castle_like(seed=2): 110 instances of 20 models, 30.0 M voxels)"
    return cached("teapot_like_s1", teapot_like), "teapot.vox STAND-IN (procedural teapot_like(seed=1))"


def ray_batch(n: int, extent: float, seed=5) -> np.ndarray:
    """Config 5 rays: origin uniform in the domain box, direction uniform on the sphere, tmin 0, tmax inf."""
    from .pipeline import RAY_DTYPE
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, dtype=RAY_DTYPE)
    rays["o"] = rng.random((n, 3), dtype=np.float32) * np.float32(extent)
    z = rng.random(n, dtype=np.float32) * 2 - 1
    phi = rng.random(n, dtype=np.float32) * np.float32(2 * np.pi)
    r = np.sqrt(np.maximum(0, 1 - z * z))
    rays["d"] = np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1).astype(np.float32)
    rays["tmin"] = 0.0
    rays["tmax"] = np.inf
    return rays


# ---- BASELINE.json configs[4] rays as SURVEY.md 8d specifies them: drawn from the reference's `pcg` hash
# (assets/shaders/headers/spatial_hash.glsl:105-111), counter-based, so ray i is the same whoever generates it and however
# the batch is split.  The same op sequence runs on numpy (CPU arms, tests) or torch (on the GPU: 10^8 rays in a blink);
# only exactly-rounded f32 operations (mul, add, sqrt, int -> float) are used, so both give the same bits.

def _pcg_x(v, xp):
    M = 0xFFFFFFFF
    state = (v * 747796405 + 2891336453) & M
    word = (((state >> ((state >> 28) + 4)) ^ state) * 277803737) & M
    return ((word >> 22) ^ word) & M


def ray_batch_pcg(n: int, extent: float, seed=5, first=0, device=None):
    """Rays first .. first+n-1 of the seeded batch: origin uniform in [0, extent)^3, direction uniform on the sphere, tmin 0, tmax inf.
    device=None: numpy structured array (RAY_DTYPE).  device="cuda": torch float32 tensor [n, 8] on that device."""
    if device is None:
        xp = np
        i = np.arange(first, first + n, dtype=np.int64)
        f32 = lambda a: a.astype(np.float32)
        f64 = lambda a: a.astype(np.float64)
        sqrt, floor, where = np.sqrt, np.floor, np.where
        stack = lambda cols: np.stack(cols, axis=1)
    else:
        import torch
        xp = torch
        i = torch.arange(first, first + n, dtype=torch.int64, device=device)
        f32 = lambda a: a.to(torch.float32)
        f64 = lambda a: a.to(torch.float64)
        sqrt, floor, where = torch.sqrt, torch.floor, torch.where
        stack = lambda cols: torch.stack(cols, dim=1)
    s = _pcg_x((i + _pcg_x_scalar(seed)) & 0xFFFFFFFF, xp)
    u = [f32(_pcg_x((s + k) & 0xFFFFFFFF, xp) >> 8) * np.float32(1.0 / 16777216.0) for k in range(5)]
    ext = np.float32(extent)
    ox, oy, oz = u[0] * ext, u[1] * ext, u[2] * ext
    z = u[3] * np.float32(2.0) - np.float32(1.0)
    # cos / sin(2 pi u) from the kernel's own mul/add polynomial (trace_math.cuh cossin_2pi): bit-reproducible everywhere
    x4 = u[4] * np.float32(4.0)
    qf = floor(x4)
    th = ((x4 - qf) - np.float32(0.5)) * np.float32(1.57079632679489661923)
    t2 = th * th
    c = [np.float32(v) for v in (-1.0 / 6.0, 1.0 / 120.0, -1.0 / 5040.0, 1.0 / 362880.0)]
    sn = th * (np.float32(1.0) + t2 * (c[0] + t2 * (c[1] + t2 * (c[2] + t2 * c[3]))))
    k = [np.float32(v) for v in (-0.5, 1.0 / 24.0, -1.0 / 720.0, 1.0 / 40320.0, -1.0 / 3628800.0)]
    cs = np.float32(1.0) + t2 * (k[0] + t2 * (k[1] + t2 * (k[2] + t2 * (k[3] + t2 * k[4]))))
    r2 = np.float32(0.70710678118654752440)
    c45, s45 = (cs - sn) * r2, (cs + sn) * r2
    q = qf  # 0, 1, 2, 3 as floats
    cosv = where(q == 0, c45, where(q == 1, -s45, where(q == 2, -c45, s45)))
    sinv = where(q == 0, s45, where(q == 1, c45, where(q == 2, -s45, -c45)))
    one_m = np.float32(1.0) - z * z
    # the one square root goes through float64 (correctly rounded by every library; vectorised float32 sqrt is not, on all of them)
    r = f32(sqrt(f64(where(one_m > 0, one_m, one_m * np.float32(0.0)))))
    if device is None:
        from .pipeline import RAY_DTYPE
        rays = np.zeros(n, dtype=RAY_DTYPE)
        rays["o"] = stack([ox, oy, oz])
        rays["d"] = stack([r * cosv, r * sinv, z])
        rays["tmax"] = np.inf
        return rays
    import torch
    zero = torch.zeros_like(ox)
    return stack([ox, oy, oz, zero, r * cosv, r * sinv, z, torch.full_like(ox, float("inf"))]).contiguous()


def _pcg_x_scalar(v: int) -> int:
    M = 0xFFFFFFFF
    state = (v * 747796405 + 2891336453) & M
    word = (((state >> ((state >> 28) + 4)) ^ state) * 277803737) & M
    return ((word >> 22) ^ word) & M
