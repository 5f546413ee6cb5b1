"""oracle/oracle_cache.c — the reference's world-space radiance cache (headers/spatial_hash.glsl) restated
for the CPU, checked against an independent numpy restatement written from the shader.  Groundwork only: the CUDA side of the
cache is not built, nothing in the product uses it."""
import ctypes as C

import numpy as np


M32 = 0xFFFFFFFF


def _pcg(v):
    state = (v * 747796405 + 2891336453) & M32
    word = (((state >> ((state >> 28) + 4)) ^ state) * 277803737) & M32
    return ((word >> 22) ^ word) & M32


def _xx(p):
    h = (p + 374761393) & M32
    h = (668265263 * (((h << 17) | (h >> 15)) & M32)) & M32
    h = (2246822519 * (h ^ (h >> 15))) & M32
    h = (3266489917 * (h ^ (h >> 13))) & M32
    return h ^ (h >> 16)


def _fingerprint(pos, d):
    h = _xx(pos[0] & M32)
    h = _xx(((pos[1] & M32) + h) & M32)
    h = _xx(((pos[2] & M32) + h) & M32)
    h = _xx((d + h) & M32)
    return max(1, h)


def _location(pos, d, cap):
    h = _pcg(pos[0] & M32)
    h = _pcg(((pos[1] & M32) + h) & M32)
    h = _pcg(((pos[2] & M32) + h) & M32)
    h = _pcg((d + h) & M32)
    return h % cap


def test_hash_functions_against_numpy(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(1)
    for v in [0, 1, 2, 12345, 0x7FFFFFFF, 0xFFFFFFFF] + rng.integers(0, 2 ** 32, 200).tolist():
        assert L.o_xxhash32(int(v)) == _xx(int(v))
    for _ in range(500):
        pos = rng.integers(-5000, 5000, 3).tolist()
        d = int(rng.integers(0, 6))
        key = oracle.RadianceCache.key(pos, d)
        assert L.o_cache_fingerprint(C.byref(key)) == _fingerprint(pos, d)
        for cap in (1024, 1000003, 32 * 1024 * 1024):
            assert L.o_cache_location(C.byref(key), cap) == _location(pos, d, cap)


def test_face_ids(oracle):
    L = oracle.lib()
    for axis in range(3):
        for sign in (-1.0, 1.0):
            n = [0.0, 0.0, 0.0]
            n[axis] = sign
            face = L.o_normal_to_face_id((C.c_float * 3)(*n))
            assert face == 2 * axis + (1 if sign > 0 else 0)                 # normal.glsl:9-18
            back = (C.c_float * 3)()
            L.o_face_id_to_normal(face, back)
            assert list(back) == n
    key = oracle.OSpatialHashKey()
    L.o_cache_key_of((C.c_float * 3)(10.0, -7.9, 123.99), (C.c_float * 3)(0.0, -1.0, 0.0), C.byref(key))
    assert list(key.position) == [2, -1, 30] and key.direction == 2         # ivec3() truncates toward zero (final_gather.rchit:49)


ACESCG2XYZ = np.array([[0.66245437, 0.13400422, 0.15618773], [0.2722288, 0.6740818, 0.05368953], [-0.0055746622, 0.00406073, 1.0103393]])
XYZ2ACESCG = np.array([[1.6410228, -0.32480323, -0.23642465], [-0.66366285, 1.6153315, 0.016756356], [0.011721907, -0.0082844375, 0.9883947]])


def _encode64(rgb):
    X, Y, Z = ACESCG2XYZ @ np.asarray(rgb, np.float64)
    if Y <= 0:
        return 0
    Le = int(np.clip(409.6 * (np.log2(Y) + 20.0), 0.0, 16383.0))
    if Le == 0:
        return 0
    inv = 1.0 / (-2.0 * X + 12.0 * Y + 3.0 * (X + Y + Z))
    ue, ve = int(np.clip(820.0 * 4.0 * X * inv, 0, 511)), int(np.clip(820.0 * 9.0 * Y * inv, 0, 511))
    return (Le << 18) | (ue << 9) | ve


def _decode64(p):
    Le = p >> 18
    if Le == 0:
        return np.zeros(3)
    Y = 2.0 ** ((Le + 0.5) / 409.6 - 20.0)
    u, v = (((p >> 9) & 0x1FF) + 0.5) / 820.0, ((p & 0x1FF) + 0.5) / 820.0
    inv = 1.0 / (6.0 * u - 16.0 * v + 12.0)
    x, y = 9.0 * u * inv, 4.0 * v * inv
    s = Y / y
    return np.maximum(XYZ2ACESCG @ np.array([s * x, Y, s * (1.0 - x - y)]), 0.0)


def test_logluv_codec_against_numpy(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(2)
    colours = np.exp(rng.uniform(np.log(1e-5), np.log(1e4), (2000, 3))).astype(np.float32)
    colours[::50] = 0.0
    same = 0
    for c in colours:
        got = L.o_logluv_encode((C.c_float * 3)(*c.tolist()))
        want = _encode64(c)
        # float32 vs float64 may move a value sitting on a quantisation step by one code in one field
        dl, du, dv = (got >> 18) - (want >> 18), ((got >> 9) & 511) - ((want >> 9) & 511), (got & 511) - (want & 511)
        assert abs(dl) <= 1 and abs(du) <= 1 and abs(dv) <= 1, (c, hex(got), hex(want))
        same += got == want
        out = (C.c_float * 3)()
        L.o_logluv_decode(got, out)
        ref = _decode64(got)
        # float32 against float64; a small component is a cancelling sum of the XYZ terms, hence the term relative to the largest one
        assert (np.abs(np.array(out[:]) - ref) <= 2e-5 * np.abs(ref) + 3e-6 * np.abs(ref).max() + 1e-12).all(), (out[:], ref)
        if got and (ref > 0).all() and 0 < ((got >> 9) & 511) < 511 and 0 < (got & 511) < 511:
            # the codec's own accuracy for colours inside its gamut: 0.17 % luminance steps, 9-bit chroma
            lum = ACESCG2XYZ[1] @ c.astype(np.float64)
            assert abs((ACESCG2XYZ[1] @ ref) / lum - 1.0) < 2.5e-3
    assert same > 0.98 * len(colours)
    assert L.o_logluv_encode((C.c_float * 3)(0.0, 0.0, 0.0)) == 0 and _encode64([0, 0, 0]) == 0


def test_insert_get_running_mean_and_eviction(oracle):
    cache = oracle.RadianceCache(4096)
    k = oracle.RadianceCache.key((10, 20, -30), 3)
    found, _, _ = cache.get(k, 0)
    assert not found
    values = [np.array([1.0, 2.0, 0.5], np.float32), np.array([3.0, 1.0, 0.5], np.float32), np.array([2.0, 3.0, 2.0], np.float32)]
    for f, v in enumerate(values):
        cache.insert(k, v, f)
    found, got, count = cache.get(k, 7)
    assert found and count == 3
    assert np.allclose(got, np.mean(values, axis=0), rtol=2e-2)              # running mean through three LogLuv round trips
    e = cache.entries()
    used = np.flatnonzero(e["fingerprint"])
    assert len(used) == 1 and e["frame"][used[0]] == 7 and e["count"][used[0]] == 3     # get() refreshes last_accessed_frame
    # sample_count saturates: the 404th and later samples are blended with weight 1 / 404 (MAX_SAMPLE_COUNT)
    for f in range(500):
        cache.insert(k, values[0], 100 + f)
    assert cache.get(k, 999)[2] == 404
    # three distinct keys forced into one probe window, then a fourth: the least recently accessed one is evicted
    small = oracle.RadianceCache(1)                                         # capacity 1: every key lands on location 0
    keys = [oracle.RadianceCache.key((i, 0, 0), 0) for i in range(4)]
    for f, kk in enumerate(keys[:3]):
        small.insert(kk, (1.0 + f, 1.0, 1.0), f)
    assert all(small.get(kk, 10 + i)[0] for i, kk in enumerate(keys[:3]))   # frames 10, 11, 12: key 0 is now the oldest
    small.insert(keys[3], (9.0, 9.0, 9.0), 20)
    assert not small.get(keys[0], 21)[0] and small.get(keys[3], 21)[0] and small.get(keys[1], 21)[0] and small.get(keys[2], 21)[0]
