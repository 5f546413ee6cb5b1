"""`.vox` ingest (crates/vox/src/{loader,collector,geometry}.rs): the product's C++ host mirror against
the oracle, byte for byte, plus direct checks of the loader's semantics on known inputs."""
import numpy as np
import pytest

from dust_b200 import scenes


def _models(vox_o, vox_h):
    ids = sorted({m for m, _ in vox_o.instances})
    return [(m, vox_o.load_model(m), vox_h.load_model(m)) for m in ids]


@pytest.mark.parametrize("maker", [
    lambda: scenes.teapot_like(),
    lambda: scenes.castle_like(tiles=2, tile=64, height=96, props=9),
], ids=["teapot", "castle_small"])
def test_host_ingest_equals_oracle(oracle, host, maker):
    data = maker()
    vo, vh = oracle.VoxFile(data), host.VoxFile(data)
    assert len(vo.instances) == len(vh.instances) > 0
    for (mo, ao), (mh, ah) in zip(vo.instances, vh.instances):
        assert mo == mh and np.array_equal(ao, ah)
    assert np.array_equal(vo.palette, vh.palette)
    for mid, o, h in _models(vo, vh):
        assert o.size == h.size
        assert np.array_equal(o.materials, h.materials)
        assert np.array_equal(o.nodes, h.nodes)                      # GPUVoxNode records, iter_leaf order
        assert np.array_equal(o.aabbs, h.aabbs)
        assert o.tree.root_bytes() == h.tree.root_bytes()
        assert o.tree.pool_bytes(1) == h.tree.pool_bytes(1)
        assert o.tree.pool_bytes(0) == h.tree.pool_bytes(0)          # includes every leaf's material_ptr


def test_swizzle_material_lookup_and_records(host):
    """loader.rs:247-272 + hit.rchit:60-72: voxel (x,y,z)_vox lands at (x, z, size.y-1-y) and its colour is
    materials[material_ptr + popcount(mask & below(bit))]."""
    rng = np.random.default_rng(3)
    size = (40, 30, 20)
    pts = np.unique(rng.integers(0, size, (600, 3)), axis=0)
    col = rng.integers(0, 255, len(pts))
    data = scenes.write_vox([(size, np.concatenate([pts, col[:, None]], axis=1))], None, scenes.default_palette())
    v = host.VoxFile(data)
    assert len(v.instances) == 1 and v.instances[0][0] == 0
    assert np.array_equal(v.instances[0][1], np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32))   # loader.rs:67-83
    m = v.load_model(0)
    assert m.size == (40, 20, 30)                                     # [x, z, y] loader.rs:292
    assert len(m.materials) == len(pts)
    want = {(int(x), int(z), int(size[1] - 1 - y)): int(c) for (x, y, z), c in zip(pts, col)}
    assert {tuple(p) for p in m.tree.iter().tolist()} == set(want)
    seen = 0
    for node, bb in zip(m.nodes, m.aabbs):
        ox, oy, oz = int(node["x"]), int(node["y"]), int(node["z"])
        assert ox % 4 == oy % 4 == oz % 4 == 0 and int(node["w"]) == 0
        assert bb.tolist() == [ox, oy, oz, ox + 4, oy + 4, oz + 4]    # geometry.rs:71-82
        mask = int(node["mask"])
        for bit in range(64):
            if mask >> bit & 1:
                p = (ox + (bit >> 4), oy + ((bit >> 2) & 3), oz + (bit & 3))
                off = bin(mask & ((1 << bit) - 1)).count("1")
                assert int(m.materials[int(node["material_ptr"]) + off]) == want[p]
                seen += 1
    assert seen == len(pts)
    # leaf order = ascending root index, then mid index (internal.rs:306-332) => gl_PrimitiveID
    keys = [((int(n["x"]) >> 4) << 8 | (int(n["y"]) >> 4) << 4 | (int(n["z"]) >> 4),
             ((int(n["x"]) >> 2) & 3) << 4 | ((int(n["y"]) >> 2) & 3) << 2 | ((int(n["z"]) >> 2) & 3)) for n in m.nodes]
    assert keys == sorted(keys)


def test_avg_albedo_packing(host):
    """geometry.rs:86-113: mean palette colour, sRGB-encoded, R10G10B10A2"""
    pal = scenes.default_palette()
    data = scenes.write_vox([((4, 4, 4), np.array([[0, 0, 0, 10], [1, 0, 0, 200]]))], None, pal)
    m = host.VoxFile(data).load_model(0)
    assert len(m.nodes) == 1
    mean = (pal[10].astype(np.float32) + pal[200].astype(np.float32)) / np.float32(2 * 255.0)

    def enc(c):
        c = np.float32(c)
        return np.float32(12.92) * c if c <= np.float32(0.0031308) else np.float32(1.055) * np.float32(c ** np.float32(1 / 2.4)) - np.float32(0.055)
    r, g, b = (int(np.float32(enc(mean[k])) * np.float32(1023.0)) for k in range(3))
    packed = int(m.nodes[0]["avg_albedo"])
    assert abs((packed >> 22) - r) <= 1 and abs(((packed >> 12) & 1023) - g) <= 1 and abs(((packed >> 2) & 1023) - b) <= 1
    assert packed & 3 == 3


def test_to_transform_all_rotations(oracle, host):
    """loader.rs:178-204: every signed permutation gives an integer-aligned affine that maps the model box
    onto a box centred at the swizzled translation; host == oracle bit for bit."""
    rots = scenes.all_rotation_bytes(False)
    assert len(rots) == 48 and len(scenes.all_rotation_bytes(True)) == 24
    for r in rots:
        for size in ((8, 6, 4), (7, 5, 3), (1, 1, 1), (256, 256, 256)):
            t = (10, -20, 30)
            a, b = oracle.to_transform(t, r, size), host.to_transform(t, r, size)
            assert np.array_equal(a, b)
            M = a.reshape(3, 4)
            L = M[:, :3]
            assert set(np.abs(L).ravel().tolist()) <= {0.0, 1.0} and abs(abs(np.linalg.det(L)) - 1) < 1e-6
            sz = np.array([size[0], size[2], size[1]], np.float64)        # object axes are (x, z, y)
            corners = np.array([[sz[0] * (c & 1), sz[1] * (c >> 1 & 1), sz[2] * (c >> 2 & 1)] for c in range(8)])
            w = corners @ L.T.astype(np.float64) + M[:, 3]
            centre = (w.min(0) + w.max(0)) / 2
            assert np.allclose(centre, [10, 30, 20], atol=0.5 + 1e-6)       # (t.x, t.z, -t.y), half-voxel offsets for odd sizes
            assert np.allclose(w * 2, np.round(w * 2))                      # integer / half-integer aligned


def test_scene_graph_instances(host, oracle):
    """loader.rs:86-176: transform -> group -> (transform -> shape)*, empty models are skipped"""
    pal = scenes.default_palette()
    a = np.array([[0, 0, 0, 1], [3, 2, 1, 2]])
    empty = np.zeros((0, 4), np.int64)
    rots = scenes.all_rotation_bytes(True)
    inst = [(0, (0, 0, 0), 4), (1, (5, 5, 5), 4), (0, (-40, 12, 7), rots[5]), (0, (100, -100, 64), rots[17])]
    data = scenes.write_vox([((4, 4, 2), a), ((6, 6, 6), empty)], inst, pal)
    vh, vo = host.VoxFile(data), oracle.VoxFile(data)
    assert [m for m, _ in vh.instances] == [0, 0, 0]                  # the empty model spawns nothing (:154-156)
    for (_, x), (_, y) in zip(vh.instances, vo.instances):
        assert np.array_equal(x, y)
    assert np.array_equal(vh.instances[0][1], host.to_transform((0, 0, 0), 4, (4, 4, 2)))
    assert np.array_equal(vh.instances[1][1], host.to_transform((-40, 12, 7), rots[5], (4, 4, 2)))


def test_bad_inputs(host, oracle):
    for cls in (host.VoxFile, oracle.VoxFile):
        with pytest.raises(ValueError):
            cls(b"not a vox file at all")
        with pytest.raises(ValueError):
            cls(scenes.teapot_like()[:200])


def test_max_size_model_corners(host, oracle):
    """256^3 model (the loader's assert bound, loader.rs:365): voxels in all eight corners"""
    pts = np.array([[x, y, z, 7] for x in (0, 255) for y in (0, 255) for z in (0, 255)])
    data = scenes.write_vox([((256, 256, 256), pts)], None, scenes.default_palette())
    mh, mo = host.VoxFile(data).load_model(0), oracle.VoxFile(data).load_model(0)
    assert len(mh.nodes) == 8 and np.array_equal(mh.nodes, mo.nodes)
    assert mh.tree.get_value((255, 255, 255)) is True and mh.tree.get_value((128, 128, 128)) is None
    assert mh.tree.pool_count(1) == 8 and mh.tree.pool_count(0) == 8


def test_file_without_rgba_chunk_gets_the_magicavoxel_default_palette(oracle, host):
    """dot_vox substitutes MagicaVoxel's default palette when a file has no RGBA chunk.  Host mirror and
    oracle must agree, and the table must show the published structure: a 6x6x6 colour cube then four ramps."""
    pts = np.array([[1, 2, 3, 0], [4, 5, 6, 215]], np.uint8)
    data = scenes.write_vox([((8, 8, 8), pts)], None, scenes.default_palette())
    k = data.index(b"RGBA")
    stripped = data[:k]                                            # RGBA is the last chunk write_vox emits
    stripped = stripped[:16] + np.uint32(len(stripped) - 20).tobytes() + stripped[20:]   # MAIN's children_bytes
    vo, vh = oracle.VoxFile(stripped), host.VoxFile(stripped)
    assert np.array_equal(vo.palette, vh.palette)
    pal = np.asarray(vh.palette).reshape(-1, 4)
    assert pal[0].tolist() == [255, 255, 255, 255]                # colour 1 = 0xffffffff
    assert pal[1].tolist() == [255, 255, 0xcc, 255]               # colour 2 = 0xffccffff (A B G R): blue steps first
    assert pal[6].tolist() == [255, 0xcc, 255, 255]               # colour 7 = 0xffffccff
    assert pal[36].tolist() == [0xcc, 255, 255, 255]              # colour 37 = 0xffffffcc
    assert pal[214].tolist() == [0, 0, 0x33, 255]                 # colour 215: last cube entry (black is omitted)
    assert pal[215].tolist() == [0xee, 0, 0, 255] and pal[225].tolist() == [0, 0xee, 0, 255]    # red / green ramps
    assert pal[235].tolist() == [0, 0, 0xee, 255] and pal[254].tolist() == [0x11, 0x11, 0x11, 255]
