"""dust_tree_build_voxels: the tree of a .vox model built ON THE DEVICE from its voxel list
must be, section by section and byte for byte, the blob that the host chain (C++ mirror of VoxLoader::load_model +
VoxGeometry::from_tree, then dust_tree_upload's re-layout) produces — and trace identically."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _same_blob(a, b):
    for k in ("nodes16", "blocks", "mids", "leaves", "palette", "materials"):
        assert a[k].shape == b[k].shape, (k, a[k].shape, b[k].shape)
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["palette_linear"].view(np.uint32), b["palette_linear"].view(np.uint32))
    assert a["bounds"] == b["bounds"]


@pytest.mark.parametrize("which", ["teapot", "castle_small"])
def test_device_build_equals_host_build(ctx, host, which):
    from dust_b200 import scenes
    from dust_b200.pipeline import DeviceTree
    data = scenes.teapot_like() if which == "teapot" else scenes.castle_like(tiles=2, tile=64, height=128, props=12)
    vox = host.VoxFile(data)
    mids = sorted({m for m, _ in vox.instances})
    assert mids
    for mid in mids:
        size, n = vox.model_info(mid)
        model = vox.load_model(mid)
        t_host = ctx.upload_tree(model.describe())
        t_dev = DeviceTree.build_voxels(ctx, vox.model_voxels(mid), size, vox.palette)
        assert t_dev.leaf_count == t_host.leaf_count == len(model.nodes)
        assert int(t_dev.layout().n_materials) == n
        _same_blob(t_dev.download(), t_host.download())


def test_device_build_edge_cases(ctx, host):
    from dust_b200.pipeline import DeviceTree
    pal = np.arange(1024, dtype=np.uint8).reshape(256, 4)
    # empty model: one empty root node, no leaves, zero bounds
    e = DeviceTree.build_voxels(ctx, np.zeros((0, 4), np.uint8), (8, 8, 8), pal)
    d = e.download()
    assert e.leaf_count == 0 and d["leaves"].shape == (0, 4) and d["mids"].shape == (0, 4) and not d["nodes16"][..., :2].any()
    assert d["bounds"] == ((0, 0, 0), (0, 0, 0))
    # full-size model corners
    v = np.array([[0, 0, 0, 5], [255, 255, 255, 7], [255, 0, 0, 9], [10, 20, 30, 2], [11, 20, 30, 3]], np.uint8)
    t = DeviceTree.build_voxels(ctx, v, (256, 256, 256), pal)
    d = t.download()
    assert t.leaf_count == 4 and len(d["materials"]) == 5
    assert d["bounds"] == ((0, 0, 0), (256, 256, 256))
    assert sorted(d["materials"].tolist()) == [2, 3, 5, 7, 9]
    # voxel (x, y, z)_vox -> tree (x, z, 255 - y): (10, 20, 30) -> (10, 30, 235): leaf (8, 28, 232), bit (z&3) | (y&3) << 2 | (x&3) << 4
    leaf = [l for l in d["leaves"] if (int(l[0]) | int(l[1]) << 32) == ((1 << ((235 & 3) | (30 & 3) << 2 | (10 & 3) << 4)) | (1 << ((235 & 3) | (30 & 3) << 2 | (11 & 3) << 4)))]
    assert len(leaf) == 1
    # A duplicated XYZI entry: ModelIndexCollector::set counts it twice in block_counts (collector.rs:29) but its grid keeps one
    # value, so every later leaf's material_ptr runs ahead of the compacted material array (a latent out-of-bounds read in the
    # reference).  Both build paths reject such a model with the same error instead of tracing wrong colours.
    dup = np.array([[0, 0, 0, 5], [255, 255, 255, 7], [255, 0, 0, 9], [10, 20, 30, 1], [10, 20, 30, 2], [11, 20, 30, 3]], np.uint8)
    with pytest.raises(Exception, match="material"):
        DeviceTree.build_voxels(ctx, dup, (256, 256, 256), pal)
    # ... unless the duplicate sits in the LAST block of the collector order (nothing comes after it): accepted by both paths, last value wins
    last = np.array([[255, 255, 255, 7], [3, 0, 255, 1], [3, 0, 255, 4]], np.uint8)           # (3, 0, 255) -> tree (3, 255, 255): block (0, 63, 63)
    t2 = DeviceTree.build_voxels(ctx, last, (256, 256, 256), pal)
    assert sorted(t2.download()["materials"].tolist()) == [4, 7]
    with pytest.raises(Exception):
        DeviceTree.build_voxels(ctx, v, (257, 1, 1), pal)


def test_scene_loaded_with_device_build_traces_identically(ctx, host, oracle):
    from dust_b200 import scenes
    from dust_b200.loader import load_vox
    data = scenes.castle_like(tiles=2, tile=64, height=128, props=12)
    a = load_vox(ctx, data)
    b = load_vox(ctx, data, device_build=True)
    assert a.num_leaves == b.num_leaves and a.num_voxels == b.num_voxels
    W, H = 160, 90
    cam = host.camera_look_at(width=W, height=H, **{k: v for k, v in scenes.CASTLE_CAMERA.items()})
    sky = host.sunlight_bake(direction=host.castle_sun_direction(0.0))
    kw = dict(spp=2, max_bounces=1, seed=5, hits=True, gbuffer=True)
    ga, gb = a.scene.trace(cam, sky, W, H, **kw), b.scene.trace(cam, sky, W, H, **kw)
    assert np.array_equal(ga["radiance"].view(np.uint32), gb["radiance"].view(np.uint32))
    for k in ("hit_instance", "hit_leaf", "hit_voxel", "hit_t"):
        assert np.array_equal(ga[k].view(np.uint8), gb[k].view(np.uint8)), k
    for k in ("albedo", "normal", "voxel_id"):
        assert np.array_equal(ga["gbuffer"][k], gb["gbuffer"][k]), k
