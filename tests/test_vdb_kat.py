"""Known-answer tests ported VERBATIM from the reference's own doc-tests / unit tests for `crates/vdb`
 — the only golden vectors the reference holds for this path.  They pin BOTH the
oracle's tree (oracle/oracle_vdb.c) and the product's host-side mirror (dust_b200/csrc/host/vdb.hpp).
"""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(params=["oracle", "host"])
def TreeCls(request, oracle, host):
    return oracle.Tree if request.param == "oracle" else host.Tree


def test_tree_get_set_doctest(TreeCls):
    """crates/vdb/src/tree.rs:15-25 — hierarchy!(2, 2)"""
    tree = TreeCls((2, 2))
    tree.set_value((0, 4, 0), True)
    tree.set_value((0, 2, 2), False)
    assert tree.get_value((0, 4, 0)) is True
    assert tree.get_value((0, 3, 0)) is None
    assert tree.get_value((0, 2, 2)) is False


def test_tree_iter_order_doctest(TreeCls):
    """crates/vdb/src/tree.rs:87-101 — hierarchy!(4, 2); pins x-major / z-minor bit order"""
    tree = TreeCls((4, 2))
    tree.set_value((0, 1, 2), True)
    tree.set_value((63, 1, 3), True)
    tree.set_value((63, 63, 63), True)
    assert tree.iter().tolist() == [[0, 1, 2], [63, 1, 3], [63, 63, 63]]


def test_bitmask_doctest(oracle):
    """crates/vdb/src/bitmask.rs:81-90 — BitMask<128>: set 12, 101 -> iter_set_bits yields 12, 101, None"""
    L = oracle.lib()
    words = (C.c_uint64 * 2)()
    L.o_bitmask_set(words, 12, 1)
    L.o_bitmask_set(words, 101, 1)
    out = (C.c_uint32 * 8)()
    n = L.o_bitmask_iter_set_bits(words, 2, out, 8)
    assert n == 2 and list(out[:2]) == [12, 101]
    assert L.o_bitmask_get(words, 12) == 1 and L.o_bitmask_get(words, 13) == 0
    assert L.o_bitmask_count_ones(words, 2) == 2
    assert words[0] == 1 << 12 and words[1] == 1 << (101 - 64)       # word = i / 64, LSB first (:44-51)
    L.o_bitmask_set(words, 12, 0)
    assert L.o_bitmask_get(words, 12) == 0


def _pool_doctest(alloc, free, num_chunks):
    """crates/vdb/src/pool.rs:22-42 — Pool of u64, 2 items per chunk"""
    assert [alloc(), alloc(), alloc(), alloc()] == [0, 1, 2, 3]
    assert num_chunks() == 2
    free(1)
    free(2)
    assert alloc() == 2
    assert alloc() == 1
    assert alloc() == 4


def test_pool_doctest_oracle(oracle):
    L = oracle.lib()
    pool = C.create_string_buffer(64)                                  # OPool
    L.o_pool_init(pool, 8, 1)
    _pool_doctest(lambda: L.o_pool_alloc(pool), lambda i: L.o_pool_free(pool, i), lambda: L.o_pool_num_chunks(pool))
    L.o_pool_destroy(pool)


def test_pool_doctest_host(host):
    p = host.Pool(8, 1)
    _pool_doctest(p.alloc, p.free, p.num_chunks)


def test_meta_mask_and_lca(TreeCls, oracle, host):
    """crates/vdb/src/accessor.rs:148-170 — hierarchy!(2, 4, 2)"""
    tree = TreeCls((2, 4, 2))
    assert tree.meta_mask() == 0b10100010
    lca = oracle.lowest_common_ancestor_level if TreeCls is oracle.Tree else host.Tree.lowest_common_ancestor_level
    assert lca((0, 0, 0), (255, 255, 255), 0b10100010, 2) == 2
    # neighbours inside one leaf share the leaf; across a leaf border they share the mid node
    assert lca((0, 0, 0), (1, 2, 3), 0b10100010, 2) == 0
    assert lca((3, 0, 0), (4, 0, 0), 0b10100010, 2) == 1


def test_accessor_random_roundtrip(TreeCls):
    """crates/vdb/src/accessor.rs:172-195 — 100 random u8^3 inserts, all read back Some(true) through the
    accessor (the reference uses an unseeded thread_rng; seeded here)."""
    rng = np.random.default_rng(1234)
    tree = TreeCls((2, 4, 2))
    locs = rng.integers(0, 256, (100, 3), dtype=np.uint32)
    for p in locs:
        tree.set_value(p, True)
    order = rng.permutation(100)
    assert tree.accessor_get_many(locs[order]) == [True] * 100
    # the accessor must also answer misses, in any order
    probe = rng.integers(0, 256, (500, 3), dtype=np.uint32)
    assert tree.accessor_get_many(probe) == [tree.get_value(p) for p in probe]


def test_hierarchy_forms(TreeCls):
    """crates/vdb/src/node/mod.rs:100-111.  hierarchy!(2) is a lone 4^3 leaf.  The (3,1) and (2,2,1) forms
    of the doc comment have 2^3-voxel leaves, whose BitMask<8> has ZERO words in the reference
    (bitmask.rs:8: SIZE / 64) — they only compile there; both implementations reject them."""
    leaf = TreeCls((2,))
    leaf.set_value((1, 2, 3), True)
    assert leaf.get_value((1, 2, 3)) is True and leaf.get_value((3, 2, 1)) is None
    assert leaf.iter().tolist() == [[1, 2, 3]]
    for bad in ((3, 1), (2, 2, 1)):
        with pytest.raises(ValueError):
            TreeCls(bad)


def test_node_byte_layouts(TreeCls):
    """#[repr(C)] sizes: leaf 24 B (leaf.rs:13-25), 4^3 internal 264 B, 16^3 internal 16 896 B (internal.rs:22-32)"""
    t = TreeCls((4, 2, 2))
    assert [t.node_bytes(lv) for lv in range(3)] == [24, 264, 16896]
    t4 = TreeCls((4, 4, 2, 2))
    assert [t4.node_bytes(lv) for lv in range(4)] == [24, 264, 16896, 16896]
    t.set_value((5, 6, 7), True)
    root = t.root_bytes()
    idx = (0 << 8) | (0 << 4) | 0                      # (5>>4, 6>>4, 7>>4)
    assert root[idx // 8] & (1 << (idx % 8))
    mid = t.pool_bytes(1)
    midx = ((5 >> 2) << 4) | ((6 >> 2) << 2) | (7 >> 2)
    assert int.from_bytes(mid[:8], "little") == 1 << midx
    leaf = t.pool_bytes(0)
    bit = ((5 & 3) << 4) | ((6 & 3) << 2) | (7 & 3)
    assert int.from_bytes(leaf[0:8], "little") == 1 << bit          # occupancy
    assert int.from_bytes(leaf[8:16], "little") == 1 << bit         # active (Some(true))


def test_clear_is_todo(TreeCls):
    """internal.rs:121-124 — clearing through an InternalNode is `todo!()` in the reference"""
    t = TreeCls((4, 2, 2))
    t.set_value((1, 1, 1), True)
    with pytest.raises(NotImplementedError):
        t.set_value((1, 1, 1), None)


def test_host_tree_bytes_equal_oracle(oracle, host):
    """Same voxels in -> byte-identical root and pools out (the DustTreeDesc contract)."""
    rng = np.random.default_rng(7)
    for hier in ((4, 2, 2), (4, 4, 2, 2), (2, 4, 2)):
        ext = 1 << sum(hier)
        pts = rng.integers(0, min(ext, 700), (4000, 3), dtype=np.uint32)
        a, b = oracle.Tree(hier), host.Tree(hier)
        for p in pts:
            a.set_value(p, True)
        b.set_many(pts)
        assert a.root_bytes() == b.root_bytes()
        for lv in range(len(hier) - 1):
            assert a.pool_bytes(lv) == b.pool_bytes(lv), (hier, lv)
        assert np.array_equal(a.iter(), b.iter())
        pa, qa = a.iter_leaf()
        pb, qb = b.iter_leaf()
        assert np.array_equal(pa, pb) and np.array_equal(qa, qb)
