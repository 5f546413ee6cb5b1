import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _build_if_missing():
    """The native libraries are built in-tree and git-ignored: a fresh checkout has none.  Build them once (make; nvcc cross-compiles
    sm_100a without a GPU) instead of failing every test with "library missing"."""
    need = [os.path.join(ROOT, "dust_b200", "lib", "libdust_b200.so"), os.path.join(ROOT, "dust_b200", "lib", "libdust_host.so"),
            os.path.join(ROOT, "oracle", "_build", "liboracle.so")]
    if all(os.path.exists(p) for p in need):
        return
    import subprocess
    subprocess.run(["make", "-C", os.path.join(ROOT, "dust_b200", "csrc"), "all"], check=True)
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")
    _build_if_missing()


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/, test infrastructure only)."""
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def host():
    from dust_b200 import host as h
    h.lib()
    return h


@pytest.fixture(scope="session")
def ctx():
    """A live CUDA context; the C ABI has no CPU fallback, so this fails (not skips) without a GPU."""
    from dust_b200.pipeline import Context
    c = Context(0)
    yield c
    c.close()


class BuiltScene:
    """A .vox scene ingested twice: by the product host mirror (+ uploaded) and by the oracle."""

    def __init__(self, data: bytes, oracle, ctx=None):
        from dust_b200 import host
        self.data = data
        self.ovox = oracle.VoxFile(data)
        self.omodels = {m: self.ovox.load_model(m) for m in sorted({m for m, _ in self.ovox.instances})}
        self.oscene = oracle.Scene([(mat, self.omodels[m].tree._h, self.omodels[m].materials, self.ovox.palette)
                                    for m, mat in self.ovox.instances])
        self.loaded = None
        if ctx is not None:
            from dust_b200.loader import load_vox
            self.loaded = load_vox(ctx, data)
        self.host = host

    def bounds(self):
        """World-space AABB of all instance boxes (object boxes are [0, size])."""
        lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
        for mid, m in self.ovox.instances:
            M = m.reshape(3, 4).astype(np.float64)
            sz = np.array(self.omodels[mid].size, np.float64)
            for c in range(8):
                p = np.array([sz[0] * (c & 1), sz[1] * ((c >> 1) & 1), sz[2] * ((c >> 2) & 1)])
                w = M[:, :3] @ p + M[:, 3]
                lo, hi = np.minimum(lo, w), np.maximum(hi, w)
        return lo, hi
