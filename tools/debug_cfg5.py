#!/usr/bin/env python3
"""Config-5 parity at scale: GPU hits of the first N rays against the oracle; mismatching rays are saved
to gpurun_out/cfg5_mismatch.npz for analysis on the CPU."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dust_b200 import host, scenes
from dust_b200.pipeline import Context, Scene
from oracle import pyoracle as po

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
dom = host.SynthDomain(seed=5, fill=0.01, domain_leaves=1024)
ctx = Context(0)
desc = dom.describe()
dtree = ctx.upload_tree(desc)
ident = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
scene = Scene(ctx, [(ident, dtree)])
rays = scenes.ray_batch(n, dom.extent, seed=5)
g = scene.trace_rays(rays)
otree = po.TreeView(desc)
osc = po.Scene([(ident, otree._h, dom.materials, dom.palette)])
t0 = time.perf_counter()
o, _ = osc.trace_rays(rays)
print("oracle %.1f s" % (time.perf_counter() - t0))
bad = np.flatnonzero((g.view(np.uint32).reshape(-1, 4) != o.view(np.uint32).reshape(-1, 4)).any(axis=1))
print("rays", n, "mismatches", len(bad))
for i in bad[:12]:
    print(i, rays[i], "gpu", g[i], "oracle", o[i])
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/cfg5_mismatch.npz", idx=bad, rays=rays[bad], gpu=g[bad], oracle=o[bad])
