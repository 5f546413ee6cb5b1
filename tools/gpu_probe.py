#!/usr/bin/env python3
"""GPU probe: one stats frame of the bench workload through a chosen pipeline, DUST_DEBUG_STATS diagnostics on stderr."""
import os, sys, time
os.environ.setdefault("DUST_DEBUG_STATS", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from dust_b200 import host
from dust_b200.loader import load_vox
from dust_b200.pipeline import Context

pipeline = sys.argv[1] if len(sys.argv) > 1 else "megakernel"
cfg = bench.CONFIGS[int(sys.argv[2]) if len(sys.argv) > 2 else 2]
data, label, camkw, sky = bench.castle_setup()
cam = host.camera_look_at(width=cfg["width"], height=cfg["height"], **camkw)
ctx = Context(0)
loaded = load_vox(ctx, data)
r = loaded.scene.trace(cam, sky, cfg["width"], cfg["height"], spp=cfg["spp"], max_bounces=cfg["max_bounces"], seed=bench.SEED, stats=True, pipeline=pipeline)
print(r["stats"])
