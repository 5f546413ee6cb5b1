#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'wf_|frame_kernel' -c 150 --csv --log-file gpurun_out/r02_v9_launches.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
cat > /tmp/gi_prof.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from dust_b200 import host
from dust_b200.pipeline import RadianceCache
class A: gpus=1
env = bench.Env(A)
S = bench.load_castle(env)
cam = host.camera_look_at(width=1920, height=1080, **S["camkw"])
cache = RadianceCache(S["ctx"], 32*1024*1024, 720*480)
img = torch.zeros((1080,1920,4), dtype=torch.float16, device="cuda")
for f in range(6):
    cache.gi_frame_device(S["scene"], cam, S["sky"], 1920, 1080, img, seed=2, frame_index=f)
torch.cuda.synchronize()
PY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'gi_|wf_trace|RadixSort|DeviceScan|Scan' -c 260 --csv --log-file gpurun_out/r02_gi_launches.csv python /tmp/gi_prof.py > gpurun_out/r02_ncu_gi.log 2>&1
tail -2 gpurun_out/r02_ncu_gi.log
