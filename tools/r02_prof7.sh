#!/bin/bash
# v12 launch list with DRAM bytes: only the wavefront kernels (the scene's device-side tree build alone is > 400 launches)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'^(wf_|resolve)' -c 400 --csv --log-file gpurun_out/r02_v12_launches.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
grep -c gpu__time gpurun_out/r02_v12_launches.csv
