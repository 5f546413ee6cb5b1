#!/bin/bash
# v14 (sun grid): launch list with DRAM bytes of a configs[2] frame
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'^(wf_|resolve)' -c 400 --csv --log-file gpurun_out/r02_v14_launches.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
grep -c gpu__time gpurun_out/r02_v14_launches.csv
