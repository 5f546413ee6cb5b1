#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^wf_shade$' -s 2 -c 2 -o gpurun_out/r02_v11_wf_shade $B > gpurun_out/r02_ncu_full3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^wf_cull$' -s 2 -c 1 -o gpurun_out/r02_v11_wf_cull $B > gpurun_out/r02_ncu_full4.log 2>&1
tail -1 gpurun_out/r02_ncu_full3.log | cut -c1-100
