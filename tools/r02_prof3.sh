#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras"
# the first non-STATS wf_trace launch = the primary queue of the first warm-up frame (launch 1 is the STATS pass' primary, 2 its shadow)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^wf_trace$' -s 2 -c 1 -o gpurun_out/r02_v10_wf_trace_primary $B > gpurun_out/r02_ncu_full2.log 2>&1
tail -2 gpurun_out/r02_ncu_full2.log | cut -c1-160
