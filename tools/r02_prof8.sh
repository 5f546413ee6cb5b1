#!/bin/bash
# v13 (dense root table, 4-wide BVH, re-tuned occupancies): launch list with DRAM bytes of a configs[2] frame, ncu --set full of the bounce tracer,
# the primary pass and wf_cull
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'^(wf_|resolve)' -c 400 --csv --log-file gpurun_out/r02_v13_launches.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wf_trace_pt -s 2 -c 1 -o gpurun_out/r02_v13_trace_pt $B > gpurun_out/r02_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^wf_trace$' -s 2 -c 1 -o gpurun_out/r02_v13_wf_trace_primary $B > gpurun_out/r02_ncu_full2.log 2>&1
tail -1 gpurun_out/r02_ncu_full.log | cut -c1-120; tail -1 gpurun_out/r02_ncu_full2.log | cut -c1-120
ls -la gpurun_out/*v13*
