#!/bin/bash
# v12 (wide BVH, wf_shade at 10 CTAs/SM): launch list with DRAM bytes of a configs[2] frame, ncu --set full of the bounce tracer and of wf_cull
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_v12_launches.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wf_trace_pt -s 2 -c 1 -o gpurun_out/r02_v12_trace_pt $B > gpurun_out/r02_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^wf_cull$' -s 2 -c 1 -o gpurun_out/r02_v12_wf_cull $B > gpurun_out/r02_ncu_full4.log 2>&1
tail -1 gpurun_out/r02_ncu_full.log | cut -c1-120; tail -1 gpurun_out/r02_ncu_full4.log | cut -c1-120
ls -la gpurun_out/*.ncu-rep
