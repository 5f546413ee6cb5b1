#!/bin/bash
# round 2, call 1: baseline numbers on today's box + ncu --set full of wf_trace_pt<4> on the config-5 batch
mkdir -p gpurun_out
python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_base_c2.json 2> gpurun_out/r02_base_c2.err
python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_base_c5.json 2> gpurun_out/r02_base_c5.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wf_trace_pt -s 2 -c 1 -o gpurun_out/r02_cfg5_pt_base \
    python bench.py --config 5 --rays 50000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_c5.log 2>&1
tail -2 gpurun_out/r02_base_c2.json | cut -c1-300; tail -1 gpurun_out/r02_base_c5.json | cut -c1-300
