#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
DUST_WF_PT_DEPTH=0 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "castle_small_gi or teapot_bounces or wavefront_sample" 2>&1 | tail -2
bash tools/ab.sh 2 3 2>&1 | tail -6
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'wf_|frame_kernel' -c 150 --csv --log-file gpurun_out/r02_v10_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
