#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -5
sed -i 's/--steps 10 --warmup 3 --no-cpu-baseline/--steps 10 --warmup 3 --no-cpu-baseline --no-extras/' tools/ab.sh
bash tools/ab.sh 2 4 2>&1 | tail -12
