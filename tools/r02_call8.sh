#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_example_cpp.py tests/test_device_build.py -m gpu -x -q 2>&1 | tail -4
bash tools/ab.sh 2 2>&1 | tail -4
