#!/bin/bash
mkdir -p gpurun_out
bash tools/ab.sh 2 4 2>&1 | tail -12
