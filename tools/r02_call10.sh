#!/bin/bash
# wide BVH: parity of the default build, then A/B against the binary BVH and width 4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gi.py -m gpu -x -q 2>&1 | tail -5
bash tools/ab.sh 2 2>&1 | tail -5
for lib in dust_b200/lib/libdust_b200.so dust_b200/lib/var/binary.so; do
  DUST_B200_LIB=$PWD/$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c '
import sys,json
d=json.loads(sys.stdin.read())
print(" ".join("%s=%.4f" % (k["kernel"], k["ms"]) for k in d["roofline"]["kernels"]))'
done
