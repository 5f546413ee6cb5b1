#!/bin/bash
# A/B of tuning builds on the GPU box: every dust_b200/lib/var/*.so (built with EXTRA_NVFLAGS variants
# of dust_b200/csrc/Makefile's command line) is benchmarked on the configs given ("2 5" by default).
# usage (through gpurun):  bash tools/ab.sh [configs...]   -> gpurun_out/ab.txt
cfgs=${@:-2}
mkdir -p gpurun_out
out=gpurun_out/ab.txt
: > $out
for lib in dust_b200/lib/libdust_b200.so dust_b200/lib/var/*.so; do
    [ -f "$lib" ] || continue
    for c in $cfgs; do
        line=$(DUST_B200_LIB=$PWD/$lib python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1)
        v=$(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("%8.1f Mrays/s %7.3f ms" % (d["value"], d["ms_per_step"]))' 2>/dev/null)
        echo "$(basename $lib) cfg$c $v" | tee -a $out
    done
done
