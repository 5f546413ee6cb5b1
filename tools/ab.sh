#!/bin/bash
# A/B on the GPU box: bench config 2 (and 1) for every library under dust_b200/lib/variants/*/
for d in dust_b200/lib/variants/*/; do
  v=$(basename $d)
  for c in ${AB_CONFIGS:-2}; do
    DUST_B200_LIB=$PWD/$d/libdust_b200.so python bench.py --config $c --steps ${AB_STEPS:-5} --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v config $c', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],3), 'ms')"
  done
done
