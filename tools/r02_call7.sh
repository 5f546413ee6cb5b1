#!/bin/bash
mkdir -p gpurun_out
bash tools/ab.sh 2 4 1 2>&1 | tail -14
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1b.json 2> gpurun_out/r02_bench_n1b.err; echo "rc=$?"; tail -c 400 gpurun_out/r02_bench_n1b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1b.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['parity']['radiance_max_rel'], d['parity']['ok'])
print(d['extra']['gi'])
PY
