#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "rc=$?"
tail -c 600 gpurun_out/r02_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','parity','cpu_baseline','e2e')})
for k,v in d.get('extra',{}).items():
    print(k, {q:v.get(q) for q in ('value','ms_per_step','parity','e2e','error')})
print(d['roofline']['frac'], d['roofline']['strict'])
PY
