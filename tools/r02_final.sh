#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-200
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2>/dev/null; tail -1 gpurun_out/r02_bench_ref.json | cut -c1-200
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['cpu_baseline']['value'], d['parity']['ok'], d['roofline']['frac'], d['roofline']['strict']['frac'], d['clocks'])
for k,v in d['extra'].items(): print(k, v.get('value'), v.get('ms_per_step'), (v.get('parity') or {}).get('ok'), (v.get('e2e') or {}).get('value'), (v.get('cpu_baseline') or {}).get('value'))
PY
