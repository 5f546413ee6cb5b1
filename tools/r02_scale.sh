#!/bin/bash
# usage: bash tools/r02_scale.sh N  -> gpurun_out/r02_bench_nN.json
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "rc=$?"
tail -c 800 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_n$N.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','scaling','sharded_equals_whole','step_parts_ms')})
print('e2e', d['e2e']['value'], 'strong', {k:d['strong'].get(k) for k in ('value','ms_per_step','sharded_equals_whole','step_parts_ms')})
for k,v in d.get('extra',{}).items():
    print(k, {q:v.get(q) for q in ('value','ms_per_step','error','sharded_equals_whole','step_parts_ms')}, (v.get('parity') or {}).get('ok'))
PY
