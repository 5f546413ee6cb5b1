#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "rc=$?"
tail -c 1500 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','scaling','sharded_equals_whole','step_parts_ms','e2e','strong')})
for k,v in d.get('extra',{}).items():
    print(k, {q:v.get(q) for q in ('value','ms_per_step','parity','e2e','error','sharded_equals_whole','step_parts_ms')})
PY
