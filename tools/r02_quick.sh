#!/bin/bash
# quick confirmation of a kernel change: the bit-exact parity tests, then the headline bench line with its step parts
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gi.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/quick_bench.json').read().strip().splitlines()[-1])
print("value %.1f  ms %.4f  e2e %.1f parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], (d.get("parity") or {}).get("ok")))
for k in d["roofline"]["kernels"]:
    print("  %-14s %.4f ms" % (k["kernel"], k["ms"]))
PY
