"""Contracts that can be checked without a GPU: every exported entry point is documented in INTEGRATION.md, and the
reference arm of bench.py (the oracle port on host threads) prints the JSON line the driver expects."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_entry_point_is_documented():
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "dust_b200.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(dust_[a-z0-9_]+)\s*\(", src)))
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in names if n not in doc]
    assert not missing, missing
    header = open(os.path.join(ROOT, "include", "dust_b200.h")).read()
    assert header.count("crates/") + header.count("assets/shaders") + header.count(".rs:") >= 15      # entry points cite what they replace


def test_reference_arm_json_contract_config5():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "5", "--rays", "150000",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mrays/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "rays" in line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["steps"] == 1 and line["n_gpus"] == 1 and line["vs_baseline"] is None
