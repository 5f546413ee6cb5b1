#!/usr/bin/env python3
"""Text summaries of an .ncu-rep for profiles/: `ncu_summary.py rep.ncu-rep out_prefix` writes
out_prefix_raw.txt (every metric of the raw page, one per line) and out_prefix_lines.txt (per CUDA source line: share of warp
instructions, active lanes per instruction, samples — needs -lineinfo + --import-source on)."""
import csv
import io
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main(rep, prefix, top=50):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(prefix + "_raw.txt", "w") as f:
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"kernel: {d.get('Kernel Name')}   grid {d.get('Grid Size')} block {d.get('Block Size')}\n")
            for k, u, v in zip(hdr, units, r):
                if k in ("ID", "Process ID", "Process Name", "Host Name", "Kernel Name", "Context", "Stream", "Block Size", "Grid Size", "Device", "CC",
                         "Section Name", "Metric Name", "Metric Unit", "Metric Value", "Rule Name", "Rule Type", "Rule Description"):
                    continue
                f.write(f"{k:95s} {u:16s} {v}\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    tmp = prefix + "_source.csv.tmp"
    open(tmp, "w").write(src)
    import ncu_lines
    old = sys.stdout
    sys.stdout = open(prefix + "_lines.txt", "w")
    try:
        ncu_lines.main(tmp, top)
    finally:
        sys.stdout.close()
        sys.stdout = old
    os.remove(tmp)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
