#!/usr/bin/env python3
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per CUDA source line:
share of warp-level instructions executed and average active threads per instruction."""
import csv
import sys

def main(path, top=40):
    rows = list(csv.reader(open(path)))
    fname, data = "?", []
    hdr = None
    for r in rows:
        if len(r) == 2 and r[0] in ("File Path", "File Name"):
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and r and r[0].isdigit():
            iI, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
            try:
                data.append((int(r[iI]), int(r[iT]), int(r[iS]), fname, int(r[0]), r[1].strip()[:100]))
            except ValueError:
                pass
    tot = sum(d[0] for d in data)
    tthr = sum(d[1] for d in data)
    tsmp = sum(d[2] for d in data)
    print(f"total warp-inst {tot:.3e}  thread-inst {tthr:.3e}  avg active {tthr / tot:.2f}  samples {tsmp}")
    data.sort(reverse=True)
    for d in data[:top]:
        print(f"{100 * d[0] / tot:5.1f}% inst {100 * d[2] / max(tsmp,1):5.1f}% smp  act {d[1] / max(d[0], 1):5.1f}  {d[3]}:{d[4]}  {d[5]}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)


def by_ranges(path, ranges):
    """ranges: {label: (file, lo, hi)} -> share of warp instructions and mean active lanes per label."""
    rows = list(csv.reader(open(path)))
    fname, hdr, acc = "?", None, {}
    tot = 0
    for r in rows:
        if len(r) == 2 and r[0] in ("File Path", "File Name"):
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and r and r[0].isdigit():
            try:
                i, t, s = int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("Thread Instructions Executed")]), int(r[hdr.index("# Samples")])
            except ValueError:
                continue
            tot += i
            label = "other:" + fname
            for k, (f, lo, hi) in ranges.items():
                if f == fname and lo <= int(r[0]) <= hi:
                    label = k
                    break
            a = acc.setdefault(label, [0, 0, 0])
            a[0] += i; a[1] += t; a[2] += s
    for k, (i, t, s) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
        print(f"{100 * i / tot:5.1f}% inst  act {t / max(i, 1):5.1f}  {k}")
