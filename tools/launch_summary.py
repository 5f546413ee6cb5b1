#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list:
launch_summary.py list.csv <first-kernel-prefix> <last-kernel-prefix>  -> the LAST complete frame between those two kernels."""
import collections
import csv
import json
import sys


def load(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    iid, ik, im, iv, iu = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    L = collections.OrderedDict()
    for r in rows[1:]:
        d = L.setdefault(int(r[iid]), {"kernel": r[ik]})
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        if r[im].startswith("gpu__time"):
            d["us"] = v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1)
        else:
            d[r[im]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    return L


def short(k):
    return k.split("(")[0].replace("void ", "").replace("dust::", "")


def main(path, first, last):
    L = load(path)
    ids = list(L)
    names = [short(L[i]["kernel"]) for i in ids]
    ends = [k for k, n in enumerate(names) if n.startswith(last)]
    b = ends[-1]
    starts = [k for k, n in enumerate(names) if n.startswith(first) and k < b and (len(ends) < 2 or k > ends[-2])]
    a = starts[0]
    tot_us = tot_b = 0.0
    for k in range(a, b + 1):
        d = L[ids[k]]
        by = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        tot_us += d["us"]; tot_b += by
        print(f"{ids[k]:5d}  {names[k][:64]:64s} {d['us']:9.1f} us  {by / 1e6:9.1f} MB DRAM")
    print(f"frame (ids {ids[a]}..{ids[b]}): {tot_us / 1e3:.3f} ms serialised under ncu (cold caches), DRAM {tot_b / 1e9:.3f} GB")
    return {"ms": tot_us / 1e3, "dram_bytes": tot_b, "ids": [ids[a], ids[b]]}


if __name__ == "__main__":
    json.dump(main(*sys.argv[1:4]), open("/tmp/launch_summary.json", "w"))
