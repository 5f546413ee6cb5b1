#!/usr/bin/env python3
"""`ncu_calls.py rep.ncu-rep [launch]`: SASS-level look at out-of-line subroutines (the IEEE division / sqrt slow paths ptxas puts
behind CALL.REL.NOINC) and divergent regions of one profiled launch: every executed CALL with its count and active lanes, the share of
the launch's warp instructions spent between a subroutine's entry and its RET, and the places where active lanes change by > 4."""
import csv
import io
import subprocess
import sys


def main(rep, launch=0):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(launch),
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if "Instructions Executed" in r)
    iI, iT, iS, iA = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Source"), hdr.index("Address")
    body, seen = [], set()
    for r in rows:
        if len(r) <= iT or not r[iI].isdigit() or r[iA] in seen:
            continue
        seen.add(r[iA])
        body.append((r[iA], r[iS], int(r[iI]), int(r[iT])))
    tot = sum(b[2] for b in body)
    print(f"{len(body)} SASS instructions, {tot:.3e} warp instructions executed")
    targets = {}
    for a, s, n, t in body:
        if "CALL" in s and n:
            tgt = s.split()[-1]
            targets.setdefault(tgt, [0, 0])
            targets[tgt][0] += n
            targets[tgt][1] += t
    addr_index = {a: k for k, (a, _, _, _) in enumerate(body)}
    for tgt, (n, t) in sorted(targets.items(), key=lambda kv: -kv[1][0]):
        k = addr_index.get(tgt)
        cost = 0
        if k is not None:
            j = k
            while j < len(body):
                cost += body[j][2]
                if "RET" in body[j][1] and not body[j][1].lstrip().startswith("@"):
                    break
                j += 1
        print(f"subroutine {tgt}: {n} warp calls, {t / max(n, 1):.1f} lanes per call, {100 * cost / tot:.1f} % of the launch's warp instructions")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
