#!/usr/bin/env python
"""Instruction mix and stall samples of one kernel from `ncu -i X.ncu-rep --page source --csv` (SASS view).
usage: summarize_ncu_source.py source.csv [rows] — `rows` divides the counts (e.g. rows processed)."""
import collections
import csv
import sys


def main(path, unit=1.0):
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.DictReader(lines[1:]))
    tot = 0
    byop = collections.Counter()
    samples = collections.Counter()
    stall_cols = [k for k in rows[0].keys() if k and k.startswith("stall_") and "Not Issued" not in k]
    stalls = collections.Counter()
    for r in rows:
        n = int(float(r.get("Instructions Executed") or 0))
        toks = (r.get("Source") or "").split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "")
        op = op.split(".")[0]
        byop[op] += n
        tot += n
        samples[op] += int(float(r.get("# Samples") or 0))
        for c in stall_cols:
            stalls[c] += int(float(r.get(c) or 0))
    print(f"{path}: {tot} warp instructions ({tot / unit:.1f} per unit)")
    print("opcode        executed   per-unit   stall-samples")
    for k, v in byop.most_common(24):
        print(f"{k:12s} {v:10d} {v / unit:9.1f} {samples[k]:8d}")
    st = sum(stalls.values()) or 1
    print("stalls: " + ", ".join(f"{k[6:]} {100 * v / st:.0f}%" for k, v in stalls.most_common(8)))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
