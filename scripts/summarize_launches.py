#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the captured window)."""
import collections
import csv
import re
import sys


def main(path, out=None):
    rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        n = re.sub(r"\(.*", "", r["Kernel Name"])
        n = re.sub(r"^void ", "", n).replace("b200::<unnamed>::", "b200::")
        agg[n][0] += 1
        agg[n][1] += float(r["Metric Value"]) / 1e3
    tot = sum(v[1] for v in agg.values())
    lines = [f"# {path}: {len(rows)} launches, {tot:.1f} us total (cold-cache, serialised: compare shares)", "",
             "| kernel | launches | total us | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k[:100]}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0]:.1f} |")
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    else:
        print(text)


if __name__ == "__main__":
    main(*sys.argv[1:3])
