#!/usr/bin/env python
"""Print the per-entry-point table of one or more bench.py JSON lines (files given on the command line)."""
import json
import sys

for f in sys.argv[1:]:
    d = None
    for line in open(f):
        if line.startswith("{"):
            d = json.loads(line)
            break
    if d is None:
        print(f, "no JSON line")
        continue
    print(f"{f}: {d['ms_per_step']:.4f} ms/step  value {d['value']:.0f}  e2e {d['e2e']['value']:.0f}  "
          f"frac {d['roofline']['frac']:.3f}  launches/step {d.get('gpu_launches_per_step')}")
    ks = d.get("kernels", {})
    for k, v in sorted(ks.items(), key=lambda kv: -kv[1]["ms_per_step"]):
        print(f"  {k:28s} {v['launches_per_step']:3d} {v['ms_per_step']:.4f}")
    print(f"  {'sum':28s}     {sum(v['ms_per_step'] for v in ks.values()):.4f}")
