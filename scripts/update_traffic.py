#!/usr/bin/env python
"""Record the DRAM traffic of the GEMM family from an ncu launch list (dram__bytes_read.sum + dram__bytes_write.sum per
gemm_tc_kernel launch, mean over the launches of one step) in profiles/gemm_dram_traffic.json — the file bench.py reads
for `roofline.traffic`.   usage: update_traffic.py launches.csv <config id> <batch> <source label>"""
import collections
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def main(path, cfg, batch, label):
    rows = list(csv.DictReader(l for l in open(path, errors="replace") if l.startswith('"')))
    per = collections.defaultdict(float)
    names = {}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows:
        if r["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            per[r["ID"]] += float(r["Metric Value"].replace(",", "")) * scale.get(r.get("Metric Unit", "byte"), 1)
            names[r["ID"]] = r["Kernel Name"]
    gemm = [v for k, v in per.items() if "gemm_tc_kernel" in names[k]]
    if not gemm:
        raise SystemExit("no gemm_tc_kernel launches in " + path)
    out = ROOT / "profiles" / "gemm_dram_traffic.json"
    d = json.loads(out.read_text()) if out.exists() else {}
    d[f"config{cfg}_B{batch}"] = {"bytes_per_launch": sum(gemm) / len(gemm), "launches": len(gemm), "source": label}
    out.write_text(json.dumps(d, indent=1, sort_keys=True) + "\n")
    print(out, d[f"config{cfg}_B{batch}"])


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
