#!/usr/bin/env python
"""Aggregate an `ncu --metrics m1,m2,... --csv` launch list by kernel: launches, total / mean duration, share of the
captured window, DRAM bytes per launch, achieved DRAM GB/s, and the mean of every percentage metric (weighted by
duration).  usage: summarize_ncu_metrics.py launches.csv [out.md] [title]"""
import collections
import csv
import re
import sys


def short(name: str) -> str:
    n = re.sub(r"^void ", "", name).replace("b200::<unnamed>::", "b200::").replace("(anonymous namespace)::", "")
    n = re.sub(r"\((const |unsigned |CUtensorMap|float|int|b200::|__nv|long|void|T1|T2).*$", "", n)
    return n[:110]


def main(path, out=None, title=""):
    rows = list(csv.DictReader(l for l in open(path, errors="replace") if l.startswith('"')))
    per = collections.OrderedDict()
    for r in rows:
        key = r["ID"]
        d = per.setdefault(key, {"name": short(r["Kernel Name"])})
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r.get("Metric Unit", "")
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            v = v / 1e3 if unit in ("nsecond", "ns") else (v if unit in ("usecond", "us") else v * 1e3 if unit in ("msecond", "ms") else v / 1e3)
        if m.startswith("dram__bytes"):
            v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        d[m] = v
    agg = collections.OrderedDict()
    for d in per.values():
        a = agg.setdefault(d["name"], collections.defaultdict(float))
        t = d.get("gpu__time_duration.sum", 0.0)
        a["n"] += 1
        a["us"] += t
        a["bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        for m, v in d.items():
            if m.endswith(("pct_of_peak_sustained_elapsed", "pct_of_peak_sustained_active", "pct")):
                a["w:" + m] += v * t
        if "launch__registers_per_thread" in d:
            a["regs"] = d["launch__registers_per_thread"]
    tot = sum(a["us"] for a in agg.values())
    pct = sorted({k[2:] for a in agg.values() for k in a if k.startswith("w:")})
    nick = {m: m.split(".")[0].replace("__", ".").replace("_cycles_active", "").replace("_throughput", "") for m in pct}
    lines = [f"# {title or path}", "",
             f"{len(per)} launches, {tot:.1f} us summed kernel time (ncu: cold caches, serialised — compare shares; "
             "percentages are duration-weighted means of ncu's pct_of_peak metrics)", "",
             "| kernel | launches | total us | share | avg us | DRAM MB/launch | DRAM GB/s | regs | " +
             " | ".join(nick[m] + " %" for m in pct) + " |",
             "|---|---:|---:|---:|---:|---:|---:|---:|" + "---:|" * len(pct)]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        gbs = a["bytes"] / (a["us"] * 1e-6) / 1e9 if a["us"] > 0 else 0.0
        cells = [f"{a['w:' + m] / a['us']:.1f}" if a["us"] > 0 and ("w:" + m) in a else "" for m in pct]
        lines.append(f"| `{k}` | {int(a['n'])} | {a['us']:.1f} | {100 * a['us'] / tot:.1f}% | {a['us'] / a['n']:.1f} | "
                     f"{a['bytes'] / a['n'] / 1e6:.2f} | {gbs:.0f} | {int(a.get('regs', 0)) or ''} | " + " | ".join(cells) + " |")
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    else:
        print(text)


if __name__ == "__main__":
    main(*sys.argv[1:4])
