#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` capture exported with `ncu -i X.ncu-rep --page raw --csv`:
duration, DRAM bytes and throughput %, tensor-pipe active %, SM throughput %, achieved occupancy, registers, the
three largest warp-stall reasons.  Launches of one kernel are aggregated (duration-weighted means).
usage: summarize_ncu_raw.py raw.csv [out.md] [title]"""
import collections
import csv
import re
import sys

WANT = {
    "gpu__time_duration.sum": "us",
    "dram__bytes_read.sum": "rd",
    "dram__bytes_write.sum": "wr",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram%",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm%",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor%",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ%",
    "launch__registers_per_thread": "regs",
    "smsp__inst_executed.sum": "inst",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "bankconf",
    "lts__t_sector_hit_rate.pct": "l2hit%",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma%",
}


def short(name: str) -> str:
    n = re.sub(r"^void ", "", name).replace("b200::<unnamed>::", "b200::").replace("(anonymous namespace)::", "")
    n = re.sub(r"\((const |unsigned |CUtensorMap|float|int|b200::|__nv|long|void|T1|T2).*$", "", n)
    return n[:100]


def fnum(s: str) -> float:
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


def main(path, out=None, title=""):
    rd = csv.reader(open(path, errors="replace"))
    rows = [r for r in rd if r]
    hdr = rows[0]
    units = rows[1]
    data = rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    name_i = col.get("Kernel Name")
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".pct")]
    if not stall_cols:
        stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    agg = collections.OrderedDict()
    for r in data:
        k = short(r[name_i])
        a = agg.setdefault(k, collections.defaultdict(float))
        t = fnum(r[col["gpu__time_duration.sum"]]) if "gpu__time_duration.sum" in col else 0.0
        u = units[col["gpu__time_duration.sum"]] if "gpu__time_duration.sum" in col else "ns"
        t_us = t / 1e3 if u.startswith("n") else (t if u.startswith("u") else t * 1e3)
        a["n"] += 1
        a["us"] += t_us
        for m, nick in WANT.items():
            if m not in col or nick == "us":
                continue
            v = fnum(r[col[m]])
            if v != v:
                continue
            if nick in ("rd", "wr"):
                mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[col[m]], 1)
                a[nick] += v * mult
            elif nick in ("regs",):
                a[nick] = v
            elif nick in ("inst", "bankconf"):
                a[nick] += v
            else:
                a["w:" + nick] += v * t_us
        for h in stall_cols:
            v = fnum(r[col[h]])
            if v == v:
                a["s:" + h] += v * t_us
    tot = sum(a["us"] for a in agg.values())
    lines = [f"# {title or path}", "",
             f"{int(sum(a['n'] for a in agg.values()))} launches, {tot:.1f} us summed kernel time under `ncu --set full` "
             "(caches flushed and clocks as found, kernels serialised: compare shares and counters, not absolute times).",
             "Percentages are duration-weighted means over the launches of a kernel.", "",
             "| kernel | n | total us | avg us | DRAM MB/launch | dram % | tensor % | sm % | fma % | occupancy % | L2 hit % | regs | top stalls |",
             "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        def w(n):
            return f"{a['w:' + n] / a['us']:.1f}" if a["us"] > 0 and ("w:" + n) in a else ""
        stalls = sorted(((v / a["us"], h) for h, v in a.items() if h.startswith("s:")), reverse=True)[:3]
        st = ", ".join(f"{h.split('stalled_')[1].split('.')[0].replace('_per_issue_active', '')} {v:.1f}" for v, h in stalls)
        lines.append(f"| `{k}` | {int(a['n'])} | {a['us']:.1f} | {a['us'] / a['n']:.1f} | {(a['rd'] + a['wr']) / a['n'] / 1e6:.2f} | "
                     f"{w('dram%')} | {w('tensor%')} | {w('sm%')} | {w('fma%')} | {w('occ%')} | {w('l2hit%')} | {int(a.get('regs', 0)) or ''} | {st} |")
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    else:
        print(text)


if __name__ == "__main__":
    main(*sys.argv[1:4])
