#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ into small tracked files under profiles/.

    python scripts/ncu_summary.py launches gpurun_out/r01_launches.csv profiles/r01_launches_gcr2d.md
    python scripts/ncu_summary.py full     gpurun_out/r01_prof.ncu-rep profiles/r01_ncu_full_gcr2d.md [workload]

`launches`: per kernel (template instance) launch count, mean device time and share of the summed kernel time of the
`--metrics gpu__time_duration.sum` pass.  `full`: per profiled launch duration, DRAM bytes read+written (the
`traffic` figure of bench.py's roofline, also merged into profiles/ncu_traffic.json when a workload is named), DRAM
throughput %, achieved occupancy, registers.
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# kernel symbol -> profile class name used by the library (KLAUNCH names)
CLASS = {"k_sell_spmv<1>": "sell_dirac", "k_sell_spmv<0>": "sell_spmv", "k_hopping": "hopping_dirac", "k_hopping_l1": "hopping_dirac",
         "k_gcr_update_xr": "gcr_update_xr", "k_gcr_dot_hist": "gcr_dot_hist", "k_gcr_dot_hist_tma": "gcr_dot_hist", "k_gcr_update_p": "gcr_update_p",
         "k_gcr_init": "gcr_init", "k_blockcsr_apply": "blockcsr_apply", "k_blockcsr_apply_ne": "blockcsr_apply", "k_restrict": "mg_restrict",
         "k_restrict_warp": "mg_restrict", "k_restrict_sub": "mg_restrict", "k_prolong": "mg_prolong", "k_hopping_tma": "hopping_dirac",
         "k_blockcsr_ring": "blockcsr_apply"}


def short(name):
    m = re.search(r"(k_\w+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:40]


def read_csv_after_header(text):
    lines = text.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    return list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))


def launches(src, dst):
    rows = read_csv_after_header(open(src).read())
    agg = collections.OrderedDict()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        t = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            t *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            t *= 1e6
        k = (short(r["Kernel Name"]), r["Grid Size"], r["Block Size"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(a[1] for a in agg.values())
    out = ["# ncu launch list summary (`--metrics gpu__time_duration.sum --clock-control none`)", "",
           "source: `%s` (%d launches, %.3f ms summed kernel time; cold-cache serialised replays: compare shares)" % (src, sum(a[0] for a in agg.values()), total / 1e6),
           "", "| kernel | grid | block | launches | mean us | share |", "|---|---|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %s | %s | %d | %.1f | %.1f%% |" % (k[0], k[1], k[2], a[0], a[1] / a[0] / 1e3, 100 * a[1] / total))
    open(dst, "w").write("\n".join(out) + "\n")
    print("wrote", dst)


def full(src, dst, workload=None):
    txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(h)}
    want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
            ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]

    def to_bytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]

    def to_us(v, u):
        v = float(v.replace(",", ""))
        return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}[u]

    out = ["# ncu `--set full` capture summary", "", "source: `%s`" % src, "",
           "| kernel | grid x block | regs | time us | DRAM read MB | DRAM write MB | traffic MB | GB/s (traffic/time) | DRAM %% of peak | achieved occupancy %% |".replace("%%", "%"),
           "|---|---|---|---|---|---|---|---|---|---|"]
    traffic = {}
    for r in data:
        name = short(r[col["Kernel Name"]])
        t = to_us(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]])
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        out.append("| `%s` | %s x %s | %s | %.1f | %.1f | %.1f | %.1f | %.0f | %s | %s |" % (
            name, r[col["launch__grid_size"]], r[col["launch__block_size"]], r[col["launch__registers_per_thread"]], t, rd / 1e6, wr / 1e6,
            (rd + wr) / 1e6, (rd + wr) / t / 1e3, r[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]],
            r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]]))
        base = re.sub(r"<[^>]*>$", "", name)
        cls = CLASS.get(name) or CLASS.get(base)
        if cls:
            traffic.setdefault(cls, []).append(rd + wr)
    open(dst, "w").write("\n".join(out) + "\n")
    print("wrote", dst)
    if workload:
        tf = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        allt = json.load(open(tf)) if os.path.exists(tf) else {}
        allt.setdefault(workload, {})
        for cls, v in traffic.items():
            allt[workload][cls] = max(v)   # the largest (fine-level) launch of the class
        allt[workload]["_source"] = (allt[workload].get("_source", "") + " " + os.path.basename(dst)).strip()
        allt[workload]["_note"] = "DRAM bytes read + written by ONE launch of the class at the fine level (the largest launch captured)"
        json.dump(allt, open(tf, "w"), indent=1, sort_keys=True)
        print("updated", tf)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
