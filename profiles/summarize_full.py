#!/usr/bin/env python
"""Summarise an `ncu --set full` report into profiles/ncu_full_<tag>.md and refresh profiles/ncu_traffic.json (DRAM bytes per
launch of every kernel = dram__bytes_read.sum + dram__bytes_write.sum, which bench.py copies into roofline.traffic).
usage: python profiles/summarize_full.py gpurun_out/full_r02a.ncu-rep r02a      (needs `ncu` on PATH: it only reads the report)"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

rep, tag = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "smsp__inst_executed.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
cols = {m: hdr.index(m) for m in METRICS if m in hdr}


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    return v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1)


agg = collections.OrderedDict()
for r in body:
    name = re.sub(r"\(.*", "", r[ki]).replace("dsa::", "").replace("void ", "")
    a = agg.setdefault(name, {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "first": r})
    a["n"] += 1
    a["us"] += to_us(num(r[cols["gpu__time_duration.sum"]]), units[cols["gpu__time_duration.sum"]])
    a["rd"] += to_bytes(num(r[cols["dram__bytes_read.sum"]]), units[cols["dram__bytes_read.sum"]])
    a["wr"] += to_bytes(num(r[cols["dram__bytes_write.sum"]]), units[cols["dram__bytes_write.sum"]])
out = [f"# ncu --set full --clock-control none ({rep}): config-2 step, averages per launch\n",
       "| kernel | launches | us | DRAM read MB | DRAM write MB | dram % | sm % | warps active % | regs | warp inst | l1tex % | lts % |",
       "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
traffic = {}
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    r, n = a["first"], a["n"]
    g = lambda m: r[cols[m]] if m in cols else ""
    out.append(f"| {name} | {n} | {a['us'] / n:.1f} | {a['rd'] / n / 1e6:.1f} | {a['wr'] / n / 1e6:.1f} | {g(METRICS[3])} | {g(METRICS[4])} | {g(METRICS[5])} | "
               f"{g(METRICS[6])} | {g(METRICS[7])} | {g(METRICS[8])} | {g(METRICS[9])} |")
    short = re.sub(r"<.*", "", name)
    short = short[2:] if short.startswith("k_") else short
    traffic[short] = {"dram_bytes": (a["rd"] + a["wr"]) / n, "ncu_us": a["us"] / n}
open(os.path.join(ROOT, "profiles", f"ncu_full_{tag}.md"), "w").write("\n".join(out) + "\n")
json.dump({"source": f"ncu --set full --clock-control none, profiles/ncu_full_{tag}.md (config-2 step, average per launch)", "kernels": traffic},
          open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print("\n".join(out))
