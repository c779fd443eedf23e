#!/usr/bin/env python
"""Warp instructions and stall samples of k_tile_merge per phase (the '// ---- X:' markers of csrc/tile.cuh).
usage: ncu -i rep --page source --csv --print-source cuda,sass -k regex:k_tile_merge | python profiles/src_phases.py NTILES_TIMES_LAUNCHES"""
import csv
import os
import sys
from collections import defaultdict

div = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
rows = list(csv.reader(sys.stdin))
cur = hdr = None
inst, samp = defaultdict(float), defaultdict(float)
for r in rows:
    if r and r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No" and len(r) > 4:
        hdr = r
    elif hdr and cur and r and len(r) == len(hdr) and r[0].strip().isdigit() and r[2] == "-":
        inst[(cur, int(r[0]))] += float(r[hdr.index("Instructions Executed")] or 0)
        samp[(cur, int(r[0]))] += float(r[hdr.index("# Samples")] or 0)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lines = open(os.path.join(root, "dynamicsparsearrays.jl_b200", "csrc", "tile.cuh")).read().splitlines()
marks = [(i + 1, l.strip()[8:30]) for i, l in enumerate(lines) if l.strip().startswith("// ---- ")]
kstart = next(i + 1 for i, l in enumerate(lines) if "k_tile_merge(TileArgs" in l)


def phase(f, n):
    if f != "tile.cuh":
        return f
    if n < kstart:
        return "helpers (tile_find, ranks)"
    p = "prologue"
    for ln, name in marks:
        if n >= ln:
            p = name
    return p


ai, as_ = defaultdict(float), defaultdict(float)
for key, v in inst.items():
    ai[phase(*key)] += v
    as_[phase(*key)] += samp[key]
ti, ts = sum(ai.values()), sum(as_.values()) or 1
print(f"total warp instructions {ti:.0f} ({ti / div:.0f} per tile), samples {ts:.0f}")
for k, v in sorted(ai.items(), key=lambda kv: -kv[1]):
    print(f"{100 * v / ti:5.1f}% inst ({v / div:6.0f}/tile)  {100 * as_[k] / ts:5.1f}% samples   {k}")
