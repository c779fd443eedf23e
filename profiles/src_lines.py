#!/usr/bin/env python
"""Instructions executed and stall samples per CUDA source line of one kernel.
usage: ncu -i rep --page source --csv --print-source cuda,sass -k regex:NAME | python profiles/src_lines.py [N] [file-substring]"""
import csv
import sys
from collections import defaultdict

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
want = sys.argv[2] if len(sys.argv) > 2 else ""
rows = list(csv.reader(sys.stdin))
cur, hdr = None, None
inst, samp, text = defaultdict(float), defaultdict(float), {}
for r in rows:
    if r and r[0] in ("File Name", "File Path"):
        cur = r[1]
    elif r and r[0] == "Line No" and len(r) > 4:
        hdr = r
    elif hdr and cur and r and len(r) == len(hdr) and r[0].strip().isdigit() and r[2] == "-":
        key = (cur.split("/")[-1], int(r[0]))
        si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
        inst[key] += float(r[ii] or 0)
        samp[key] += float(r[si] or 0)
        if r[1].strip():
            text[key] = r[1].strip()
ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
print(f"total warp instructions {ti:.0f}, samples {ts:.0f}")
for key in sorted(inst, key=lambda k: (k[0], k[1])):
    if want and want not in key[0]:
        continue
    if inst[key] / ti < 0.004 and samp[key] / ts < 0.004:
        continue
    print(f"{key[0]:>14}:{key[1]:<4} inst {100 * inst[key] / ti:5.1f}%  samples {100 * samp[key] / ts:5.1f}%  {text.get(key, '')[:100]}")
