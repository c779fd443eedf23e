#!/usr/bin/env python
"""Top stall-sample lines of one kernel from `ncu -i X.ncu-rep --page source --csv -k regex:NAME` (SASS or source view).
usage: ncu -i rep --page source --csv -k regex:k_name | python profiles/hot_lines.py [N]"""
import csv
import sys

n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": [], "hdr": None}
        blocks.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and r:
        cur["rows"].append(r)
for b in blocks[:1]:
    hdr = b["hdr"]
    si, ci = hdr.index("Source"), hdr.index("# Samples")
    ei = hdr.index("Instructions Executed")
    tot = sum(float(r[ci] or 0) for r in b["rows"]) or 1.0
    print(b["name"][:100], "samples", tot)
    reasons = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    for r in sorted(b["rows"], key=lambda r: -float(r[ci] or 0))[:n]:
        top = sorted(((float(r[i] or 0), hdr[i]) for i in reasons), reverse=True)[:2]
        print(f"{100 * float(r[ci] or 0) / tot:5.1f}%  exec={r[ei]:>8}  {r[si][:90]:90s} {top[0][1]}:{top[0][0]:.0f} {top[1][1]}:{top[1][0]:.0f}")
