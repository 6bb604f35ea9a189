#!/usr/bin/env python
"""Hot source lines of one kernel from `ncu -i rep --page source --csv --kernel-name regex:NAME --print-source cuda,sass`:
rows that carry a line number (not the per-SASS rows, whose first column is '-') sorted by executed warp instructions.
usage: ncu_hot_lines.py x.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], newline="")))
cur = ""; hdr = None; out = []; 
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit(): continue
    ci = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); ti = hdr.index("Avg. Threads Executed")
    try: n = int(r[ci]); s = int(r[si])
    except ValueError: continue
    out.append((n, s, cur, r[0], r[1].strip()[:140]))
T = sum(o[0] for o in out) or 1; S = sum(o[1] for o in out) or 1
for n, s, f, ln, src in sorted(out, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{n/T:6.1%} smp={s/S:6.1%} {f}:{ln:>4s} {src}")
print("total warp instructions", T, "samples", S)
