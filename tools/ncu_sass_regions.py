#!/usr/bin/env python
"""Prints the SASS listing of an `ncu --page source --csv` export with executed-instruction counts, average active
threads and stall samples per instruction, plus totals; used to split a kernel's issue slots into code regions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
tot_inst = tot_thr = tot_samp = 0
out = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    inst = int(r[ix["Instructions Executed"]]); thr = int(r[ix["Thread Instructions Executed"]]); samp = int(r[ix["# Samples"]])
    tot_inst += inst; tot_thr += thr; tot_samp += samp
    out.append((r[ix["Source"]].strip(), inst, thr, samp))
print(f"total warp-inst {tot_inst/1e6:.1f}M  thread-inst {tot_thr/1e6:.1f}M  avg threads {tot_thr/max(1,tot_inst):.2f}  samples {tot_samp}")
if len(sys.argv) > 2 and sys.argv[2] == "list":
    cum = 0
    for i, (src, inst, thr, samp) in enumerate(out):
        cum += inst
        print(f"{i:4d} {inst/1e6:8.2f}M thr={thr/max(1,inst):5.1f} smp={samp:6d} cum={cum/tot_inst:5.3f}  {src[:90]}")
