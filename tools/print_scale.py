#!/usr/bin/env python
"""One-line summary of a multi-GPU bench line: python tools/print_scale.py <file.json>"""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
b = d.get("strong_scaling_base")
print("N=%d ms_per_step=%.4f value=%.1f e2e_ms=%.3f base1=%s check=%s" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], b, d.get("exchange_check_vs_1gpu")))
print(d.get("per_rank_ms"), d["update_ms"])
