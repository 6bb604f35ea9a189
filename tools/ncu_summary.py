#!/usr/bin/env python
"""Aggregates an `ncu --csv --metrics gpu__time_duration.sum[,...]` launch list per kernel: launches, average duration,
share of the summed GPU time.  Usage: ncu_summary.py launches.csv [skip_first_n_launches_per_kernel]"""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        rows.append(r)
    per = OrderedDict()
    for r in rows:
        name = r.get("Kernel Name", "?")
        metric = r.get("Metric Name", "")
        val = r.get("Metric Value", "0").replace(",", "")
        unit = r.get("Metric Unit", "")
        d = per.setdefault(name, {})
        d.setdefault(metric, []).append((float(val) if val not in ("", "n/a") else 0.0, unit))
    total = 0.0
    out = []
    for name, d in per.items():
        t = d.get("gpu__time_duration.sum", [])
        t = t[skip:] if len(t) > skip else t
        if not t:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(t[0][1], 1e-3)
        us = [v * scale for v, _ in t]
        out.append((name, len(us), sum(us) / len(us), sum(us)))
        total += sum(us)
    for name, n, avg, s in sorted(out, key=lambda x: -x[3]):
        print(f"{name[:64]:64s} n={n:4d} avg={avg:9.1f}us share={s / total:.3f}")
    print(f"total {total:.1f} us")


if __name__ == "__main__":
    main()
