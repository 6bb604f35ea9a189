#!/bin/bash
# Parity tests + short bench of every workload on one GPU.  Usage (through gpurun): bash tools/gpu_all.sh <tag> "<workloads>"
TAG=${1:-all}; WLS=${2:-"c3 c2 c4 c5"}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
for WL in $WLS; do
  ST=30; [ "$WL" = "c5" ] && ST=5
  python bench.py --workload $WL --steps $ST --warmup 3 --no-alt > gpurun_out/${TAG}_${WL}.json 2> gpurun_out/${TAG}_${WL}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_${WL}.json"))
    r=d["roofline"]
    print("$WL ms_per_step=%.4f e2e_ms=%.4f trace=%.4f conv=%.4f build=%.2f nodes/primary=%.2f nodes/shadow=%.2f frac=%.3f kernels=%s cpu=%s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["update_ms"]["trace"], d["update_ms"]["convolve"], d["details"]["accel_build_ms"], r["nodes_per_primary_ray"], r["nodes_per_shadow_ray"], r["frac"], [round(v,4) for v in r["kernels_ms_avg"].values()], d["cpu_baseline"] and round(d["cpu_baseline"]["value"],2)))
except Exception as e:
    print("$WL failed", e); print(open("gpurun_out/${TAG}_${WL}.err").read()[-2000:])
PY
done
