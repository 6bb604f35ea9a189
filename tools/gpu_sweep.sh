#!/bin/bash
# Sweep DDGI_TUNE values (kernel tunables) over the bench; prints ms per full-grid update for each.
# Usage (through gpurun): bash tools/gpu_sweep.sh <tag> <workload> "<tune1> <tune2> ..."
TAG=${1:-sweep}; WL=${2:-c3}; shift; shift
mkdir -p gpurun_out
for T in $1; do
  DDGI_TUNE=$T python bench.py --workload $WL --steps 20 --warmup 5 --no-cpu-baseline --no-alt > gpurun_out/${TAG}_$T.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_$T.json"))
print("tune=$T ms_per_step=%.4f trace=%.4f conv=%.4f nodes/primary=%.2f" % (d["ms_per_step"], d["update_ms"]["trace"], d["update_ms"]["convolve"], d["roofline"]["nodes_per_primary_ray"]))
PY
done
