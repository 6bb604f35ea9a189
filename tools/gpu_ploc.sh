#!/bin/bash
# Builder sweep on the GPU box: parity tests, then the C3 bench for several PLOC radii (0 = plain LBVH).
# Usage (through gpurun): bash tools/gpu_ploc.sh <tag> "<r1> <r2> ..." [workload]
TAG=${1:-ploc}; RS=${2:-"0 16"}; WL=${3:-c3}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
for R in $RS; do
  DDGI_PLOC_RADIUS=$R python bench.py --workload $WL --steps 30 --warmup 5 --no-cpu-baseline --no-alt > gpurun_out/${TAG}_r$R.json 2> gpurun_out/${TAG}_r$R.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_r$R.json"))
    r=d["roofline"]
    print("radius=$R ms_per_step=%.4f trace=%.4f conv=%.4f build=%.2f nodes/primary=%.2f nodes/shadow=%.2f kernels=%s" % (d["ms_per_step"], d["update_ms"]["trace"], d["update_ms"]["convolve"], d["config"]["accel_build_ms"], r["nodes_per_primary_ray"], r["nodes_per_shadow_ray"], [round(v,4) for v in r["kernels_ms_avg"].values()]))
except Exception as e:
    print("radius=$R failed", e); print(open("gpurun_out/${TAG}_r$R.err").read()[-2000:])
PY
done
