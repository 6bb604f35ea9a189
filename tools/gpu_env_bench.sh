#!/bin/bash
# The C3 bench (short form) under each "VAR=value[;VAR2=value2]" setting.  Usage (through gpurun): bash tools/gpu_env_bench.sh <tag> "<settings>" [workload]
TAG=${1:-env}; SETS=${2:-"X=0"}; WL=${3:-c3}
mkdir -p gpurun_out
for S in $SETS; do
  NAME=$(echo $S | tr ',=;' '___')
  env $(echo $S | tr ';' ' ') python bench.py --workload $WL --steps 30 --warmup 5 --no-alt --no-extras --no-cpu-baseline > gpurun_out/${TAG}_${WL}_$NAME.json 2> gpurun_out/${TAG}_${WL}_$NAME.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_${WL}_$NAME.json")); r=d["roofline"]
    print("$S ms_per_step=%.4f e2e_ms=%.4f build=%.2f nodes/shadow=%.2f kernels=%s conv=%.4f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["details"]["accel_build_ms"], r["nodes_per_shadow_ray"], [round(v,4) for v in r["kernels_ms_avg"].values()], d["update_ms"]["convolve"]), "prefiltered", r.get("shadow_rays_prefiltered"), "candidate", r.get("shadow_rays_ended_by_candidate"), "traced", r.get("shadow_rays_traced"))
except Exception as e:
    print("$S failed", e); print(open("gpurun_out/${TAG}_${WL}_$NAME.err").read()[-1500:])
PY
done
