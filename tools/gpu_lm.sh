#!/bin/bash
# Light-map check: parity tests of the maps + C3 full size, then the C3 bench under each map resolution (0 = maps off).
# Usage (through gpurun): bash tools/gpu_lm.sh <tag> "<resolutions>"
TAG=${1:-lm}; RES=${2:-"0 256 512 1024"}
mkdir -p gpurun_out
python -m pytest tests/test_light_maps.py tests/test_gpu_full_size.py -m gpu -x -q -k "light or config3" 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
for R in $RES; do
  DDGI_LM_RES=$R python bench.py --workload c3 --steps 30 --warmup 5 --no-alt --no-extras --no-cpu-baseline > gpurun_out/${TAG}_c3_r$R.json 2> gpurun_out/${TAG}_c3_r$R.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_c3_r$R.json")); r=d["roofline"]
    print("R=$R ms_per_step=%.4f e2e_ms=%.4f build=%.2f shadow/primary=%.3f nodes/shadow=%.2f kernels=%s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["details"]["accel_build_ms"], r["shadow_rays_per_primary"], r["nodes_per_shadow_ray"], [round(v,4) for v in r["kernels_ms_avg"].values()]), "prefiltered", r.get("shadow_rays_prefiltered"), "candidate", r.get("shadow_rays_ended_by_candidate"), "traced", r.get("shadow_rays_traced"))
except Exception as e:
    print("R=$R failed", e); print(open("gpurun_out/${TAG}_c3_r$R.err").read()[-1500:])
PY
done
