#!/bin/bash
# Sweep one environment knob over the bench.  Usage (through gpurun): bash tools/gpu_env_sweep.sh <tag> <VAR> "<v1> <v2> ..." [workload]
TAG=${1:-env}; VAR=$2; VALS=$3; WL=${4:-c3}
mkdir -p gpurun_out
for V in $VALS; do
  env $VAR=$V python bench.py --workload $WL --steps 30 --warmup 5 --no-cpu-baseline --no-alt > gpurun_out/${TAG}_$V.json 2>gpurun_out/${TAG}_$V.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_$V.json"))
    print("$VAR=$V ms_per_step=%.4f trace=%.4f conv=%.4f kernels=%s" % (d["ms_per_step"], d["update_ms"]["trace"], d["update_ms"]["convolve"], [round(v,4) for v in d["roofline"]["kernels_ms_avg"].values()]))
except Exception as e:
    print("$VAR=$V failed", e); print(open("gpurun_out/${TAG}_$V.err").read()[-1500:])
PY
done
