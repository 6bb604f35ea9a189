#!/bin/bash
# ncu --set full capture of the non-traversal kernels (k_shade, k_shade_prep, k_probe_convolve_fast, k_conv_weights) of one
# c3 update after warm-up, a details page for each, then a DDGI_TUNE sweep.
# Usage (through gpurun): bash tools/gpu_prof3.sh <tag> "<tune1> <tune2> ..."
TAG=${1:-prof3}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_probe_convolve_fast|k_conv_weights' -s 12 -c 4 -f -o gpurun_out/${TAG} \
    python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --no-alt > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
ncu -i gpurun_out/${TAG}.ncu-rep --page details > gpurun_out/${TAG}_details.txt 2>&1
for T in $2; do
  DDGI_TUNE=$T python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline --no-alt > gpurun_out/${TAG}_$T.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_$T.json"))
print("tune=$T ms_per_step=%.4f trace=%.4f kernels=%s" % (d["ms_per_step"], d["update_ms"]["trace"], [round(v,4) for v in d["roofline"]["kernels_ms_avg"].values()]))
PY
done
