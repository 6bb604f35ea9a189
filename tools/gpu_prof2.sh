#!/bin/bash
# ncu --set full capture of the IrradianceProbeRT / Convolve kernels of one c3 update (one launch each, after warm-up).
# Usage (through gpurun): bash tools/gpu_prof2.sh <tag> [workload]
TAG=${1:-prof}; WL=${2:-c3}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_shade|k_probe_convolve_fast' -s 20 -c 5 -f -o gpurun_out/${TAG} \
    python bench.py --workload $WL --steps 2 --warmup 5 --no-cpu-baseline --no-alt > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/${TAG}.ncu-rep
