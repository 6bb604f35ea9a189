#!/bin/bash
# ncu --set full capture of one kernel (regex) of the c3 bench; brings the .ncu-rep back under gpurun_out/.
# Usage (through gpurun): bash tools/gpu_prof.sh <tag> <kernel-regex> [skip] [workload]
TAG=${1:-prof}; PAT=${2:-k_trace_shadow}; SKIP=${3:-4}; WL=${4:-c3}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$PAT -s $SKIP -c 1 -f -o gpurun_out/${TAG} \
    python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/${TAG}.ncu-rep
