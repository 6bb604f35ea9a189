#!/bin/bash
# One GPU-box round trip: parity tests, a short bench line, and an ncu launch list of the same bench command.
# Usage (through gpurun): bash tools/gpu_check.sh <tag> [workload]
TAG=${1:-run}; WL=${2:-c3}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
python bench.py --workload $WL --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_${WL}.json 2> gpurun_out/${TAG}_bench_${WL}.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_${WL}.csv \
    python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_launches_${WL}.csv > gpurun_out/${TAG}_launches_${WL}_summary.txt 2>&1
tail -3 gpurun_out/${TAG}_tests.log; cat gpurun_out/${TAG}_bench_${WL}.json; tail -5 gpurun_out/${TAG}_bench_${WL}.err; cat gpurun_out/${TAG}_launches_${WL}_summary.txt
