#!/bin/bash
# One GPU-box round trip: parity tests, bench line, ncu launch list of the same bench command, and one
# `ncu --set full` capture of the two traversal kernels (after warm-up).
# Usage (through gpurun): bash tools/gpu_round.sh <tag> [workload] [skip-tests]
TAG=${1:-run}; WL=${2:-c3}; SKIPT=${3:-0}
mkdir -p gpurun_out
if [ "$SKIPT" = "0" ]; then
  python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
fi
python bench.py --workload $WL --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_${WL}.json 2> gpurun_out/${TAG}_bench_${WL}.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_${WL}.csv \
    python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --no-alt > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_launches_${WL}.csv > gpurun_out/${TAG}_launches_${WL}_summary.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_trace' -s 6 -c 2 -f -o gpurun_out/${TAG}_trace \
    python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --no-alt > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_trace.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,smsp__thread_inst_executed_per_inst_executed.ratio,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active > gpurun_out/${TAG}_trace_raw.csv 2>&1
tail -3 gpurun_out/${TAG}_tests.log; cat gpurun_out/${TAG}_bench_${WL}.json; tail -5 gpurun_out/${TAG}_bench_${WL}.err; cat gpurun_out/${TAG}_launches_${WL}_summary.txt; cat gpurun_out/${TAG}_trace_raw.csv | cut -c1-600
