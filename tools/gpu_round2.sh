#!/bin/bash
# Round-2 profile set on one GPU: the default bench line (as the driver runs it), the launch list of the same command, one
# `ncu --set full` capture of the five kernels of an update (after warm-up) with details / raw metrics / SASS-level source
# pages, and the other single-GPU workloads.  Usage (through gpurun): bash tools/gpu_round2.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/${TAG}_launches_c3.csv \
    python bench.py --workload c3 --steps 4 --warmup 3 --no-cpu-baseline --no-alt --no-extras > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_launches_c3.csv > gpurun_out/${TAG}_launches_c3_summary.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_shade|k_probe_convolve_fast' -s 20 -c 5 -f -o gpurun_out/${TAG}_kernels \
    python bench.py --workload c3 --steps 2 --warmup 5 --no-cpu-baseline --no-alt --no-extras > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_kernels.ncu-rep --page details > gpurun_out/${TAG}_ncu_kernels_details.txt 2>&1
ncu -i gpurun_out/${TAG}_kernels.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,smsp__thread_inst_executed_per_inst_executed.ratio,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active > gpurun_out/${TAG}_ncu_kernels_raw.csv 2>&1
for WL in c2 c4 c5; do
  ST=20; [ "$WL" = "c5" ] && ST=5
  python bench.py --workload $WL --steps $ST --warmup 3 --no-extras > gpurun_out/${TAG}_bench_${WL}.json 2> gpurun_out/${TAG}_bench_${WL}.err
done
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
python tools/build_time.py c2 c3 c5 > gpurun_out/${TAG}_build_time.txt 2>&1
python tools/two_level_bench.py instanced4096 c3 c5 > gpurun_out/${TAG}_two_level.jsonl 2> gpurun_out/${TAG}_two_level.err
python tools/two_level_stats.py c3 instanced4096 > gpurun_out/${TAG}_two_level_stats.txt 2>&1
python - <<PY
import json
for f in ("bench_default", "bench_c2", "bench_c4", "bench_c5", "bench_reference"):
    try:
        d = json.loads([l for l in open("gpurun_out/${TAG}_%s.json" % f) if l.startswith("{")][-1])
        print(f, "ms_per_step=%.4f value=%.1f e2e_ms=%s" % (d["ms_per_step"], d["value"], d["e2e"].get("ms_per_step")))
    except Exception as e:
        print(f, "failed", e)
PY
cat gpurun_out/${TAG}_launches_c3_summary.txt | head -12; cat gpurun_out/${TAG}_build_time.txt
