#!/bin/bash
# C5 strong-scaling bench on N GPUs of one box: fused exchange (with the 1-GPU base + exchange check) and NCCL all-gather.
# Usage (through gpurun --gpus N): bash tools/gpu_scale.sh <tag> <N> [steps]
TAG=${1:-scale}; N=${2:-2}; ST=${3:-10}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $ST --warmup 3 \
  > gpurun_out/${TAG}_n${N}_fused.json 2> gpurun_out/${TAG}_n${N}_fused.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps $ST --warmup 3 --exchange nccl --no-alt \
  > gpurun_out/${TAG}_n${N}_nccl.json 2> gpurun_out/${TAG}_n${N}_nccl.err
python - <<PY
import json
for x in ("fused", "nccl"):
    try:
        d = json.loads([l for l in open("gpurun_out/${TAG}_n${N}_%s.json" % x) if l.startswith("{")][-1])
        b = d.get("strong_scaling_base")
        print(x, "N=%d ms_per_step=%.4f value=%.1f trace=%.4f conv=%.4f e2e_ms=%.3f base1=%s check=%s other=%s" % (d["n_gpus"], d["ms_per_step"], d["value"], d["update_ms"]["trace"], d["update_ms"]["convolve"], d["e2e"]["ms_per_step"], b and round(b["ms_per_step"], 3), d.get("exchange_check_vs_1gpu"), d["other_build"] and round(d["other_build"]["ms_per_step"], 4)))
    except Exception as e:
        print(x, "failed", e); print(open("gpurun_out/${TAG}_n${N}_%s.err" % x).read()[-1500:])
PY
