"""bench.py's reference arm (the CPU implementation of the path, timed on the host cores) prints ONE JSON line carrying the contract's
keys; under a multi-rank launch only rank 0 runs.  (The CUDA arm needs a device; its line is checked on the B200 by the driver.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *flags):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--workload", "c2", *flags],
                          capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    out = _run({})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "ddgi_probe_update_primary_mrays_per_s" and j["unit"] == "Mrays/s"
    assert j["higher_is_better"] is True and j["n_gpus"] == 1 and j["steps"] == 1 and j["value"] > 0 and j["ms_per_step"] > 0
    assert j["config"]["workload"] == "c2" and j["data"] == "synthetic" and j["dtype"] == "f32" and j["vs_baseline"] is None
    cb, e2e = j["cpu_baseline"], j["e2e"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "Y-layer" in cb["sample"]
    assert e2e == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_without_work():
    out = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29571"}, "--gpus", "2")
    assert out.returncode == 0 and out.stdout.strip() == "", out.stdout + out.stderr[-2000:]
