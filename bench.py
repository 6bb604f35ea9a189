#!/usr/bin/env python
"""bench.py — DDGI probe-update throughput (BASELINE.json metric) on N B200s of one node.

A "step" is one full-grid DDGI probe update (translate if needed -> trace+shade all layers -> convolve) of one
synthetic workload:
    c3 (default at N=1): Sponza-scale mesh (~262k tris, 64 materials, 8 lights), 24x12x24 probes x 256 rays + shadow rays
    c5 (default at N>1): large volume (~2M tris), 64x32x64 probes x 256 rays, probe layers dealt to the ranks + atlas all-gather
    c4                 : camera-scroll sequence: c3 + 16 moving rigid nodes; every step = new camera UBO (x: -8 -> +8 over 120
                         steps at 30 Hz sim time), BVH refit of the moved nodes, probe-grid translation, full-grid update
    c2                 : Cornell box (~10k tris), 16x8x16 probes x 128 rays
`value` = primary probe rays of all ranks / device time (CUDA events on the handle's stream, max over ranks), inputs
resident in HBM.  `e2e` = the same through the public API with host buffers: per step the scene UBO + rotation matrices
go host->device and the atlas comes back to pinned host memory, wall-clocked.  `--impl reference` times the CPU oracle
(the reference has no CPU path; the oracle is its shader transliteration) on the box's host cores instead.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ddgi_probe_update_primary_mrays_per_s"
UNIT = "Mrays/s"


def build_workload(name):
    from anerend_b200 import scenes
    if name == "c2":
        return scenes.cornell_box(), (16, 8, 128), 2
    if name == "c3":
        return scenes.sponza_like(), (24, 12, 256), 3
    if name == "c4":
        return scenes.sponza_like(seed=4, movers=16), (24, 12, 256), 4
    if name == "c5":
        return scenes.large_volume(), (64, 32, 256), 5
    raise SystemExit(f"unknown workload {name}")


def workload_desc(name, scene, grid, world):
    nxz, ny, nr = grid
    return {"workload": name, "scene": scene.name, "triangles": int(scene.num_triangles),
            "probes": f"{nxz}x{ny}x{nxz}", "rays_per_probe": nr, "primary_rays_per_update": nxz * nxz * ny * nr,
            "lights": int(len(scene.lights)), "point_lights_E1": True, "multi_bounce": True, "dead_shadow_rays_skipped": True,
            "mode": "jacobi full-grid update", "parallelism": f"probe layers dealt to {world} ranks" if world > 1 else "single GPU",
            # timing rule of the CUDA arm (the CPU arm has no device cache to flush; the key is kept so both arms print the same config)
            "l2": "working set larger than L2, no flush" if name == "c5" else "flushed between timed iterations (256 MiB write)"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz, self._stop_evt, self.index = [], set(), None, threading.Event(), index

    def run(self):
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                     0x80: "hw_power_brake_slowdown"}
            while not self._stop_evt.is_set():
                self.samples.append(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM))
                try:
                    r = N.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.01)
        except Exception as e:  # NVML unavailable: report that instead of guessing
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def algorithmic_bytes(stats, rays):
    """80 B per BVH8 node fetched + 48 B per triangle tested (closest-hit and any-hit), per hit
    12 B indices + 288 B vertices + 64 B material + 256 B of previous-atlas texels, per ray 8 B RGBA16F radiance out."""
    nodes = stats["primary_nodes_visited"] + stats["shadow_nodes_visited"]
    tris = stats["primary_tris_tested"] + stats["shadow_tris_tested"]
    return 80 * nodes + 48 * tris + stats["primary_hits"] * (12 + 288 + 64 + 256) + rays * 8


def run_oracle_sample(scene, grid, layers, threads=0, point_lights=True):
    """Times the CPU oracle on `layers` of the workload; returns (rays, seconds, threads)."""
    from anerend_b200 import layouts as L, make_params, scenes
    from oracle import binding as ob
    nxz, ny, nr = grid
    o = ob.Oracle(make_params(nxz, ny, nr, flags=(L.FLAG_POINT_LIGHTS | L.FLAG_SKIP_DEAD_SHADOW_RAYS) if point_lights else L.FLAG_SKIP_DEAD_SHADOW_RAYS))
    nthreads = o.set_threads(threads)
    o.upload_scene(scene)
    o.set_scene_ubo(scene.ubo)
    o.build_accel()
    rots = o.rotations(scenes.rotation_sequence(scene.seed, ny))
    t0 = time.perf_counter()
    for layer in layers:
        o.update(layer, rots[layer])
    dt = time.perf_counter() - t0
    o.close()
    return len(layers) * nxz * nxz * nr, dt, nthreads


def vulkan_reference_status():
    """north_star: 'the reference Vulkan path, reported additionally if the host driver exposes ray-tracing pipelines'.  It never can be
    run from this repo (the reference's host code is Win32-only and its shaders need VK_KHR_ray_tracing_pipeline); this records why."""
    try:
        import ctypes.util
        import shutil
        loader = ctypes.util.find_library("vulkan")
        info = shutil.which("vulkaninfo")
        if not loader and not info:
            return "n/a: no Vulkan loader (libvulkan) and no vulkaninfo on this box; B200 exposes no VK_KHR_ray_tracing_pipeline"
        return ("n/a: a Vulkan loader is present (%s) but the reference's host code is Win32-only and B200 exposes no "
                "VK_KHR_ray_tracing_pipeline" % (loader or info))
    except Exception as ex:                                         # noqa: BLE001
        return "n/a: probe failed (%s)" % str(ex)[:80]


def run_config1(dev, cpu_baseline=True):
    """BASELINE config 1 — probe_conv.comp alone: 8 x 8 x 8 probes x 64 rays on synthetic radiance U[0, 4)^3.  The CUDA convolution runs
    through _upload_ray_radiance + _convolve_layer (radiance resident, one launch set per layer, all 8 layers = one pass); next to it the
    single-thread CPU transliteration the config names (the oracle; cpu_baseline leg).  Parity of the two: tests/test_config1.py."""
    import torch
    import anerend_b200 as A
    from anerend_b200 import make_params, scenes
    nxz, ny, nr = 8, 8, 64
    ubo = scenes.empty_scene(grid=(nxz, ny, nr)).ubo
    rng = np.random.RandomState(1)
    irr = np.ones((ny, 8 * nxz, 8 * nxz, 4), np.float16)
    irr[..., :3] = rng.uniform(0, 4, irr[..., :3].shape)
    hc = A.Ddgi(make_params(nxz, ny, nr))
    out = {"workload": "c1", "probes": "8x8x8", "rays_per_probe": nr, "texels": nxz * nxz * ny * 64}
    try:
        hc.set_scene_ubo(ubo)
        rot = hc.make_rotation(*scenes.rotation_sequence(1, 1)[0])
        for layer in range(ny):
            hc.upload_ray_radiance(layer, irr[layer], rot)
        stream = torch.cuda.ExternalStream(hc.stream(), device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(23):
            if it == 3:
                e0.record(stream)
            for layer in range(ny):
                hc.convolve_layer(layer)
        e1.record(stream)
        hc.sync()
        ms = e0.elapsed_time(e1) / 20.0
        out["gpu"] = {"ms_per_pass": ms, "mprobes_per_s": nxz * nxz * ny / ms / 1e3, "launches_per_pass": 2 * ny,
                      "note": "8 per-layer calls of a 64-probe layer each: launch-latency bound at this size"}
    finally:
        hc.close()
    if cpu_baseline:
        from oracle import binding as ob
        o = ob.Oracle(make_params(nxz, ny, nr))
        o.set_threads(1)
        o.set_scene_ubo(ubo)
        for layer in range(ny):
            o.upload_ray_radiance(layer, irr[layer], rot)
        best = float("inf")
        for it in range(5):
            t0 = time.perf_counter()
            for layer in range(ny):
                o.convolve_layer(layer)
            best = min(best, time.perf_counter() - t0)
        o.close()
        out["cpu_single_thread"] = {"ms_per_pass": best * 1e3, "ns_per_texel": best * 1e9 / (nxz * nxz * ny * 64), "kind": "port", "cores": 1}
    return out


def sample_layers(name, ny):
    """The Y-layers one CPU step covers: the whole grid where that takes about a second (c2, c3, c4), every 8th layer of c5 (an
    equal-weight sample over the height of the volume).  Used by BOTH CPU figures (--impl reference and cpu_baseline)."""
    return list(range(ny)) if name != "c5" else list(range(0, ny, 8))


def reference_arm(args, rank, world):
    """The reference's CPU implementation of the path = the oracle (the path itself — GLSL + Vulkan RT — cannot be built here;
    oracle/_ref only holds the reference's struct headers and glm::rotate).  Rank 0 alone runs; each step is one pass over
    sample_layers() of the workload — the same measurement as the `cpu_baseline` leg of the CUDA arm."""
    if rank != 0:
        return
    name = args.workload or ("c3" if world == 1 else "c5")
    scene, grid, _ = build_workload(name)
    from anerend_b200 import layouts as L, make_params, scenes
    from oracle import binding as ob
    nxz, ny, nr = grid
    o = ob.Oracle(make_params(nxz, ny, nr, flags=L.FLAG_POINT_LIGHTS | L.FLAG_SKIP_DEAD_SHADOW_RAYS))     # the same work as the CUDA arm
    threads = o.set_threads(0)
    o.upload_scene(scene)
    o.set_scene_ubo(scene.ubo)
    o.build_accel()
    rots = o.rotations(scenes.rotation_sequence(scene.seed, ny))
    layers = sample_layers(name, ny)
    per_step = len(layers) * nxz * nxz * nr
    steps, warmup = min(args.steps, 20), min(args.warmup, 3)       # bounded: a step is 0.3 - 5 s of CPU work
    for w in range(warmup):
        for layer in layers:
            o.update(layer, rots[layer])
    t0 = time.perf_counter()
    for k in range(steps):
        for layer in layers:
            o.update(layer, rots[layer])
    dt = time.perf_counter() - t0
    value = per_step * steps / dt / 1e6
    sample = (f"{steps} steps x {len(layers)} of {ny} Y-layers ({per_step} primary rays + their shadow rays per step) of {name}, "
              f"multithreaded C++ oracle (shader transliteration)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / steps * 1e3 * (ny / len(layers)), "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_desc(name, scene, grid, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "steps_timed": steps, "note": "ms_per_step is scaled to a full-grid update where a step covers a sample of the layers"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, "c2", "c3", "c4", "c5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fast-math", action="store_true", help="time the FAST_MATH build as the headline (default: exact build)")
    ap.add_argument("--no-extras", action="store_true", help="skip the atlas-consumer / ShadowRT / SpecularGIRT figures and the C5 1-GPU base")
    ap.add_argument("--no-alt", action="store_true", help="skip the second-build and 1-GPU strong-scaling-base measurements")
    ap.add_argument("--deal", default=None, choices=["stripes", "blocks", "measured", "static"],
                    help="N>1, fused exchange: every rank takes the probe rows z = rank mod N of every layer (default; needs nXZ "
                         "divisible by N), or whole layers dealt by measured per-layer cost, or by the static mirrored/zigzag rule")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="N>1: atlas exchange fused into the convolution over NVLink peer memory (default) or NCCL all-gather")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import anerend_b200 as A
    from anerend_b200 import layouts as L, scenes, sharding

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    name = args.workload or ("c3" if world == 1 else "c5")
    scene, grid, seed = build_workload(name)
    nxz, ny, nr = grid
    # multi-GPU: interleaved layer sets (balanced) when nY divides evenly, contiguous slabs otherwise
    mirrored = world > 1 and ny % (2 * world) == 0
    interleaved = world > 1 and (mirrored or ny % world == 0)
    if args.deal is None:                      # peer push: row stripes; NCCL: the static mirrored slabs (the measured collective baseline)
        args.deal = "stripes" if args.exchange == "fused" else "static"
    blocks = args.deal == "blocks"              # contiguous probe rows per rank instead of interleaved ones (same code path otherwise)
    if blocks:
        args.deal = "stripes"
    stripes = world > 1 and args.deal == "stripes" and nxz % world == 0
    measured_deal = world > 1 and args.exchange == "fused" and (args.deal == "measured" or (args.deal == "stripes" and not stripes))
    slab = (0, ny) if (interleaved or measured_deal or stripes) else sharding.slab_for_rank(ny, rank, world)
    my_layers = (list(range(ny)) if stripes else sharding.mirrored_slabs_for_rank(ny, rank, world) if mirrored else
                 sharding.layers_for_rank(ny, rank, world) if interleaved else list(range(slab[0], slab[1])))
    # point lights (extension E1) + dead-shadow-ray elimination (radiance bit-identical, tests/test_oracle_kat.py);
    # the default kernels are the exact build (atlas bit-identical to the oracle); FAST_MATH is timed as a second figure
    flags = L.FLAG_POINT_LIGHTS | L.FLAG_SKIP_DEAD_SHADOW_RAYS
    if args.fast_math:
        flags |= L.FLAG_FAST_MATH
    dev = torch.device("cuda", local_rank)
    total_steps = args.warmup + args.steps
    seq = scenes.rotation_sequence(seed, ny * total_steps)
    all_rays = nxz * nxz * ny * nr

    if name == "c4":
        c4_ubos = [scene.with_camera(np.array([-8.0 + 16.0 * k / 120.0, 5.0, 0.0])) for k in range(120)]
        c4_rend = [scene.animated_renderables(k / 30.0) for k in range(120)]

    fused = world > 1 and args.exchange == "fused"

    def make_handle(hflags, hslab, connect=False, striped=True):
        hh = A.Ddgi(A.make_params(nxz, ny, nr, flags=hflags, slab=hslab, device=local_rank))
        if world == 1 and os.environ.get("DDGI_BENCH_STRIPE"):      # profiling aid: ONE rank's share of an R-GPU run on one GPU, no exchange ("r,R")
            r_, R_ = (int(v) for v in os.environ["DDGI_BENCH_STRIPE"].split(","))
            hh.set_row_stripe(r_, R_)
        if stripes and striped:
            if blocks:
                hh.set_row_block(rank * (nxz // world), nxz // world)
            else:
                hh.set_row_stripe(rank, world)
        hh.upload_scene(scene)
        hh.set_scene_ubo(scene.ubo)
        hh.build_accel()
        if connect:
            connect_handle(hh)
        return hh

    def close_handle(hh, connected=False):
        if connected:
            hh.sync()
            dist.barrier()
            hh.exchange_disconnect()
            dist.barrier()
        hh.close()

    def connect_handle(hh):                   # fused exchange: map every peer's atlas (cudaIpc over NVLink)
        blobs = [None] * world
        dist.all_gather_object(blobs, hh.exchange_export())
        hh.exchange_connect(blobs, world, rank)
        dist.barrier()

    h = make_handle(flags, slab, connect=False)
    accel_ms = h.timings()["accel_ms"]
    info = h.bvh_info()
    rots = h.rotations(seq).reshape(total_steps, ny, 16)
    if os.environ.get("DDGI_BENCH_FIXED_ROT"):        # experiments only: the same rotations in every step
        rots[:] = rots[0]
    balanced = None
    if measured_deal:
        # measured-cost deal of the layers (sharding.balanced_layers): every rank times one update of each layer alone on
        # its own GPU (replicated scene: same numbers everywhere), rank 0's figures are broadcast so all ranks derive
        # the same deal.  The peer-store exchange takes arbitrary layer sets; the NCCL gathers keep the static deals.
        costs = [sharding.measure_layer_costs(h, rots[0])]
        dist.broadcast_object_list(costs, src=0)
        deal = sharding.balanced_layers(costs[0], world)
        my_layers = deal[rank]
        loads = [sum(costs[0][y] for y in d) for d in deal]
        balanced = {"layer_cost_ms_min": min(costs[0]), "layer_cost_ms_max": max(costs[0]), "rank_load_ms_min": min(loads),
                    "rank_load_ms_max": max(loads), "layers_per_rank": [len(d) for d in deal]}
        h.reset()
    if fused:
        connect_handle(h)
    my_rays = nxz * nxz * len(my_layers) * nr // (world if stripes else 1)
    nbytes = h.atlas_device_ptr()[1]
    host_atlas = torch.empty((h.atlas_height, h.atlas_width, 4), dtype=torch.int16, pin_memory=True)   # pinned result buffer
    host_np = host_atlas.numpy().view(np.uint16)

    # working set vs L2 (126 MB): flush between timed iterations when everything would stay cached
    working_set = info.num_nodes * 80 + info.num_triangles * (48 + 80) + nbytes + my_rays * (8 + 16 + 48 + 4)
    flush = working_set < 256e6
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if flush else None

    def timed(hh, steps, warmup, do_gather, multi):
        """`steps` full-grid updates of handle hh, CUDA events on the handle's stream; returns device ms and per-pass ms."""
        stream = torch.cuda.ExternalStream(hh.stream(), device=dev)
        ptr, _ = hh.atlas_device_ptr()
        atlas_t = torch.as_tensor(sharding.DeviceArray(ptr, (hh.atlas_height, hh.atlas_width, 4)), device=dev)

        def step(k):
            if name == "c4":                     # camera scroll + moving nodes (host renderable buffer -> refit -> translate)
                hh.set_scene_ubo(c4_ubos[k % 120])
                hh.refit_accel(c4_rend[k % 120])
            else:
                hh.set_scene_ubo(scene.ubo)
            hh.translate()
            if do_gather and fused:           # the convolution stores into every peer's atlas; barriers inside the call
                hh.update_layers(my_layers, rots[k])
            elif do_gather and stripes:       # row stripes + one NCCL all-gather of the packed stripe rows
                hh.update_layers(my_layers, rots[k])
                with torch.cuda.stream(stream):
                    sharding.all_gather_atlas_stripes(atlas_t, nxz, ny, rank, world)
            elif do_gather and interleaved:
                hh.update_layers(my_layers, rots[k])
                with torch.cuda.stream(stream):
                    if mirrored:
                        sharding.all_gather_atlas_mirrored(atlas_t, nxz, ny, rank, world)
                    else:
                        sharding.all_gather_atlas_layers(atlas_t, nxz, ny, rank, world)
            else:
                hh.update(-1, rots[k])
                if do_gather:
                    with torch.cuda.stream(stream):
                        sharding.all_gather_atlas(atlas_t, nxz, ny, rank, world)

        # Two passes over the same steps.  (A) the timed region: no timing events inside the library (set_timers(0)), no host
        # synchronisation between the steps — K steps between two synchronisations, every step bracketed by its own pair of events on the
        # handle's stream (the L2 flush sits outside the pairs).  (B) the same steps again with events around every pass and kernel
        # (set_timers(2)) and one synchronisation per step: where rt / convolve / per-kernel times come from.  The events of (B) are not
        # free (C3: +26 us per update), which is why (A) is the headline.
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        hh.set_timers(0)
        # N > 1 (C5: no L2 flush, the steps follow each other on the stream): consecutive updates pipelined, as in the e2e loop — the K
        # steps still lie between two synchronisations; one GPU keeps every step on its own between two events with the flush in between
        hh.set_pipeline(multi and not flush)
        for k in range(warmup):
            step(k)
        hh.sync()
        if multi:
            dist.barrier()
        torch.cuda.synchronize()
        for k in range(steps):
            if flush:
                with torch.cuda.stream(stream):
                    flush_buf.fill_(k & 0xFF)
            ev[k][0].record(stream)
            step(warmup + k)
            ev[k][1].record(stream)
        if multi and fused:
            # the last update's push into the peers runs on the exchange stream: the timed region ends when it has landed
            tail0, tail1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tail0.record(stream)
            hh.exchange_fence()
            tail1.record(stream)
            ev.append((tail0, tail1))
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        if multi:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        # (B) per-pass / per-kernel times
        hh.set_pipeline(False)
        hh.set_timers(2)
        rt, cv, nl = [], [], 0
        hh.kernel_ms = []                        # per step: (closest-hit trace, shade_prep, any-hit trace, shade) in ms
        evb = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for k in range(steps):
            if flush:
                with torch.cuda.stream(stream):
                    flush_buf.fill_(k & 0xFF)
            evb[k][0].record(stream)
            step(warmup + k)
            evb[k][1].record(stream)
            t = hh.timings()                     # synchronises the handle's stream
            rt.append(t["rt_ms"])
            cv.append(t["convolve_ms"])
            hh.kernel_ms.append(t["rt_kernel_ms"])
            nl += t["kernel_launches"]          # our kernels only (the NCCL collective and torch's pack/unpack copies are not counted)
        if multi and fused:
            hh.exchange_fence()
        hh.sync()
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
        hh.ms_with_kernel_timers = sum(a.elapsed_time(b) for a, b in evb) / steps
        hh.set_timers(0)
        return ms, rt, cv, nl, step

    do_gather = world > 1
    sampler = ClockSampler(local_rank)
    sampler.start()
    dev_ms, rt_ms, conv_ms, launches, step = timed(h, args.steps, args.warmup, do_gather, world > 1)
    clocks = sampler.stop()
    per_rank = None
    if world > 1:                                  # own trace / convolve(+exchange) ms of every rank: shows the load balance
        mine = torch.tensor([float(np.mean(rt_ms)), float(np.mean(conv_ms))], device=dev)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"trace_ms": [round(float(t[0]), 4) for t in allr], "convolve_exchange_ms": [round(float(t[1]), 4) for t in allr]}
    ms_per_step = dev_ms / args.steps
    value = all_rays / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: host buffers in, atlas back to pinned host memory, wall clock ---------------------------------
    e2e_steps = max(3, min(args.steps, 50))
    # consecutive updates are pipelined here (anerend_cuda_ddgi_set_pipeline): ray generation and the closest-hit trace of update k+1
    # read nothing update k writes, so they run on their own stream under the tail of update k's kernels; the timed region of
    # `value` above keeps every step on its own between two events (active on one GPU only: a connected peer exchange switches it off)
    h.set_pipeline(True)
    for k in range(2):                 # untimed: the first pipelined updates
        step(k)
    h.sync()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    # result read-back: the whole atlas on one GPU; at N>1 every rank reads back the layers it computed (together the
    # ranks return the whole atlas once per step)
    if stripes:
        own_np = host_np.reshape(-1)[:host_np.size // world].reshape(ny, nxz // world, 10, h.atlas_width, 4)
    else:
        own_np = host_np.reshape(-1)[:len(my_layers) * h.atlas_width * h.atlas_width * 4].reshape(len(my_layers), h.atlas_width, h.atlas_width, 4)
    for k in range(e2e_steps):
        step(args.warmup + (k % args.steps))
        if stripes:
            # own stripe on the read-back stream: overlaps the next step's tracing; the next own convolution waits for it on the device
            h.download_atlas_stripe_async(own_np)
        elif world > 1:
            h.download_atlas_layers(my_layers, own_np)
        else:
            # D2H of every step's result into pinned memory on the library's copy stream: it overlaps the next step's
            # tracing, and the next convolution waits for it on the device (anerend_cuda_ddgi_download_atlas_async)
            h.download_atlas_async(host_np)
    if world == 1 or stripes:
        h.download_wait()
    h.sync()
    e2e_s = time.perf_counter() - t0
    h.set_pipeline(False)
    if world > 1:
        tt = torch.tensor([e2e_s], device=f"cuda:{local_rank}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = all_rays * e2e_steps / e2e_s / 1e6
    h2d_bytes = 568 + 64 * ny                          # scene UBO + one rotation matrix per layer (they travel as kernel arguments; the
                                                       # direction table is made on the device from them)
    d2h_bytes = nbytes if world == 1 else int(own_np.nbytes)

    # ---- second figure: the other arithmetic build (FAST_MATH <-> exact), same workload, fewer steps ----------
    alt = None
    if not args.no_alt:
        ha = make_handle(flags ^ L.FLAG_FAST_MATH, slab, connect=fused)
        asteps = max(3, min(args.steps, 20))
        ams, art, acv, _, _ = timed(ha, asteps, 3, do_gather, world > 1)
        alt = {"build": "exact" if args.fast_math else "fast_math", "ms_per_step": ams / asteps,
               "value": all_rays / (ams / asteps * 1e-3) / 1e6, "unit": UNIT, "trace_ms": float(np.mean(art)),
               "convolve_ms": float(np.mean(acv)), "steps": asteps}
        close_handle(ha, fused)

    # ---- the faithful subset: what rchit:85-94 shades — the sun only, no point lights (extension E1 off) — and the same with the light
    # maps off (every shadow ray traced), so the headline can be read as reference work + E1 + what the maps save
    subsets = None
    if world == 1 and not args.no_alt:
        subsets = {}
        for key, fl in (("sun_only_faithful", flags & ~L.FLAG_POINT_LIGHTS), ("sun_only_no_light_maps", (flags & ~L.FLAG_POINT_LIGHTS) | L.FLAG_NO_LIGHT_MAPS),
                        ("point_lights_no_light_maps", flags | L.FLAG_NO_LIGHT_MAPS)):
            hs = make_handle(fl, slab)
            ssteps = max(3, min(args.steps, 20))
            sms_, srt, scv, _, _ = timed(hs, ssteps, 3, False, False)
            subsets[key] = {"ms_per_step": sms_ / ssteps, "value": all_rays / (sms_ / ssteps * 1e-3) / 1e6, "unit": UNIT, "trace_ms": float(np.mean(srt)),
                            "kernels_ms_avg": [round(float(v), 4) for v in np.mean(np.asarray(hs.kernel_ms, dtype=np.float64), axis=0)], "steps": ssteps}
            close_handle(hs)

    # ---- strong-scaling base: the SAME workload on one GPU (rank 0 alone), so the N-GPU figure has its own denominator
    base1 = None
    verified = None
    if world > 1 and not args.no_alt:
        # exchange check: two updates from a cleared atlas on all ranks must give rank 0 the atlas ONE GPU computes
        # (jacobi updates do not depend on how the layers are dealt; the exact build is bit-reproducible)
        h.reset()
        for k in range(2):
            step(k)
        h.sync()
        dist.barrier()
        if rank == 0:
            multi = np.empty_like(host_np)
            h.download_atlas(multi)
            hb = make_handle(flags, (0, ny), striped=False)
            bsteps = 3
            bms, _, _, _, bstep = timed(hb, bsteps, 1, False, False)
            base1 = {"n_gpus": 1, "ms_per_step": bms / bsteps, "value": all_rays / (bms / bsteps * 1e-3) / 1e6, "unit": UNIT, "steps": bsteps}
            # the same workload end to end on one GPU (host buffers in, whole atlas back to pinned memory every step, wall clock): the
            # denominator of the N-GPU e2e figure
            hb.sync()
            tb = time.perf_counter()
            for k in range(bsteps):
                bstep(k)
                hb.download_atlas_async(host_np)
            hb.download_wait()
            hb.sync()
            base1["e2e_ms_per_step"] = (time.perf_counter() - tb) / bsteps * 1e3
            base1["e2e_value"] = all_rays / (base1["e2e_ms_per_step"] * 1e-3) / 1e6
            if name != "c4":
                hb.reset()
                for k in range(2):
                    hb.set_scene_ubo(scene.ubo)
                    hb.translate()
                    hb.update(-1, rots[k])
                single = np.empty_like(host_np)
                hb.download_atlas(single)
                verified = {"updates": 2, "texels": int(single.size // 4), "mismatching_texels": int((multi != single).any(axis=-1).sum())}
            hb.close()
        dist.barrier()

    # ---- atlas consumer: calcIndirectDiffuseLight for a 1920x1080 synthetic G-buffer against the atlas
    consumer = None
    if world == 1 and not args.no_extras:
        npts = 1920 * 1080
        gen = torch.Generator(device=dev); gen.manual_seed(11)
        pts = torch.zeros((npts, 12), dtype=torch.float32, device=dev)
        half = 0.5 * nxz
        pts[:, 0:3] = torch.rand((npts, 3), generator=gen, device=dev) * torch.tensor([2 * half, 2.0 * ny, 2 * half], device=dev) \
            + torch.tensor([float(scene.ubo["camera_pos"][0][0]) - half, 0.0, float(scene.ubo["camera_pos"][0][2]) - half], device=dev)
        nrm = torch.randn((npts, 3), generator=gen, device=dev)
        pts[:, 3:6] = nrm / nrm.norm(dim=1, keepdim=True)
        pts[:, 6:9] = torch.rand((npts, 3), generator=gen, device=dev) * 0.9 + 0.05
        pts[:, 10] = torch.rand(npts, generator=gen, device=dev) * 0.9 + 0.1
        outc = torch.empty((npts, 4), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        cstream = torch.cuda.ExternalStream(h.stream(), device=dev)
        view = [float(v) for v in scene.ubo["camera_pos"][0][:3]]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(13):
            if it == 3:
                e0.record(cstream)
            h.indirect_diffuse_device(pts.data_ptr(), npts, view, True, outc.data_ptr())
        e1.record(cstream)
        h.sync()
        cms = e0.elapsed_time(e1) / 10.0
        cbytes = npts * (48 + 16 + 32 * 8)
        consumer = {"points": npts, "sampler": "clamp_to_edge (deferred lighting)", "ms": cms, "mpoints_per_s": npts / cms / 1e3,
                    "algorithmic_bytes": cbytes, "achieved_gbs": cbytes / (cms * 1e-3) / 1e9, "finite": bool(torch.isfinite(outc).all().item())}

    # ---- ShadowRT on the same BVH: one any-hit ray per pixel of a 1920x1080 depth buffer (top-down view)
    shadow_rt = None
    if world == 1 and not args.no_extras:
        from anerend_b200 import layouts as _L
        sw, sh = 1920, 1080
        u = scene.ubo.copy()
        m = np.zeros((4, 4)); m[0, 0], m[2, 1], m[1, 2], m[3, 3] = 0.5 * nxz, 0.5 * nxz, 2.0 * ny, 1.0     # clip -> world, w = 1
        m[0, 3], m[2, 3] = float(u["camera_pos"][0][0]), float(u["camera_pos"][0][2])
        u["inv_view_proj"][0] = m.T.astype(np.float32).ravel()
        u["screen_width"], u["screen_height"] = sw, sh
        u["flags"] |= _L.UBO_RT_SHADOWS_FLAG
        h.set_scene_ubo(u)
        gen = torch.Generator(device=dev); gen.manual_seed(12)
        ddepth = torch.rand((sh, sw), generator=gen, device=dev) * 0.96 + 0.02
        dvis = torch.empty((sh, sw), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        sstream = torch.cuda.ExternalStream(h.stream(), device=dev)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(13):
            if it == 3:
                s0.record(sstream)
            h.shadow_rt_device(ddepth.data_ptr(), sw, sh, dvis.data_ptr())
        s1.record(sstream)
        h.sync()
        sms = s0.elapsed_time(s1) / 10.0
        shadow_rt = {"pixels": sw * sh, "ms": sms, "mrays_per_s": sw * sh / sms / 1e3, "lit_fraction": float(dvis.mean().item())}
        # SpecularGIRT: same camera and depth, random G-buffer normals (incoherent reflections); wavefront kernels in external-ray mode
        u["flags"] |= _L.UBO_SPECULAR_GI_FLAG
        h.set_scene_ubo(u)
        dgeom = torch.randn((sh, sw, 4), generator=gen, device=dev)
        dspec = torch.empty((sh, sw, 4), dtype=torch.int16, device=dev)
        torch.cuda.synchronize()
        for it in range(5):
            if it == 2:
                s0.record(sstream)
            h.specular_gi_device(ddepth.data_ptr(), dgeom.data_ptr(), sw, sh, dspec.data_ptr())
        s1.record(sstream)
        h.sync()
        shadow_rt["specular_gi_ms"] = s0.elapsed_time(s1) / 3.0
        h.set_scene_ubo(scene.ubo)

    # ---- BASELINE config 1 (the reference's CPU-runnable case): an extra figure, never allowed to take the headline line down with it
    config1 = None
    if world == 1 and args.workload is None and not args.no_extras:
        try:
            config1 = run_config1(dev, cpu_baseline=not args.no_cpu_baseline)
        except Exception as ex:                                     # noqa: BLE001
            config1 = {"workload": "c1", "error": str(ex)[:200]}

    # ---- base of the multi-GPU series: the N>1 runs of this script time C5 (the configuration BASELINE.json shards over 2/4/8
    # GPUs), so the default N=1 run also reports C5 on one GPU — the denominator of that series — next to its own C3 headline
    series_base = None
    if world == 1 and args.workload is None and not args.no_extras:
        import subprocess
        try:
            outp = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", "c5", "--steps", "5", "--warmup", "3", "--no-alt",
                                   "--no-cpu-baseline", "--no-extras"], capture_output=True, text=True, timeout=600)
            d5 = json.loads([ln for ln in outp.stdout.splitlines() if ln.startswith("{")][-1])
            series_base = {"workload": "c5", "n_gpus": 1, "value": d5["value"], "unit": UNIT, "ms_per_step": d5["ms_per_step"], "steps": d5["steps"],
                           "probes": d5["config"]["probes"], "triangles": d5["config"]["triangles"]}
        except Exception as e:      # the headline line must not depend on this extra
            series_base = {"workload": "c5", "error": f"{type(e).__name__}: {e}"[:200]}

    # ---- the other single-GPU BASELINE configurations next to the C3 headline: C2 (Cornell box) and C4 (camera scroll: translation +
    # refit of 16 movers + dynamic light maps every step), each a short run of this script
    other_workloads = None
    if world == 1 and args.workload is None and not args.no_extras:
        import subprocess
        other_workloads = {}
        for wl in ("c2", "c4"):
            try:
                outp = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", wl, "--steps", "20", "--warmup", "3", "--no-alt",
                                       "--no-cpu-baseline", "--no-extras"], capture_output=True, text=True, timeout=600)
                dw = json.loads([ln for ln in outp.stdout.splitlines() if ln.startswith("{")][-1])
                other_workloads[wl] = {"value": dw["value"], "unit": UNIT, "ms_per_step": dw["ms_per_step"], "e2e_ms_per_step": dw["e2e"]["ms_per_step"],
                                       "steps": dw["steps"], "probes": dw["config"]["probes"], "triangles": dw["config"]["triangles"]}
            except Exception as e:      # the headline line must not depend on this extra
                other_workloads[wl] = {"error": f"{type(e).__name__}: {e}"[:200]}

    # ---- two-level acceleration structure (FLAG_TWO_LEVEL): memory, build, per-frame cost of moving instances and
    # full-grid update against the one-level structure on a scene of 4096 instances of 5 meshes, and on the C3 scene
    two_level = None
    if world == 1 and args.workload is None and not args.no_extras:
        import subprocess
        try:
            outp = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "two_level_bench.py"), "instanced4096", "c3"],
                                  capture_output=True, text=True, timeout=600)
            two_level = [json.loads(ln) for ln in outp.stdout.splitlines() if ln.startswith("{")]
        except Exception as e:      # the headline line must not depend on this extra
            two_level = {"error": f"{type(e).__name__}: {e}"[:200]}

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel: counted, untimed run of the same update ----------
        hc = make_handle(flags | L.FLAG_COUNTERS, slab)
        if interleaved or measured_deal or stripes:
            hc.update_layers(my_layers, rots[args.warmup])
        else:
            hc.update(-1, rots[args.warmup])
        stats = hc.stats()
        hc.close()
        peak, peak_src = peaks()
        # whole IrradianceProbeRT pass (4 launches): SURVEY.md §8(d) bytes / CUDA-event time of the pass
        abytes = algorithmic_bytes(stats, my_rays)
        rt_avg = float(np.mean(rt_ms))
        kms = np.mean(np.asarray(h.kernel_ms, dtype=np.float64), axis=0)      # live CUDA events around each launch
        knames = ["k_trace<ANY=0> (closest hit)", "k_shade_prep", "k_trace<ANY=1> (any-hit shadow rays)", "k_shade"]
        # dominant kernel = the any-hit traversal of the shadow-ray queues: 80 B per BVH8 node fetched + 48 B per
        # triangle tested + per shadow ray 4 B queue entry, 16 B origin (surface record) and 4 B visibility word
        kbytes = [80 * stats["primary_nodes_visited"] + 48 * stats["primary_tris_tested"] + my_rays * (16 + 16),
                  None,
                  80 * stats["shadow_nodes_visited"] + 48 * stats["shadow_tris_tested"] + stats["shadow_rays"] * 24,
                  None]
        dom = int(np.argmax(kms))
        if kbytes[dom] is None:
            dom = 2
        achieved = kbytes[dom] / (kms[dom] * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(name, {}).get("k_trace_any_dram_bytes_per_launch" if dom == 2 else
                                                                "k_trace_closest_dram_bytes_per_launch")
        roofline = {"bound": "hbm", "kernel": knames[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": kbytes[dom], "kernel_ms_avg": float(kms[dom]),
                    "kernel_share_of_step": float(kms[dom]) / ms_per_step,
                    "kernels_ms_avg": {knames[i]: float(kms[i]) for i in range(4)},
                    "rt_pass": {"launches": 4, "algorithmic_bytes": abytes, "ms_avg": rt_avg,
                                "achieved": abytes / (rt_avg * 1e-3) / 1e9, "frac": abytes / (rt_avg * 1e-3) / 1e9 / peak,
                                "share_of_step": rt_avg / ms_per_step},
                    "frac_meaning": "ALGORITHMIC bytes (80 B per BVH8 node visited + 48 B per triangle tested + per-ray records) / HBM peak, the figure "
                                    "north_star asks for; it is NOT DRAM utilisation: `traffic` (ncu dram__bytes) is a few per cent of it, the working set "
                                    "is L2-resident and the kernel is instruction-issue bound (profiles/)",
                    "nodes_per_primary_ray": stats["primary_nodes_visited"] / max(1, stats["primary_rays"]),
                    "tris_per_primary_ray": stats["primary_tris_tested"] / max(1, stats["primary_rays"]),
                    "nodes_per_shadow_ray": stats["shadow_nodes_visited"] / max(1, stats["shadow_rays"]),
                    "tris_per_shadow_ray": stats["shadow_tris_tested"] / max(1, stats["shadow_rays"]),
                    "shadow_rays_per_primary": stats["shadow_rays"] / max(1, stats["primary_rays"]),
                    "hit_fraction": stats["primary_hits"] / max(1, stats["primary_rays"]),
                    # light maps (csrc/ddgi_kernels.cuh): shadow rays answered without a traversal / ended by the candidate triangle
                    "shadow_rays_needed": stats["shadow_rays"] + stats["shadow_rays_prefiltered"],
                    "shadow_rays_prefiltered": stats["shadow_rays_prefiltered"], "shadow_rays_traced": stats["shadow_rays"],
                    "shadow_rays_ended_by_candidate": stats["shadow_rays_candidate"]}
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            layers = sample_layers(name, ny)
            rays, secs, threads = run_oracle_sample(scene, grid, layers)
            cpu_baseline = {"value": rays / secs / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
                            "sample": f"{len(layers)} of {ny} Y-layers of {name} ({rays} primary rays + shadow rays), "
                                      f"{secs:.2f} s, multithreaded C++ oracle (shader transliteration)"}
        cfg = workload_desc(name, scene, grid, world)
        deal_desc = None
        if balanced is not None:
            deal_desc = f"probe layers dealt to {world} ranks by measured per-layer cost (LPT)"
        if stripes:
            deal_desc = (f"row blocks: rank r updates the probe rows z in [r*{nxz // world}, (r+1)*{nxz // world}) of every layer" if blocks else
                         f"row stripes: rank r updates the probe rows z = r mod {world} of every layer")
        details = {}
        assert flush == (name != "c5"), "config.l2 (workload_desc) must state what the timed loop does"
        details.update({"bvh8_nodes": int(info.num_nodes), "accel_build_ms": accel_ms,
                    "atlas_exchange": ("peer push: after the convolution a copy kernel stores this rank's atlas rows into all peers' atlases "
                                       "over NVLink (cudaIpc-mapped), device-side barrier before and after; no collective" if fused else
                                       "nccl all_gather of the packed stripe rows on the handle's stream" if stripes else
                                       "2 in-place nccl all_gathers of mirrored slabs on the handle's stream" if mirrored else
                                       "nccl all_gather of interleaved layers (pack, gather, scatter) on the handle's stream" if interleaved
                                       else "nccl all_gather of contiguous slabs, in place") if world > 1 else "none"})
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "details": dict(details, deal=deal_desc), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                        "ms_per_step": e2e_s / e2e_steps * 1e3, "steps": e2e_steps,
                        "pipelined": "closest-hit trace of update k+1 under the tail of update k (set_pipeline)"},
                "pipelined_value": bool(world > 1 and not flush),
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "update_ms": {"trace": rt_avg, "convolve": float(np.mean(conv_ms)), "full_grid": ms_per_step,
                              "full_grid_with_kernel_timers": float(h.ms_with_kernel_timers)},
                "timers": ("value / ms_per_step / e2e: no timing events inside the library (anerend_cuda_ddgi_set_timers(h, 0)), the steps bracketed by "
                           "bench.py's own events; update_ms.trace / convolve and roofline.kernels_ms_avg: the same steps again with events around every "
                           "pass and kernel (level 2), which costs update_ms.full_grid_with_kernel_timers - full_grid per update"),
                "arithmetic": "fast_math (1e-4 relative per update)" if args.fast_math else "exact (atlas bit-identical to the oracle)",
                "other_build": alt, "subsets": subsets, "strong_scaling_base": base1, "exchange_check_vs_1gpu": verified, "layer_deal": balanced, "per_rank_ms": per_rank, "atlas_consumer": consumer, "shadow_rt": shadow_rt, "multi_gpu_series_base": series_base,
                "other_workloads": other_workloads, "two_level_structure": two_level, "config1_conv_only": config1, "vulkan_reference": vulkan_reference_status()}
    close_handle(h, fused)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
