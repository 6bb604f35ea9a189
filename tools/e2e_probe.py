"""Where does the e2e step time go?  Host enqueue time vs device time of the C3 update loop (run on the GPU box)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import anerend_b200 as A
from anerend_b200 import layouts as L, scenes
scene = scenes.sponza_like(); nxz, ny, nr = 24, 12, 256
h = A.Ddgi(A.make_params(nxz, ny, nr, flags=L.FLAG_POINT_LIGHTS | L.FLAG_SKIP_DEAD_SHADOW_RAYS))
h.upload_scene(scene); h.set_scene_ubo(scene.ubo); h.build_accel()
h.set_timers(int(os.environ.get("DDGI_TIMERS", "0")))      # 0 none (production), 1 passes, 2 passes + kernels
h.set_pipeline(os.environ.get("DDGI_PIPELINE", "0") == "1")  # closest-hit trace of the next update under the tail of this one
rots = h.rotations(scenes.rotation_sequence(3, ny * 60)).reshape(60, ny, 16)
host = torch.empty((h.atlas_height, h.atlas_width, 4), dtype=torch.int16, pin_memory=True).numpy().view(np.uint16)
for mode in ("no_download", "async", "sync"):
    for k in range(5):
        h.update(-1, rots[k])
    h.sync()
    t0 = time.perf_counter(); tq = []
    for k in range(50):
        a = time.perf_counter()
        h.set_scene_ubo(scene.ubo); h.translate(); h.update(-1, rots[k])
        if mode == "async": h.download_atlas_async(host)
        elif mode == "sync": h.download_atlas(host)
        tq.append(time.perf_counter() - a)
    t1 = time.perf_counter()
    h.download_wait(); h.sync()
    t2 = time.perf_counter()
    t = h.timings()
    print(f"   last step: rt {t['rt_ms']:.4f} conv {t['convolve_ms']:.4f} kernels {[round(x, 4) for x in t['rt_kernel_ms']]}")
    print(f"{mode}: host enqueue {1e3 * (t1 - t0) / 50:.3f} ms/step (median call {1e3 * np.median(tq):.3f}), total {1e3 * (t2 - t0) / 50:.3f} ms/step")
