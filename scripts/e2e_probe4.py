"""Outlier frames: is the GPU slow (per-stage CUDA events) or the host late?  Direct C-ABI, synchronous frames."""
import time
import numpy as np
import vulkanhybridraytracer_b200 as vk
from vulkanhybridraytracer_b200 import PLANE

W, H = 1920, 1080
sc = vk.HostScene(0)
ctx = vk.Context(0); ctx.upload_scene(sc)
ctx.set_camera(sc.camera(W, H))
ctx.set_profiling(1)
for refit in (0, 1):
    for chain in (2, 0):
        rows = []
        for k in range(240):
            t0 = time.perf_counter()
            if refit:
                ctx.set_lights(sc.lights); ctx.build_tlas(True)
            ctx.set_params(vk.make_params(W, H, frame=k, wavefront=chain))
            ctx.render(); t1 = time.perf_counter()
            ctx.sync(); t2 = time.perf_counter()
            c = ctx.counters()
            rows.append(((t2 - t0) * 1e3, (t1 - t0) * 1e3, c.ms_frame, *[c.ms_stage[i] for i in range(8)]))
        r = np.array(rows[40:])
        out = np.where(r[:, 0] > 1.3 * np.median(r[:, 0]))[0]
        print(f"refit={refit} chain={chain}: wall median {np.median(r[:,0]):.3f} mean {r[:,0].mean():.3f} | gpu ms_frame median {np.median(r[:,2]):.3f} mean {r[:,2].mean():.3f} | outliers {len(out)}/200")
        for i in out[:5]:
            print("    wall %.3f submit %.3f gpu_frame %.3f stages" % tuple(r[i, :3]), np.round(r[i, 3:], 3))
        i = int(np.argsort(r[:, 0])[len(r) // 2])
        print("    typical: wall %.3f submit %.3f gpu_frame %.3f stages" % tuple(r[i, :3]), np.round(r[i, 3:], 3))
