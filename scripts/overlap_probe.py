"""Where does the overlapped frame spend its time?  Per-stage CUDA-event times of the main stream and the chain's own
start-to-end time on its stream, for the chain schedules, with and without reflection paths (max_bounces = 0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vulkanhybridraytracer_b200 as vk

W, H = 1920, 1080
sc = vk.HostScene(0)
ctx = vk.Context(0); ctx.upload_scene(sc)
ctx.set_camera(sc.camera(W, H))
ctx.set_profiling(1)
for chain, mb in ((7, 8),):
    rows = []
    for k in range(24):
        ctx.flush_l2()
        ctx.set_params(vk.make_params(W, H, frame=3 + k, wavefront=chain, max_bounces=mb))
        ctx.render(); ctx.sync()
        c = ctx.counters()
        rows.append(list(c.ms_stage)[:4] + [c.ms_stage[7], c.ms_tlas, c.ms_frame])
    r = np.median(np.array(rows[4:]), axis=0)
    print(f"chain {chain} max_bounces {mb}: gbuffer {r[0]:.3f} shadow {r[1]:.3f} denoise {r[2]:.3f} resolve {r[3]:.3f} | chain stream {r[4]:.3f} ms, ends at {r[5]:.3f} | frame {r[6]:.3f}", flush=True)
