"""Which thread sets the tail of k_shadow in the slow frames?"""
import time
import numpy as np
import vulkanhybridraytracer_b200 as vk
from vulkanhybridraytracer_b200 import PLANE, STAGE

W, H = 1920, 1080
sc = vk.HostScene(0)
ctx = vk.Context(0); ctx.upload_scene(sc)
ctx.set_camera(sc.camera(W, H))
rows = []
for k in range(120):
    ctx.set_params(vk.make_params(W, H, frame=k, wavefront=2))
    ctx.set_profiling(1); ctx.render(); ctx.sync(); c = ctx.counters()
    ms_shadow = c.ms_stage[STAGE.SHADOW]
    ctx.set_profiling(2); ctx.run_stage(STAGE.SHADOW); ctx.sync(); c2 = ctx.counters()
    rows.append((k, ms_shadow, c2.worst_thread_nodes, c2.worst_thread_tag, c2.node_fetches))
r = np.array(rows, dtype=np.float64)
print("corr(ms_shadow, worst_thread_nodes) = %.3f" % np.corrcoef(r[:, 1], r[:, 2])[0, 1])
for k, ms, wn, tag, nf in rows:
    tag = int(tag); pix = tag & ((1 << 23) - 1)
    if ms > 0.6 or k < 12:
        print(f"frame {k:3d} k_shadow {ms:.3f} ms  worst thread: {int(wn):6d} node fetches, light {(tag >> 23) & 31}, pixel ({pix % W},{pix // W}); total fetches {int(nf)}")
ids = ctx.readback(PLANE.HIT_IDS)
pos = ctx.readback(PLANE.POSITION).astype(np.float32)
for k, ms, wn, tag, nf in rows:
    if ms > 0.6:
        pix = int(tag) & ((1 << 23) - 1); x, y = pix % W, pix // W
        print("   pixel", (x, y), "hit ids", ids[y, x], "pos", pos[y, x])
