"""Small frames through every schedule and the build, for compute-sanitizer (memcheck / racecheck / synccheck).
`quick` as first argument: the default schedule and schedule 2 only, three frames each (reset frame + two accumulation frames, where k_shadow's threads
apply the filter themselves), no group."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vulkanhybridraytracer_b200 as vk
from vulkanhybridraytracer_b200 import PLANE

sc = vk.HostScene(0, detail=(4, 3, 4, 6))
ctx = vk.Context(0); ctx.upload_scene(sc)
w, h = 160, 96
ctx.set_camera(sc.camera(w, h))
ref = None
QUICK = len(sys.argv) > 1 and sys.argv[1] == "quick"
for chain in ((7, 2) if QUICK else (0, 1, 2, 3, 4, 5, 6, 7, 8)):
    ctx.set_params(vk.make_params(w, h)); ctx.reset_history()
    for f in range(3 if QUICK else 2):
        ctx.set_params(vk.make_params(w, h, frame=f, wavefront=chain, faithful_rng=0, shadow_samples=2 if chain % 2 else 1)); ctx.render()
    img = ctx.readback(PLANE.FINAL_F32).copy()
    print("chain", chain, "checksum", float(img.sum()), flush=True)
if QUICK:
    ctx.close(); print("SANITIZE_PROBE_DONE"); sys.exit(0)
ctx.set_params(vk.make_params(w, h, frame=0, wavefront=7, denoiser=1)); ctx.reset_history(); ctx.render(); ctx.sync()
ctx.set_params(vk.make_params(w, h, frame=0, wavefront=7, shadow_samples=8, faithful_rng=0)); ctx.reset_history(); ctx.render(); ctx.sync()      # persistent k_shadow
ctx.set_params(vk.make_params(w, h, mode=vk.MODE_RAYTRACING, spp=2, shadow_samples=9, faithful_rng=0, frame=0)); ctx.reset_history(); ctx.render(); ctx.sync()
ctx.set_params(vk.make_params(w, h, mode=vk.MODE_DEFERRED)); ctx.render(); ctx.sync()
g = vk.Group([0, 0]); g.upload_scene(sc); g.set_camera(sc.camera(w, h))
for f in range(2):
    g.set_params(vk.make_params(w, h, frame=f, wavefront=7)); g.render()
print("group checksum", float(g.readback(PLANE.FINAL_F32).sum()), flush=True)
g.close(); ctx.close()
print("SANITIZE_PROBE_DONE")
