"""Development aid: per-stage times for different views / resolutions (where do the kernel floors come from?)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vulkanhybridraytracer_b200 as vk
sc = vk.HostScene(0)
ctx = vk.Context(0); ctx.upload_scene(sc); ctx.set_profiling(True)
def run(name, W, H, pos, yaw, pitch, chain=0):
    sc.set_camera(pos, yaw, pitch); cam = sc.camera(W, H); ctx.set_camera(cam)
    acc = np.zeros(8); n = 10
    for f in range(n + 3):
        ctx.set_params(vk.make_params(W, H, frame=f + 1, wavefront=chain)); ctx.flush_l2(); ctx.render(); c = ctx.counters()
        if f >= 3: acc += np.array(list(c.ms_stage))
    acc /= n
    print("%-28s %4dx%-4d gbuffer %.3f shadow %.3f denoise %.3f resolve %.3f | rays p %d s %d r %d ss %d" % (name, W, H, acc[0], acc[1], acc[2], acc[3], c.rays_primary, c.rays_shadow, c.rays_reflect, c.rays_secondary_shadow))
for W, H in ((1920, 1080), (960, 540), (480, 270)):
    run("default", W, H, (0, 5, 10), -90, 0)
    run("sky only (look up)", W, H, (0, 5, 10), -90, 80)
    run("floor only (look down)", W, H, (8, 6, 4), -90, -80)
    run("lucy close-up", W, H, (0, 1.3, 2.2), -90, 0)
    run("helmet close-up", W, H, (-5, 1, -2.8), -90, 0)
