"""Does the present copy overlap the next frame?  Direct C-ABI loop with / without per-frame uniform uploads."""
import time
import numpy as np
import torch
import vulkanhybridraytracer_b200 as vk
from vulkanhybridraytracer_b200 import PLANE

W, H = 1920, 1080
sc = vk.HostScene(0)
ctx = vk.Context(0); ctx.upload_scene(sc)
cam = sc.camera(W, H); ctx.set_camera(cam)
pinned = [torch.empty((H, W, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
n = pinned[0].numel()
f = 0
def frame(upload, refit):
    global f
    if upload:
        ctx.set_lights(sc.lights)
    if refit:
        ctx.build_tlas(True)
    ctx.set_params(vk.make_params(W, H, frame=f, wavefront=2)); f += 1
    ctx.render()
for k in range(5):
    frame(True, True); ctx.readback(PLANE.FINAL_SRGB8)
for upload, refit, present in ((0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 0, 2)):
    ctx.sync(); ctx.readback_wait(0)
    t0 = time.perf_counter()
    for k in range(60):
        ctx.set_output_buffer(k & 1)
        frame(upload, refit)
        if present == 1:
            ctx.readback_async(PLANE.FINAL_SRGB8, pinned[k & 1].data_ptr(), n)
            ctx.readback_wait(1)
        elif present == 2:
            ctx.L.vkp_readback(ctx.h, PLANE.FINAL_SRGB8, pinned[0].data_ptr(), n)
        elif k % 4 == 3:
            ctx.sync()
    ctx.readback_wait(0); ctx.sync()
    dt = (time.perf_counter() - t0) / 60 * 1e3
    print(f"upload={upload} refit={refit} present={present}: {dt:.3f} ms/frame")
