"""Development aid: render one Hybrid frame on the GPU through the C-ABI, compare every stage with the oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vulkanhybridraytracer_b200 as vk
from vulkanhybridraytracer_b200 import PLANE
from oracle.oracle_py import Oracle

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (640, 360)
sc = vk.HostScene(0)
t = time.time(); ctx = vk.Context(0); ctx.upload_scene(sc); ctx.sync(); print("gpu scene upload+build %.3fs" % (time.time() - t))
cam = sc.camera(W, H); p = vk.make_params(W, H)
ctx.set_camera(cam); ctx.set_params(p); ctx.set_profiling(True)
for k in range(3):
    ctx.reset_history(); ctx.render(); c = ctx.counters()
print("gpu frame ms", c.ms_frame, "stages", list(c.ms_stage)[:4], "rays", c.rays_primary, c.rays_shadow, c.rays_reflect, c.rays_secondary_shadow)
o = Oracle(sc); t = time.time(); fr = o.hybrid_frame(cam, p); print("oracle frame %.2fs" % (time.time() - t), list(o.counters.rays))
g = fr["g"]
def cmp(name, a, b):
    a = np.asarray(a); b = np.asarray(b)
    if a.dtype == np.float16: a = a.view(np.uint16); b = b.view(np.uint16)
    neq = (a != b)
    if neq.ndim == 3: neq = neq.any(-1)
    print("%-10s mismatching pixels: %d / %d" % (name, int(neq.sum()), neq.size))
cmp("hit_ids", ctx.readback(PLANE.HIT_IDS), g.hit_ids)
cmp("position", ctx.readback(PLANE.POSITION), g.position)
cmp("normal", ctx.readback(PLANE.NORMAL), g.normal)
cmp("albedo", ctx.readback(PLANE.ALBEDO), g.albedo)
cmp("motion", ctx.readback(PLANE.MOTION), g.motion)
cmp("material", ctx.readback(PLANE.MATERIAL), g.material)
cmp("emissive", ctx.readback(PLANE.EMISSIVE), g.emissive)
for l in range(len(sc.lights)):
    cmp("shadow%d" % l, ctx.readback(PLANE.SHADOW0 + l), fr["shadow"][l])
    cmp("denoised%d" % l, ctx.readback(PLANE.DENOISED0 + l), fr["denoised"][l])
fin = ctx.readback(PLANE.FINAL_F32)
d = np.abs(fin[..., :3] - fr["final"][..., :3])
print("final: MAE %.3e max %.3e  bit-identical pixels %d / %d" % (d.mean(), d.max(), int((fin.view(np.uint32) == fr["final"].view(np.uint32)).all(-1).sum()), W * H))
srgb = ctx.readback(PLANE.FINAL_SRGB8)
cmp("srgb8", srgb, o.post(p, fr["final"]))
os.makedirs("gpurun_out", exist_ok=True)
import cv2
cv2.imwrite("gpurun_out/first_light_gpu.png", srgb[:, :, [2, 1, 0]])
