"""Development aid: traversal statistics of (approximate) shadow rays seeded from the GPU G-buffer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vulkanhybridraytracer_b200 as vk
from vulkanhybridraytracer_b200 import types as T, PLANE, STAGE
W, H = 1920, 1080
sc = vk.HostScene(0)
ctx = vk.Context(0); ctx.set_builder(int(os.environ.get("BUILDER", "1"))); ctx.upload_scene(sc)
cam = sc.camera(W, H); ctx.set_camera(cam); ctx.set_params(vk.make_params(W, H)); ctx.run_stage(STAGE.GBUFFER)
pos = ctx.readback(PLANE.POSITION).astype(np.float32); nrm = ctx.readback(PLANE.NORMAL).astype(np.float32); ids = ctx.readback(PLANE.HIT_IDS)[..., 0]
P = pos[..., :3].reshape(-1, 3); N = nrm[..., :3].reshape(-1, 3) * 2 - 1; N /= np.maximum(np.linalg.norm(N, axis=1, keepdims=True), 1e-9)
idf = ids.reshape(-1)
for l in range(3):
    L = sc.lights["position"][l][:3] - P; dist = np.linalg.norm(L, axis=1); D = L / dist[:, None]
    m = ((N * D).sum(1) > 0) & (idf >= 0)
    rays = np.zeros(int(m.sum()), T.ray_dtype)
    rays["origin"] = P[m] + D[m] * 0.1; rays["dir"] = D[m]; rays["tmin"] = 0.001; rays["tmax"] = dist[m] + 1
    st = ctx.trace_stats(rays, any_hit=True)
    print("light %d: rays %d steps mean %.1f p50 %d p99 %d p99.9 %d max %d | tris mean %.1f p99 %d max %d" % (l, len(rays), st[:, 0].mean(), np.median(st[:, 0]), np.percentile(st[:, 0], 99), np.percentile(st[:, 0], 99.9), st[:, 0].max(), st[:, 1].mean(), np.percentile(st[:, 1], 99), st[:, 1].max()))
    src = idf[m]
    for inst in range(6):
        k = src == inst
        if k.any(): print("    from instance %d: rays %7d steps mean %6.1f max %4d tris mean %5.1f max %4d" % (inst, k.sum(), st[k, 0].mean(), st[k, 0].max(), st[k, 1].mean(), st[k, 1].max()))
