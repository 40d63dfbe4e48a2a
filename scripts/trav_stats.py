"""Development aid: traversal statistics (node-pair fetches, triangle tests) for the primary rays of the default scene."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vulkanhybridraytracer_b200 as vk
from vulkanhybridraytracer_b200 import types as T
W, H = 1920, 1080
sc = vk.HostScene(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ctx = vk.Context(0); ctx.set_builder(int(os.environ.get("BUILDER", "1"))); ctx.upload_scene(sc)
cam = sc.camera(W, H)
iv = np.array(cam.inv_view[:], np.float32).reshape(4, 4).T; ip = np.array(cam.inv_proj[:], np.float32).reshape(4, 4).T
ys, xs = np.mgrid[0:H:2, 0:W:2]
d = np.stack([(xs + 0.5) / W * 2 - 1, (ys + 0.5) / H * 2 - 1, np.ones_like(xs, float), np.ones_like(xs, float)], -1).astype(np.float32)
t = d @ ip.T; tn = t[..., :3] / np.linalg.norm(t[..., :3], axis=-1, keepdims=True)
dirs = (tn @ iv[:3, :3].T).reshape(-1, 3)
rays = np.zeros(len(dirs), T.ray_dtype)
rays["origin"] = iv[:3, 3] + dirs * 1e-2; rays["dir"] = dirs; rays["tmin"] = 0.001; rays["tmax"] = 1e4
st = ctx.trace_stats(rays)
hits = ctx.trace_rays(rays)
print("primary rays %d: steps mean %.1f p50 %d p99 %d max %d | tris mean %.1f p99 %d max %d" % (len(rays), st[:, 0].mean(), np.median(st[:, 0]), np.percentile(st[:, 0], 99), st[:, 0].max(), st[:, 1].mean(), np.percentile(st[:, 1], 99), st[:, 1].max()))
it = st[:, 0] + st[:, 1]
slow = np.argsort(-st[:, 2].astype(np.int64))[:8]
print("cycles: mean %.0f p50 %d p99 %d max %d | cycles per iteration (steps+tris): mean %.0f ; slowest rays: %s" % (st[:, 2].mean(), np.median(st[:, 2]), np.percentile(st[:, 2], 99), st[:, 2].max(), (st[:, 2] / np.maximum(it, 1)).mean(), [(int(st[k, 0]), int(st[k, 1]), int(st[k, 2])) for k in slow]))
heavy = it > 40
print("rays with > 40 iterations: %d, cycles/iteration %.0f" % (heavy.sum(), (st[heavy, 2] / it[heavy]).mean()))
for inst in range(len(sc.instances)):
    m = hits["instance"] == inst
    if m.any(): print("  hit instance %d (%d tris): rays %d steps mean %.1f tris mean %.1f" % (inst, sc.instances["index_count"][inst] // 3, m.sum(), st[m, 0].mean(), st[m, 1].mean()))
m = hits["instance"] < 0
print("  miss: rays %d steps mean %.1f tris mean %.1f" % (m.sum(), st[m, 0].mean(), st[m, 1].mean()))
for b in range(5):
    nodes, prims = ctx.bvh(b)
    print("  blas %d: prims %d nodes %d" % (b, len(prims), len(nodes)))
