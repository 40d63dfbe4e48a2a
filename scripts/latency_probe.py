"""Development aid: single-warp traversal latency vs throughput (device timer around vkp_trace_rays)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vulkanhybridraytracer_b200 as vk
from vulkanhybridraytracer_b200 import types as T
sc = vk.HostScene(0)
ctx = vk.Context(0); ctx.upload_scene(sc)
def mk(n, target, jitter):
    rng = np.random.default_rng(0)
    rays = np.zeros(n, T.ray_dtype)
    o = np.array([0, 5, 10], np.float32)
    t = np.array(target, np.float32) + rng.normal(scale=jitter, size=(n, 3)).astype(np.float32)
    d = t - o; d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["origin"] = o; rays["dir"] = d; rays["tmin"] = 0.001; rays["tmax"] = 1e4
    return rays
def timeit(rays, any_hit=False, reps=20):
    ctx.trace_rays(rays, any_hit)
    best = 1e9
    for _ in range(reps):
        ctx.timer_begin(); ctx.trace_rays(rays, any_hit); best = min(best, ctx.timer_end())
    return best
for name, target, jit in (("sky (miss)", [0, 50, -5], 0.01), ("floor", [3, 0, 0], 0.01), ("lucy", [0, 1.2, 0], 0.02), ("helmet", [-5, 1, -5], 0.02), ("mirror sphere", [5, 1, -5], 0.02)):
    r32 = mk(32, target, jit); st = ctx.trace_stats(r32)
    t32 = timeit(r32); t1m = timeit(mk(1 << 20, target, jit), reps=5)
    print("%-14s steps %.1f tris %.1f | 1 warp %.1f us | 1M rays %.1f us (%.2f Grays/s)" % (name, st[:, 0].mean(), st[:, 1].mean(), t32 * 1e3, t1m * 1e3, (1 << 20) / t1m / 1e6))
