"""Shared helpers of the test-suite (test infrastructure)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import vulkanhybridraytracer_b200 as vk  # noqa: E402
from vulkanhybridraytracer_b200 import types as T  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_DETAIL = (3, 2, 3, 4)
GOLDEN_W, GOLDEN_H = 64, 36


def random_rays(scene, n, seed=0, tmax=200.0):
    """Rays aimed at the scene from a shell around it (a mix of hits and misses), plus some starting inside geometry."""
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, T.ray_dtype)
    org = rng.normal(size=(n, 3)); org /= np.linalg.norm(org, axis=1, keepdims=True)
    org = org * rng.uniform(3.0, 25.0, size=(n, 1)) + np.array([0.0, 3.0, -5.0])
    org[:, 1] = np.abs(org[:, 1]) + 0.05
    target = rng.uniform([-8, 0, -12], [8, 6, 2], size=(n, 3))
    d = target - org
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    k = n // 10          # some axis-aligned directions (zero components) and un-normalised ones
    d[:k] = np.eye(3)[rng.integers(0, 3, size=k)] * rng.choice([-1.0, 1.0], size=(k, 1))
    d[k:2 * k] *= rng.uniform(0.2, 5.0, size=(k, 1))
    rays["origin"] = org.astype(np.float32); rays["dir"] = d.astype(np.float32)
    rays["tmin"] = 0.001; rays["tmax"] = tmax
    return rays


def golden_case(name):
    """Deterministic tiny renders of the default scene at reduced mesh detail; returns {array name: array}."""
    from oracle.oracle_py import Oracle
    sc = vk.HostScene(0, detail=GOLDEN_DETAIL)
    o = Oracle(sc)
    W, H = GOLDEN_W, GOLDEN_H
    cam = sc.camera(W, H)
    out = {}
    if name in ("hybrid_f0", "hybrid_f1"):
        f0 = o.hybrid_frame(cam, vk.make_params(W, H, frame=0))
        fr = f0
        if name == "hybrid_f1":
            fr = o.hybrid_frame(cam, vk.make_params(W, H, frame=1), history=f0["denoised"])
        g = fr["g"]
        out = {"position": g.position.view(np.uint16), "normal": g.normal.view(np.uint16), "albedo": g.albedo, "motion": g.motion.view(np.uint16),
               "material": g.material, "emissive": g.emissive, "hit_ids": g.hit_ids, "shadow": fr["shadow"], "denoised": fr["denoised"],
               "final_bits": fr["final"].view(np.uint32)}
    elif name == "raytrace":
        a = o.raytrace(cam, vk.make_params(W, H, mode=vk.MODE_RAYTRACING, spp=2, shadow_samples=2, faithful_rng=0))
        out = {"final_bits": a.view(np.uint32)}
    elif name == "deferred":
        g = o.gbuffer(cam, vk.make_params(W, H))
        d = o.deferred(cam, vk.make_params(W, H, mode=vk.MODE_DEFERRED), g)
        out = {"final_bits": d.view(np.uint32)}
    else:
        raise KeyError(name)
    o.close(); sc.close()
    return out


def upload_gbuffer(ctx, g):
    from vulkanhybridraytracer_b200 import PLANE
    ctx.upload_plane(PLANE.POSITION, g.position); ctx.upload_plane(PLANE.NORMAL, g.normal); ctx.upload_plane(PLANE.ALBEDO, g.albedo)
    ctx.upload_plane(PLANE.MOTION, g.motion); ctx.upload_plane(PLANE.MATERIAL, g.material); ctx.upload_plane(PLANE.EMISSIVE, g.emissive)


def bits(a):
    a = np.asarray(a)
    if a.dtype == np.float16:
        return a.view(np.uint16)
    if a.dtype == np.float32:
        return a.view(np.uint32)
    return a
