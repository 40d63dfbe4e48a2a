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


def make_test_glb(path, seed=0):
    """A small self-contained .glb for the real-asset path: a UV sphere split into two primitives (textured
    upper half with base-colour / normal / emissive / metallic-roughness PNGs embedded in the binary chunk, plain lower
    half) under a rotated + scaled root node, and a child node holding a tilted quad.  Every primitive has normals."""
    import io
    import json
    import struct
    from PIL import Image

    rng = np.random.default_rng(seed)
    n_lon, n_lat = 24, 12
    lat = np.linspace(0, np.pi, n_lat + 1); lon = np.linspace(0, 2 * np.pi, n_lon + 1)
    la, lo = np.meshgrid(lat, lon, indexing="ij")
    pos = np.stack([np.sin(la) * np.cos(lo), np.cos(la), np.sin(la) * np.sin(lo)], -1).reshape(-1, 3).astype(np.float32)
    nrm = pos.copy()
    uv = np.stack([lo / (2 * np.pi), la / np.pi], -1).reshape(-1, 2).astype(np.float32)

    def band(r0, r1):
        tri = []
        for r in range(r0, r1):
            for c in range(n_lon):
                a, b = r * (n_lon + 1) + c, (r + 1) * (n_lon + 1) + c
                if r > 0:
                    tri += [a, b, a + 1]
                if r < n_lat - 1:
                    tri += [a + 1, b, b + 1]
        return np.array(tri, np.uint32)
    upper, lower = band(0, n_lat // 2), band(n_lat // 2, n_lat)
    quad_p = np.array([[-1.5, 0, -1.5], [1.5, 0, -1.5], [1.5, 0, 1.5], [-1.5, 0, 1.5]], np.float32)
    quad_n = np.array([[0, 1, 0]] * 4, np.float32)
    quad_uv = np.array([[0, 0], [2, 0], [2, 2], [0, 2]], np.float32)            # REPEAT addressing
    quad_i = np.array([0, 2, 1, 0, 3, 2], np.uint16)

    chunks, views = [], []

    def add(arr):
        off = sum(len(c) for c in chunks)
        b = np.ascontiguousarray(arr).tobytes()
        chunks.append(b + b"\0" * (-len(b) % 4))
        views.append({"buffer": 0, "byteOffset": off, "byteLength": len(b)})
        return len(views) - 1

    def png(a):
        buf = io.BytesIO(); Image.fromarray(a).save(buf, "PNG"); return np.frombuffer(buf.getvalue(), np.uint8)
    th, tw = 24, 48
    yy, xx = np.mgrid[0:th, 0:tw]
    base = np.stack([64 + 191 * ((xx // 6 + yy // 6) % 2), 255 - 4 * xx, 40 + 8 * yy], -1).astype(np.uint8)
    bump = rng.normal(0, 0.25, size=(th, tw, 2))
    nz = np.sqrt(np.clip(1 - (bump ** 2).sum(-1), 0.2, 1))
    normal = np.clip(np.dstack([bump, nz]) * 127.5 + 127.5, 0, 255).astype(np.uint8)
    emissive = np.zeros((th, tw, 3), np.uint8); emissive[::5, ::7] = [255, 120, 30]
    mr = np.stack([np.zeros_like(xx), 40 + 4 * xx, 255 - 9 * yy], -1).astype(np.uint8)     # G = roughness, B = metallic
    acc = [
        {"bufferView": add(pos), "componentType": 5126, "count": len(pos), "type": "VEC3", "min": pos.min(0).tolist(), "max": pos.max(0).tolist()},
        {"bufferView": add(nrm), "componentType": 5126, "count": len(nrm), "type": "VEC3"},
        {"bufferView": add(uv), "componentType": 5126, "count": len(uv), "type": "VEC2"},
        {"bufferView": add(upper), "componentType": 5125, "count": len(upper), "type": "SCALAR"},
        {"bufferView": add(lower), "componentType": 5125, "count": len(lower), "type": "SCALAR"},
        {"bufferView": add(quad_p), "componentType": 5126, "count": 4, "type": "VEC3"},
        {"bufferView": add(quad_n), "componentType": 5126, "count": 4, "type": "VEC3"},
        {"bufferView": add(quad_uv), "componentType": 5126, "count": 4, "type": "VEC2"},
        {"bufferView": add(quad_i), "componentType": 5123, "count": 6, "type": "SCALAR"},
    ]
    images = [{"bufferView": add(png(t)), "mimeType": "image/png"} for t in (base, normal, emissive, mr)]
    s = float(np.sin(np.radians(20) / 2)); c = float(np.cos(np.radians(20) / 2))
    doc = {
        "asset": {"version": "2.0", "generator": "tests/util.py"},
        "scenes": [{"nodes": [0]}],
        "nodes": [{"translation": [0.0, 0.6, 0.0], "rotation": [0.0, 0.0, s, c], "scale": [1.4, 1.4, 1.4], "mesh": 0, "children": [1]},
                  {"translation": [0.0, -0.9, 0.0], "rotation": [s, 0.0, 0.0, c], "mesh": 1}],
        "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "NORMAL": 1, "TEXCOORD_0": 2}, "indices": 3, "material": 0},
                                   {"attributes": {"POSITION": 0, "NORMAL": 1, "TEXCOORD_0": 2}, "indices": 4, "material": 1}]},
                   {"primitives": [{"attributes": {"POSITION": 5, "NORMAL": 6, "TEXCOORD_0": 7}, "indices": 8, "material": 0}]}],
        "accessors": acc, "bufferViews": views,
        "materials": [{"pbrMetallicRoughness": {"baseColorTexture": {"index": 0}, "metallicRoughnessTexture": {"index": 3}, "metallicFactor": 0.9, "roughnessFactor": 0.8},
                       "normalTexture": {"index": 1}, "emissiveTexture": {"index": 2}},
                      {"pbrMetallicRoughness": {"baseColorFactor": [0.9, 0.3, 0.2, 1.0], "metallicFactor": 0.2, "roughnessFactor": 0.4}}],
        "textures": [{"source": k} for k in range(4)], "images": images,
        "buffers": [{"byteLength": sum(len(c_) for c_ in chunks)}],
    }
    blob = b"".join(chunks)
    js = json.dumps(doc).encode(); js += b" " * (-len(js) % 4)
    body = struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(blob), 0x004E4942) + blob
    with open(path, "wb") as f:
        f.write(struct.pack("<III", 0x46546C67, 2, 12 + len(body)) + body)
    return {"triangles": (len(upper) + len(lower) + 6) // 3, "textures": [base, normal, emissive, mr]}


def blur_inputs():
    """Seeded inputs of the blur.comp vectors (shared by tests/golden/make_spirv_vectors.py and tests/test_oracle_spirv.py): byte planes
    at the shader's compiled-in 1700 x 900 x 3 lights, a motion plane with static, shifted, noisy and sky (0 -> decoded -1) bands,
    and the invocations that were executed."""
    W, H, L = 1700, 900, 3
    rng = np.random.default_rng(77)
    shadow = rng.choice(np.array([0, 255, 255, 255, 64, 191], np.uint8), size=(L, H, W))
    hist = rng.integers(0, 256, size=(L, H, W), dtype=np.uint8)
    motion = np.full((H, W, 2), 0.5, np.float16)
    ys, xs = np.mgrid[0:H, 0:W]
    band = (ys // 60) % 3
    motion[band == 1] = np.float16(0.5 + 3.0 / W)
    motion[band == 2, 0] = (0.5 + rng.uniform(-0.02, 0.02, size=int((band == 2).sum()))).astype(np.float16)
    motion[:40] = 0
    pix = [(0, 0), (W - 1, 0), (0, H - 1), (W - 1, H - 1), (1, 1), (W // 2, 0), (0, H // 2), (W - 1, H // 2), (W // 2, H - 1)]
    pix += [(int(x), int(y)) for x, y in zip(rng.integers(0, W, 260), rng.integers(0, H, 260))]
    return W, H, L, shadow, hist, motion, np.array(pix, np.int32)
