"""Which shadow rays make thousands of node fetches?  Rebuilds the shadow ray of the (frame, pixel, light) triples that
e2e_probe5.py reported as kernel tails (shadowRaygen.rgen:31,70-80 + helpers.glsl:82-93, faithful seed-by-value RNG)."""
import numpy as np
import vulkanhybridraytracer_b200 as vk
from vulkanhybridraytracer_b200 import types as T, PLANE, STAGE

W, H = 1920, 1080
f32 = np.float32
def tea(v0, v1):
    s0 = 0
    for _ in range(16):
        s0 = (s0 + 0x9E3779B9) & 0xFFFFFFFF
        v0 = (v0 + ((((v1 << 4) & 0xFFFFFFFF) + 0xA341316C) ^ (v1 + s0) ^ ((v1 >> 5) + 0xC8013EA4))) & 0xFFFFFFFF
        v1 = (v1 + ((((v0 << 4) & 0xFFFFFFFF) + 0xAD90777D) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7E95761E))) & 0xFFFFFFFF
    return v0
def rnd(seed):
    seed = (1664525 * seed + 1013904223) & 0xFFFFFFFF
    return seed, f32(seed & 0xFFFFFF) / f32(16777216.0)
def normalize(v): return (v / np.linalg.norm(v)).astype(f32)

sc = vk.HostScene(0)
ctx = vk.Context(0); ctx.upload_scene(sc)
ctx.set_camera(sc.camera(W, H)); ctx.set_params(vk.make_params(W, H)); ctx.run_stage(STAGE.GBUFFER)
pos = ctx.readback(PLANE.POSITION).astype(f32)
cases = [(19, 894, 877, 0), (48, 158, 1022, 1), (52, 1888, 1064, 0), (55, 90, 1007, 1), (94, 133, 1031, 1), (6, 399, 850, 1), (1, 1145, 1063, 1)]
for frame, x, y, l in cases:
    P = pos[y, x, :3]
    lp = np.array(sc.lights["position"][l][:3], f32); rad = f32(sc.lights["radius"][l])
    Lv = lp - P; dist = f32(np.linalg.norm(Lv)); Ld = normalize(Lv)
    seed = tea(y * W + x, frame)
    s, u1 = rnd(seed); s, u2 = rnd(s)
    radius = rad * np.sqrt(u1); angle = u2 * f32(2.0) * f32(3.14159265359)
    tang = normalize(np.cross(Ld, np.array([0, 1, 0], f32))); bit = normalize(np.cross(tang, Ld))
    target = P + Ld + radius * np.cos(angle) * tang + radius * np.sin(angle) * bit
    D = normalize(target - P)
    rays = np.zeros(1, T.ray_dtype)
    rays["origin"] = P + D * f32(0.1); rays["dir"] = D; rays["tmin"] = 0.001; rays["tmax"] = dist + 1
    st = ctx.trace_stats(rays, any_hit=True); hit = ctx.trace_rays(rays, any_hit=True)
    print(f"frame {frame} pixel ({x},{y}) light {l}: u=({u1:.6f},{u2:.6f}) dir {D} origin {rays['origin'][0]} -> node fetches {st[0,0]} tris {st[0,1]} hit {hit[0]}")
