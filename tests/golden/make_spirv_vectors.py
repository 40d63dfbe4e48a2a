"""Generates tests/golden/spirv_*.npz by EXECUTING THE REFERENCE'S OWN COMPILED SHADERS
(/root/reference/data/shaders/output/*.spv) in oracle/spirv_interp.py.  Run in the build container, where /root/reference
exists; the vectors are committed and tests/test_oracle_spirv.py holds the oracle to them without the reference present.

    python tests/golden/make_spirv_vectors.py [functions] [shadow] [resolve] [blur] [post] [deferred] [gbuffer]     (default: all)

What pins what (SURVEY 8a row -> reference binary):
  functions  random.glsl / helpers.glsl: tea, lcg, rnd, sampleDisk, DistributionGGX, GeometrySchlickGGX, GeometrySmith,
             FresnelSchlick as compiled into shadowRaygen.rgen.spv / hybridRaygen.rgen.spv / hybridHit.rchit.spv / deferred.frag.spv;
             cotangent_frame / perturbNormal of geometry_shader.frag.spv
  shadow     a4: shadowRaygen.rgen.spv main() per pixel; traceRayEXT is answered by the oracle's any-hit query on the ray the SHADER
             built (origin, tmin, direction, tmax are recorded and compared too)
  blur       a5: blur.comp.spv main() at its compiled-in 1700 x 900 x 3 lights, reset and accumulation frames
  resolve    a6-a8: hybridRaygen.rgen.spv main(); its traceRayEXT runs hybridHit.rchit.spv / miss.rmiss.spv main() in the same
             interpreter on the oracle's closest hit, their shadow rays again through the oracle's any-hit query
  post       a10: postFragment.frag.spv main()
  deferred   a12: deferred.frag.spv main(), every debug target
  gbuffer    a3: geometry_shader.frag.spv main() on the oracle's interpolated varyings (normal mapping, material packing)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.spirv_interp import F32, I32, U32, Image2D, Interpreter, Module, Ptr  # noqa: E402

SPV = "/root/reference/data/shaders/output"
OUT = os.path.join(ROOT, "tests", "golden")


def module(name):
    return Module(os.path.join(SPV, name + ".spv"))


def vec(*xs):
    return np.array(xs, dtype=F32)


def light_value(pos4, color4, radius):
    return [vec(*pos4), vec(*color4), F32(radius)]


# ------------------------------------------------------------------------------------------------------- functions
def gen_functions():
    rng = np.random.default_rng(2024)
    out = {}
    it = Interpreter(module("shadowRaygen.rgen"))
    # tea / lcg / rnd
    a = rng.integers(0, 1 << 32, size=64, dtype=np.uint64).astype(np.uint32)
    b = rng.integers(0, 1 << 32, size=64, dtype=np.uint64).astype(np.uint32)
    a[:6] = [0, 1, 0, 1919, 2073599, 1037760]; b[:6] = [0, 0, 1, 0, 0, 7]
    out["tea_in"] = np.stack([a, b], 1)
    out["tea_out"] = np.array([int(it.call("tea", U32(x), U32(y))[0]) for x, y in zip(a, b)], dtype=np.uint32)
    seeds = rng.integers(0, 1 << 32, size=64, dtype=np.uint64).astype(np.uint32)
    seeds[:3] = [0, 0xffffffff, 0x741c187d]
    lcg_val, lcg_state, rnd_val = [], [], []
    for s in seeds:
        r, st = it.call("lcg", U32(s)); lcg_val.append(int(r)); lcg_state.append(int(st[0]))
        r, st = it.call("rnd", U32(s)); rnd_val.append(np.float32(r))
    out["seed_in"] = seeds
    out["lcg_out"] = np.array(lcg_val, np.uint32); out["lcg_state"] = np.array(lcg_state, np.uint32); out["rnd_out"] = np.array(rnd_val, np.float32)
    # sampleDisk (seed by value: the caller's seed must come back unchanged)
    n = 96
    lights = np.zeros((n, 9), np.float32)
    pos = rng.uniform(-10, 10, size=(n, 3)).astype(np.float32)
    L = rng.normal(size=(n, 3)); L /= np.linalg.norm(L, axis=1, keepdims=True); L = L.astype(np.float32)
    L[0] = [0, 1, 0]; L[1] = [1, 0, 0]; L[2] = [0, 0, -1]                            # the special case and two axis directions
    lights[:, 8] = rng.uniform(0.01, 1.0, size=n); lights[:, :3] = rng.uniform(-15, 15, size=(n, 3)); lights[:, 3] = 500
    sd_seed = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    dirs, seeds_after = [], []
    for k in range(n):
        r, st = it.call("sampleDisk", light_value(lights[k, :4], lights[k, 4:8], lights[k, 8]), vec(*pos[k]), vec(*L[k]), U32(sd_seed[k]))
        dirs.append(r); seeds_after.append(int(st[3]))
    # inside the function the parameter copy advances by two LCG steps; the CALLER's seed does not (no copy-back after
    # OpFunctionCall sampleDisk -- SURVEY App. A #1): that shows in the whole-shader vectors below ("shadow")
    lcg2 = [(1664525 * ((1664525 * int(s) + 1013904223) & 0xffffffff) + 1013904223) & 0xffffffff for s in sd_seed]
    assert seeds_after == lcg2
    out["disk_light"] = lights; out["disk_pos"] = pos; out["disk_L"] = L; out["disk_seed"] = sd_seed; out["disk_dir"] = np.array(dirs, np.float32)
    # BRDF terms, one set per module that carries them (the source is shared: they must agree)
    n = 128
    N = rng.normal(size=(n, 3)); N /= np.linalg.norm(N, axis=1, keepdims=True)
    V = rng.normal(size=(n, 3)); V /= np.linalg.norm(V, axis=1, keepdims=True)
    Lb = rng.normal(size=(n, 3)); Lb /= np.linalg.norm(Lb, axis=1, keepdims=True)
    H = V + Lb; H /= np.linalg.norm(H, axis=1, keepdims=True)
    N, V, Lb, H = [x.astype(np.float32) for x in (N, V, Lb, H)]
    rough = rng.uniform(0.0, 1.0, size=n).astype(np.float32); rough[:2] = [0.0, 1.0]
    cos_t = rng.uniform(0.0, 1.0, size=n).astype(np.float32); cos_t[:2] = [0.0, 1.0]
    F0 = rng.uniform(0.0, 1.0, size=(n, 3)).astype(np.float32)
    out["brdf_N"], out["brdf_V"], out["brdf_L"], out["brdf_H"], out["brdf_rough"], out["brdf_cos"], out["brdf_F0"] = N, V, Lb, H, rough, cos_t, F0
    per_module = {}
    for name in ("hybridRaygen.rgen", "hybridHit.rchit", "closesthit.rchit", "deferred.frag"):
        im = Interpreter(module(name))
        D = [np.float32(im.call("DistributionGGX", vec(*N[k]), vec(*H[k]), F32(rough[k]))[0]) for k in range(n)]
        G1 = [np.float32(im.call("GeometrySchlickGGX", F32(cos_t[k]), F32(rough[k]))[0]) for k in range(n)]
        G = [np.float32(im.call("GeometrySmith", vec(*N[k]), vec(*V[k]), vec(*Lb[k]), F32(rough[k]))[0]) for k in range(n)]
        Fr = [im.call("FresnelSchlick", F32(cos_t[k]), vec(*F0[k]))[0] for k in range(n)]
        per_module[name] = (np.array(D, np.float32), np.array(G1, np.float32), np.array(G, np.float32), np.array(Fr, np.float32))
    ref = per_module["hybridRaygen.rgen"]
    for name, vals in per_module.items():
        for x, y in zip(ref, vals):
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), name       # same source, same binary code
    out["brdf_D"], out["brdf_G1"], out["brdf_G"], out["brdf_F"] = ref
    np.savez_compressed(os.path.join(OUT, "spirv_functions.npz"), **out)
    print("spirv_functions.npz:", {k: v.shape for k, v in out.items()})


GENERATORS = {"functions": gen_functions}

if __name__ == "__main__":
    which = sys.argv[1:] or list(GENERATORS)
    for w in which:
        GENERATORS[w]()
