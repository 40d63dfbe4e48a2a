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



# ------------------------------------------------------------------------------------------------------- scene plumbing
def golden_scene(index=0, detail=None):
    import vulkanhybridraytracer_b200 as vk
    from oracle.oracle_py import Oracle
    from util import GOLDEN_DETAIL, GOLDEN_W, GOLDEN_H
    sc = vk.HostScene(index, detail=detail or GOLDEN_DETAIL)
    return vk, sc, Oracle(sc), GOLDEN_W, GOLDEN_H


def mat4_cols(m16):
    a = np.asarray(m16, dtype=F32).reshape(4, 4)          # column-major float[16]: row k of the reshape = column k
    return [a[k].copy() for k in range(4)]


def lights_block(sc, lights=None):
    return [[light_value(l["position"], l["color"], l["radius"]) for l in (sc.lights if lights is None else lights)]]


def materials_block(sc):
    return [[[vec(*m["diffuse"]), vec(*m["textures"]), vec(*m["shadingMetallicRoughness"])] for m in sc.materials]]


def gbuffer_images(g):
    """gbuffers[0..2] as bound by the shadow pass (position | material index, normal, motion), fp16 storage read as float."""
    h, w = g.position.shape[:2]
    motion = np.zeros((h, w, 4), F32); motion[..., :2] = g.motion.astype(F32); motion[..., 3] = 1
    return [Image2D(g.position.astype(F32)), Image2D(g.normal.astype(F32)), Image2D(motion)]


class ShadowPipeline:
    """traceRayEXT of a ray-gen shader that only asks "occluded?": the oracle answers the any-hit query on the ray the shader
    built; on a miss shadow.rmiss.spv main() runs on the caller's payload."""

    def __init__(self, oracle, T):
        self.o, self.T = oracle, T
        self.miss = Interpreter(module("shadow.rmiss"))
        self.miss_payload = self.miss.m.global_by_name("shadowed")
        self.rays = []

    def __call__(self, interp, flags, cull, sbt_offset, sbt_stride, miss_index, origin, tmin, direction, tmax, payload):
        assert flags == (0x01 | 0x04 | 0x08) and cull == 0xFF and miss_index in (0, 1)      # Opaque | TerminateOnFirstHit | SkipClosestHit
        ray = np.zeros(1, self.T.ray_dtype)
        ray["origin"] = origin; ray["dir"] = direction; ray["tmin"] = tmin; ray["tmax"] = tmax
        self.rays.append(ray[0].copy())
        hit = self.o.trace_rays(ray, any_hit=True)
        if hit["t"][0] < 0:
            assert not payload.path
            self.miss.cells[self.miss_payload] = payload.cell
            self.miss.run_entry()


# ------------------------------------------------------------------------------------------------------- a4: shadowRaygen.rgen
def gen_shadow():
    vk, sc, o, W, H = golden_scene()
    from vulkanhybridraytracer_b200 import types as T
    cam = sc.camera(W, H)
    m = module("shadowRaygen.rgen")
    out = {}
    for tag, frame, samples, step in (("f0", 0, 1, 1), ("f3s4", 3, 4, 3)):
        p = vk.make_params(W, H, frame=frame, shadow_samples=samples)
        g = o.gbuffer(cam, p)
        pipe = ShadowPipeline(o, T)
        shadow_imgs = [Image2D(np.zeros((H, W, 4), F32), unorm8=True) for _ in sc.lights]
        res = {"cam": [mat4_cols(cam.inv_view), mat4_cols(cam.inv_proj), F32(frame)], "Lights": lights_block(sc), "SampleBuffer": [I32(samples)],
               "MaterialBuffer": materials_block(sc), "gbuffers": gbuffer_images(g), "shadowImage": shadow_imgs,
               "LaunchSizeKHR": np.array([W, H, 1], U32)}
        it = Interpreter(m, res, trace_ray=pipe)
        launch = it.m.builtin_globals()["LaunchIdKHR"]
        pixels, bytes_, ray_rows = [], [], []
        for y in range(0, H, step):
            for x in range(0, W, step):
                it.cells[launch][0] = np.array([x, y, 0], U32)
                n0 = len(pipe.rays)
                it.run_entry()
                pixels.append((x, y))
                bytes_.append([int(np.rint(img.a[y, x, 0] * 255)) for img in shadow_imgs])
                rays = pipe.rays[n0:]
                if samples > 1 and rays:
                    # App. A #1 straight from the binary: every sample of every light of a pixel re-uses the same (r1, r2), so the
                    # `samples` rays of one light are one and the same ray
                    per_light = len(rays) // samples
                    assert per_light * samples == len(rays)
                    for k in range(per_light):
                        grp = rays[k * samples:(k + 1) * samples]
                        assert all(np.array_equal(grp[0]["dir"], r["dir"]) for r in grp)
                for r in rays:
                    ray_rows.append([x, y] + r["origin"].tolist() + r["dir"].tolist() + [float(r["tmax"])])
        out[f"{tag}_params"] = np.array([W, H, frame, samples], np.int32)
        out[f"{tag}_pixels"] = np.array(pixels, np.int32)
        out[f"{tag}_shadow"] = np.array(bytes_, np.uint8)
        out[f"{tag}_rays"] = np.array(ray_rows, np.float32)
        print(tag, "pixels", len(pixels), "rays", len(ray_rows), "interpreted instructions", it.steps)
    np.savez_compressed(os.path.join(OUT, "spirv_shadow.npz"), **out)
    o.close(); sc.close()


# ------------------------------------------------------------------------------------------------------- a6-a8: hybridRaygen.rgen + hybridHit.rchit + miss.rmiss
def unorm_image(a_u8, repeat=True):
    return Image2D(np.asarray(a_u8, np.uint8).astype(F32) / F32(255), repeat=repeat)


def hybrid_gbuffer_images(g):
    """gbuffers[0..5] of the hybrid descriptor set: position, normal, albedo, motion, material, emissive (renderer.cpp:2932-2958)."""
    h, w = g.position.shape[:2]
    motion = np.zeros((h, w, 4), F32); motion[..., :2] = g.motion.astype(F32); motion[..., 3] = 1
    return [Image2D(g.position.astype(F32)), Image2D(g.normal.astype(F32)), unorm_image(g.albedo), Image2D(motion), unorm_image(g.material), unorm_image(g.emissive)]


class HybridPipeline:
    """traceRayEXT of hybridRaygen.rgen: closest hit by the oracle on the ray the shader built, then hybridHit.rchit.spv or
    miss.rmiss.spv main() on the caller's payload; hybridHit's own shadow rays go through ShadowPipeline."""

    def __init__(self, sc, oracle, T, hit_module="hybridHit.rchit", env_name="environmentTexture", extra=None, lights=None):
        self.o, self.T, self.sc = oracle, T, sc
        self.shadow = ShadowPipeline(oracle, T)
        verts, inds = {}, {}
        for mid, (v, ix) in enumerate(sc.meshes):
            # rtVertexAttribute: normal.xyz1, color.xyz1, uv.xy11 (renderer.cpp:3027-3029)
            verts[mid] = [[[vec(n[0], n[1], n[2], 1), vec(c[0], c[1], c[2], 1), vec(u[0], u[1], 1, 1)] for n, c, u in zip(v["normal"], v["color"], v["uv"])]]
            inds[mid] = [[I32(int(k)) for k in ix]]
        inst = sc.instances
        env = [unorm_image(sc.textures[sc.sky_tex]), unorm_image(sc.textures[sc.irradiance_tex])]
        res = {"Lights": lights_block(sc, lights), "MaterialBuffer": materials_block(sc),
               "vertices": [verts[int(i["mesh_id"])] for i in inst], "indices": [inds[int(i["mesh_id"])] for i in inst],
               "textures": [unorm_image(t) for t in sc.textures], env_name: env,
               "sceneBuffer": [[vec(k, int(i["material_id"]), k, int(i["first_index"])) for k, i in enumerate(inst)]],
               "Matrices": [[mat4_cols(i["model"]) for i in inst]]}
        res.update(extra or {})
        self.hit = Interpreter(module(hit_module), res, trace_ray=self.shadow)
        self.miss = Interpreter(module("miss.rmiss"), {"skybox": env})
        self.hit_prd = self.hit.m.global_by_name("prd"); self.miss_prd = self.miss.m.global_by_name("prd")
        self.hb = self.hit.m.builtin_globals(); self.mb = self.miss.m.builtin_globals()
        self.hit_attr = self.hit.m.global_by_name("attribs")
        self.rays, self.hits = [], []

    def __call__(self, interp, flags, cull, sbt_offset, sbt_stride, miss_index, origin, tmin, direction, tmax, payload):
        assert flags == 0x01 and cull == 0xFF and miss_index == 0 and not payload.path
        ray = np.zeros(1, self.T.ray_dtype)
        ray["origin"] = origin; ray["dir"] = direction; ray["tmin"] = tmin; ray["tmax"] = tmax
        h = self.o.trace_rays(ray, any_hit=False)[0]
        self.rays.append(ray[0].copy()); self.hits.append((int(h["instance"]), int(h["primitive"]), float(h["t"])))
        if h["instance"] < 0:
            self.miss.cells[self.miss_prd] = payload.cell
            self.miss.cells[self.mb["WorldRayDirectionKHR"]][0] = np.asarray(direction, F32).copy()
            self.miss.run_entry()
            return
        it = self.hit
        it.cells[self.hit_prd] = payload.cell
        it.cells[self.hit_attr][0] = vec(h["u"], h["v"], 0)
        it.cells[self.hb["InstanceCustomIndexKHR"]][0] = I32(int(h["instance"]))
        it.cells[self.hb["PrimitiveId"]][0] = I32(int(h["primitive"]))
        it.cells[self.hb["WorldRayOriginKHR"]][0] = np.asarray(origin, F32).copy()
        it.cells[self.hb["WorldRayDirectionKHR"]][0] = np.asarray(direction, F32).copy()
        it.cells[self.hb["RayTmaxKHR"]][0] = F32(h["t"])          # gl_HitTEXT is RayTmaxKHR in a hit shader
        it.run_entry()


def gen_resolve():
    vk, sc, o, W, H = golden_scene()
    from vulkanhybridraytracer_b200 import types as T
    cam = sc.camera(W, H)
    out = {}
    frame = 0
    p = vk.make_params(W, H, frame=frame)
    fr = o.hybrid_frame(cam, p)
    g, den = fr["g"], fr["denoised"]
    pipe = HybridPipeline(sc, o, T)
    image = Image2D(np.zeros((H, W, 4), F32))          # the float value BEFORE the store into the B8G8R8A8 target (App. A #7)
    env = [unorm_image(sc.textures[sc.sky_tex]), unorm_image(sc.textures[sc.irradiance_tex])]
    shadow_imgs = []
    for l in range(len(sc.lights)):
        a = np.zeros((H, W, 4), F32); a[..., 0] = den[l].astype(F32) / F32(255); a[..., 3] = 1
        shadow_imgs.append(Image2D(a))
    res = {"cam": [mat4_cols(cam.inv_view), mat4_cols(cam.inv_proj), vec(frame, 0, 0, 0)], "gbuffers": hybrid_gbuffer_images(g), "Lights": lights_block(sc),
           "MaterialBuffer": materials_block(sc), "environmentTexture": env, "shadowImage": shadow_imgs, "image": image,
           "LaunchSizeKHR": np.array([W, H, 1], U32)}
    it = Interpreter(module("hybridRaygen.rgen"), res, trace_ray=pipe)
    launch = it.m.builtin_globals()["LaunchIdKHR"]
    pixels, colors, n_rays, chain = [], [], [], []
    for y in range(H):
        for x in range(W):
            it.cells[launch][0] = np.array([x, y, 0], U32)
            n0 = len(pipe.rays); s0 = len(pipe.shadow.rays)
            it.run_entry()
            pixels.append((x, y)); colors.append(image.a[y, x].copy())
            n_rays.append((len(pipe.rays) - n0, len(pipe.shadow.rays) - s0))
            for k in range(n0, len(pipe.rays)):
                chain.append([x, y, k - n0, pipe.hits[k][0], pipe.hits[k][1]])
    out["params"] = np.array([W, H, frame], np.int32)
    out["pixels"] = np.array(pixels, np.int32); out["final"] = np.array(colors, np.float32); out["ray_counts"] = np.array(n_rays, np.int32)
    out["chain_hits"] = np.array(chain, np.int32).reshape(-1, 5)
    np.savez_compressed(os.path.join(OUT, "spirv_resolve.npz"), **out)
    print("resolve: pixels", len(pixels), "reflection rays", len(pipe.rays), "secondary shadow rays", len(pipe.shadow.rays),
          "instructions", it.steps + pipe.hit.steps + pipe.miss.steps)
    o.close(); sc.close()


# ------------------------------------------------------------------------------------------------------- a12: deferred.frag
def io_by_location(m, storage):
    """{location: global id} of the Input (1) or Output (3) interface variables."""
    return {m.decor[g][30][0]: g for g, (t, st, _i) in m.globals.items() if st == storage and 30 in m.decor.get(g, {})}


def gen_deferred():
    vk, sc, o, W, H = golden_scene()
    cam = sc.camera(W, H)
    g = o.gbuffer(cam, vk.make_params(W, H))
    imgs = hybrid_gbuffer_images(g)
    m = module("deferred.frag")
    out = {"params": np.array([W, H], np.int32)}
    for target in range(7):
        res = {"positionTexture": imgs[0], "normalTexture": imgs[1], "albedoTexture": imgs[2], "motionTexture": imgs[3], "materialTexture": imgs[4],
               "emissiveTexture": imgs[5], "environmentTexture": unorm_image(sc.textures[sc.irradiance_tex]), "LightBuffer": lights_block(sc),
               "debugInfo": [I32(target)]}
        it = Interpreter(m, res)
        ins, outs = io_by_location(m, 1), io_by_location(m, 3)
        step = 1 if target == 0 else 5
        pixels, colors = [], []
        for y in range(0, H, step):
            for x in range(0, W, step):
                it.cells[ins[0]][0] = vec((x + 0.5) / W, (y + 0.5) / H)            # quad.vert: the full-screen triangle's uv at the pixel centre
                it.cells[ins[1]][0] = vec(*cam.position[:3])
                it.run_entry()
                pixels.append((x, y)); colors.append(it.cells[outs[0]][0].copy())
        out[f"t{target}_pixels"] = np.array(pixels, np.int32); out[f"t{target}_color"] = np.array(colors, np.float32)
        print("deferred target", target, "pixels", len(pixels))
    np.savez_compressed(os.path.join(OUT, "spirv_deferred.npz"), **out)
    o.close(); sc.close()


# ------------------------------------------------------------------------------------------------------- a10: postFragment.frag
def gen_post():
    m = module("postFragment.frag")
    levels = np.arange(256, dtype=np.uint8)
    tex = np.zeros((1, 256, 4), np.uint8)
    tex[0, :, 0] = levels; tex[0, :, 1] = levels[::-1]; tex[0, :, 2] = (levels.astype(int) * 7 + 3) % 256; tex[0, :, 3] = 255
    it = Interpreter(m, {"finalTexture": unorm_image(tex)})
    ins, outs = io_by_location(m, 1), io_by_location(m, 3)
    res = []
    for k in range(256):
        it.cells[ins[0]][0] = vec((k + 0.5) / 256, 0.5)
        it.run_entry()
        res.append(it.cells[outs[0]][0].copy())
    np.savez_compressed(os.path.join(OUT, "spirv_post.npz"), texels=tex, frag_color=np.array(res, np.float32))
    print("post: 256 texels")


# ------------------------------------------------------------------------------------------------------- a5: blur.comp
def gen_blur():
    """blur.comp.spv at its compiled-in size (WIDTH 1700, HEIGHT 900, NLIGHTS 3) on seeded byte planes.  imageLoad / imageStore work
    on R8_UNORM images (renderer.cpp:1538) declared rgba8 in the shader: reads give (r, 0, 0, 1).  Stores go to a second buffer so
    that every invocation sees last frame's history, as on a GPU where the neighbours have not run yet (the oracle copies it too).
    The motion texture is sampled exactly on texel corners (uv = x / W): which texel that returns is implementation-defined, the
    stub snaps to texel x like the oracle's integer fetch."""
    from util import blur_inputs
    W, H, L, shadow, hist, motion, pix = blur_inputs()
    pix = [(int(x), int(y)) for x, y in pix]
    m = module("blur.comp")
    out = {"pixels": np.array(pix, np.int32)}
    for frame in (0, 5):
        def rgba(plane):
            a = np.zeros((H, W, 4), F32); a[..., 0] = plane.astype(F32) / F32(255); a[..., 3] = 1
            return a
        inputs = [Image2D(rgba(shadow[l])) for l in range(L)]
        outs_buf = [rgba(hist[l]) for l in range(L)]
        outputs = [Image2D(rgba(hist[l]), unorm8=True, write_to=outs_buf[l]) for l in range(L)]
        mo = np.zeros((H, W, 4), F32); mo[..., :2] = motion.astype(F32); mo[..., 3] = 1
        res = {"inputImage": inputs, "outputImage": outputs, "FrameCount": [I32(frame)], "motionTexture": Image2D(mo, repeat=False, corner_eps=1e-3)}
        it = Interpreter(m, res)
        gid = it.m.builtin_globals()["GlobalInvocationId"]
        vals = []
        for (x, y) in pix:
            it.cells[gid][0] = np.array([x, y, 0], U32)
            it.run_entry()
            vals.append([int(np.rint(outs_buf[l][y, x, 0] * 255)) for l in range(L)])
        out[f"f{frame}_denoised"] = np.array(vals, np.uint8)
        print("blur frame", frame, "pixels", len(pix), "instructions", it.steps)
    np.savez_compressed(os.path.join(OUT, "spirv_blur.npz"), **out)


# ------------------------------------------------------------------------------------------------------- a3: basic.vert + geometry_shader.frag
def camera_ray(cam, x, y, W, H):
    """raygen.rgen:32-40 in float32 (the oracle's primary ray: DESIGN.md section 2, "G-buffer by ray casting")."""
    u, v = F32(F32(x) + F32(0.5)) / F32(W), F32(F32(y) + F32(0.5)) / F32(H)
    d = vec(u * F32(2) - F32(1), v * F32(2) - F32(1), 1, 1)
    ip, iv = mat4_cols(cam.inv_proj), mat4_cols(cam.inv_view)
    t = ip[0] * d[0] + ip[1] * d[1] + ip[2] * d[2] + ip[3] * d[3]
    tn = t[:3] / np.sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2])
    dirv = (iv[0] * tn[0] + iv[1] * tn[1] + iv[2] * tn[2])[:3]
    return iv[3][:3].copy(), dirv.astype(F32)


def gen_gbuffer():
    """The raster G-buffer pass, per covered pixel: basic.vert.spv main() on the three vertices of the triangle the oracle's primary
    ray hits, varyings interpolated with the hit's barycentrics (what the rasteriser's perspective-correct interpolation gives at
    that point) and handed over BY LOCATION to geometry_shader.frag.spv main().  dFdx / dFdy of the world position -- only used by
    the normal-mapped material -- are supplied as the oracle defines them (neighbour camera rays on the triangle's plane)."""
    vk, sc, o, W, H = golden_scene()
    from vulkanhybridraytracer_b200 import types as T
    cam0 = sc.camera(W, H)
    sc.move_camera(3, 6.0); sc.rotate_camera(12.0, 4.0)
    cam = sc.camera(W, H, prev=cam0)                      # a moving camera: motion vectors are not zero
    p = vk.make_params(W, H)
    g = o.gbuffer(cam, p)
    vm, fm = module("basic.vert"), module("geometry_shader.frag")
    cam_block = [mat4_cols(cam.view), mat4_cols(cam.proj), mat4_cols(cam.prev_view), mat4_cols(cam.prev_proj)]
    vs = Interpreter(vm, {"CameraBuffer": cam_block})
    v_in, v_out = io_by_location(vm, 1), io_by_location(vm, 3)
    v_push = vm.global_by_name("pushC")
    deriv = {}
    fs = Interpreter(fm, {"textures": [unorm_image(t) for t in sc.textures]}, derivative=lambda rid, axis, value: deriv["d"][axis])
    f_in, f_out = io_by_location(fm, 1), io_by_location(fm, 3)
    f_push = fm.global_by_name("pushC")
    rows = {k: [] for k in ("pixels", "position", "normal", "albedo", "motion", "material", "emissive", "normal_mapped")}
    for y in range(H):
        for x in range(W):
            inst, prim = [int(t) for t in g.hit_ids[y, x]]
            if inst < 0:
                continue
            org, dirv = camera_ray(cam, x, y, W, H)
            ray = np.zeros(1, T.ray_dtype); ray["origin"] = org + dirv * F32(1e-2); ray["dir"] = dirv; ray["tmin"] = 0.001; ray["tmax"] = 10000.0
            h = o.trace_rays(ray)[0]
            if int(h["instance"]) != inst or int(h["primitive"]) != prim:
                continue                                  # the float32 re-derivation of the ray picked a neighbour at an edge: skip
            I = sc.instances[inst]
            verts, idx = sc.meshes[int(I["mesh_id"])]
            tri = idx[int(I["first_index"]) + 3 * prim: int(I["first_index"]) + 3 * prim + 3]
            vs.cells[v_push][0] = [mat4_cols(I["model"]), mat4_cols(I["inv_model"])]
            outs = []
            for k in tri:
                vtx = verts[int(k)]
                vs.cells[v_in[0]][0] = vec(*vtx["position"]); vs.cells[v_in[1]][0] = vec(*vtx["normal"])
                vs.cells[v_in[2]][0] = vec(*vtx["color"]); vs.cells[v_in[3]][0] = vec(*vtx["uv"])
                vs.run_entry()
                outs.append({loc: vs.cells[gid][0].copy() for loc, gid in v_out.items()})
            b = [F32(1) - F32(h["u"]) - F32(h["v"]), F32(h["u"]), F32(h["v"])]
            for loc, gid in f_in.items():
                fs.cells[gid][0] = outs[0][loc] * b[0] + outs[1][loc] * b[1] + outs[2][loc] * b[2]
            mat = sc.materials[int(I["material_id"])]
            fs.cells[f_push][0] = [vec(*mat["diffuse"]), vec(*mat["textures"]), vec(*mat["shadingMetallicRoughness"])]
            mapped = mat["textures"][1] > -1
            if mapped:
                P = [outs[k][0] for k in range(3)]
                ng = np.cross(P[1] - P[0], P[2] - P[0]).astype(F32)
                wp = fs.cells[f_in[0]][0]
                d = []
                for (nx, ny) in ((x + 1, y), (x, y + 1)):
                    o2, d2 = camera_ray(cam, nx, ny, W, H)
                    q = o2 + d2 * (np.dot(P[0] - o2, ng) / np.dot(d2, ng))
                    d.append((q - wp).astype(F32))
                deriv["d"] = d
            fs.run_entry()
            res = {loc: fs.cells[gid][0].copy() for loc, gid in f_out.items()}
            rows["pixels"].append((x, y)); rows["position"].append(res[0]); rows["normal"].append(res[1]); rows["albedo"].append(res[2])
            rows["motion"].append(res[3]); rows["material"].append(res[4]); rows["emissive"].append(res[5]); rows["normal_mapped"].append(bool(mapped))
    out = {k: np.array(v, np.int32 if k == "pixels" else (np.bool_ if k == "normal_mapped" else np.float32)) for k, v in rows.items()}
    out["params"] = np.array([W, H], np.int32)
    np.savez_compressed(os.path.join(OUT, "spirv_gbuffer.npz"), **out)
    print("gbuffer: pixels", len(rows["pixels"]), "normal-mapped", int(np.sum(out["normal_mapped"])))
    o.close(); sc.close()


# ------------------------------------------------------------------------------------------------------- a11: raygen.rgen + closesthit.rchit + miss.rmiss
def gen_raytrace():
    vk, sc, o, W, H = golden_scene()
    from vulkanhybridraytracer_b200 import types as T
    cam = sc.camera(W, H)
    out = {}
    for tag, frame, samples, step in (("f0s1", 0, 1, 1), ("f2s3", 2, 3, 3)):
        pipe = HybridPipeline(sc, o, T, hit_module="closesthit.rchit", env_name="environment_texture",
                              extra={"SampleBuffer": [I32(samples)], "shadowImage": [Image2D(np.zeros((1, 1, 4), F32)) for _ in sc.lights]})
        rng = np.random.default_rng(5)
        prev = rng.uniform(0, 1, size=(H, W, 4)).astype(F32) if frame else np.zeros((H, W, 4), F32)      # the accumulation image
        image = Image2D(prev)
        res = {"cam": [mat4_cols(cam.inv_view), mat4_cols(cam.inv_proj), vec(frame, 0, 0, 0)], "image": image, "LaunchSizeKHR": np.array([W, H, 1], U32)}
        it = Interpreter(module("raygen.rgen"), res, trace_ray=pipe)
        launch = it.m.builtin_globals()["LaunchIdKHR"]
        pixels, colors, counts = [], [], []
        for y in range(0, H, step):
            for x in range(0, W, step):
                it.cells[launch][0] = np.array([x, y, 0], U32)
                n0, s0 = len(pipe.rays), len(pipe.shadow.rays)
                it.run_entry()
                pixels.append((x, y)); colors.append(image.a[y, x].copy()); counts.append((len(pipe.rays) - n0, len(pipe.shadow.rays) - s0))
        out[f"{tag}_params"] = np.array([W, H, frame, samples], np.int32); out[f"{tag}_pixels"] = np.array(pixels, np.int32)
        out[f"{tag}_color"] = np.array(colors, np.float32); out[f"{tag}_ray_counts"] = np.array(counts, np.int32)
        if frame:
            out[f"{tag}_previous"] = prev
        print("raytrace", tag, "pixels", len(pixels), "rays", len(pipe.rays), "shadow rays", len(pipe.shadow.rays))
    np.savez_compressed(os.path.join(OUT, "spirv_raytrace.npz"), **out)
    o.close(); sc.close()


# ------------------------------------------------------------------------------------------------------- light edge cases through a4 + a6-a8
def variant_lights(sc):
    """The branches the default scene never takes: a directional light (pos.w < 0: L = pos.xyz, attenuation 0.3 in hybridHit,
    intensity 1), a point light whose maxDistance most of the scene lies beyond (unshadowed in shadowRaygen.rgen:81-83 but
    shadowed in hybridHit.rchit:115-125, attenuation clamps to 0), and a light of zero intensity (attenuation = 0, the
    irradiance branch of the mirror / glass terms)."""
    lights = sc.lights.copy()
    lights["position"][0] = [0.3, 0.8, 0.5, -1.0]
    lights["position"][1][3] = 14.0
    lights["color"][2][3] = 0.0
    return lights


def gen_variants():
    vk, sc, o, W, H = golden_scene()
    from vulkanhybridraytracer_b200 import types as T
    lights = variant_lights(sc)
    o.set_lights(lights)
    cam = sc.camera(W, H)
    p = vk.make_params(W, H, frame=1, shadow_samples=2)
    g = o.gbuffer(cam, p)
    # a4
    pipe = ShadowPipeline(o, T)
    shadow_imgs = [Image2D(np.zeros((H, W, 4), F32), unorm8=True) for _ in lights]
    res = {"cam": [mat4_cols(cam.inv_view), mat4_cols(cam.inv_proj), F32(1)], "Lights": lights_block(sc, lights), "SampleBuffer": [I32(2)],
           "MaterialBuffer": materials_block(sc), "gbuffers": gbuffer_images(g), "shadowImage": shadow_imgs, "LaunchSizeKHR": np.array([W, H, 1], U32)}
    it = Interpreter(module("shadowRaygen.rgen"), res, trace_ray=pipe)
    launch = it.m.builtin_globals()["LaunchIdKHR"]
    pixels, sh = [], []
    for y in range(0, H, 2):
        for x in range(0, W, 2):
            it.cells[launch][0] = np.array([x, y, 0], U32)
            it.run_entry()
            pixels.append((x, y)); sh.append([int(np.rint(img.a[y, x, 0] * 255)) for img in shadow_imgs])
    out = {"params": np.array([W, H, 1, 2], np.int32), "lights": lights, "pixels": np.array(pixels, np.int32), "shadow": np.array(sh, np.uint8)}
    # a6-a8 on the oracle's filtered planes of that frame
    fr = o.hybrid_frame(cam, p)
    den = fr["denoised"]
    hp = HybridPipeline(sc, o, T, lights=lights)
    image = Image2D(np.zeros((H, W, 4), F32))
    env = [unorm_image(sc.textures[sc.sky_tex]), unorm_image(sc.textures[sc.irradiance_tex])]
    simgs = []
    for l in range(len(lights)):
        a = np.zeros((H, W, 4), F32); a[..., 0] = den[l].astype(F32) / F32(255); a[..., 3] = 1
        simgs.append(Image2D(a))
    res = {"cam": [mat4_cols(cam.inv_view), mat4_cols(cam.inv_proj), vec(1, 0, 0, 0)], "gbuffers": hybrid_gbuffer_images(g), "Lights": lights_block(sc, lights),
           "MaterialBuffer": materials_block(sc), "environmentTexture": env, "shadowImage": simgs, "image": image, "LaunchSizeKHR": np.array([W, H, 1], U32)}
    it = Interpreter(module("hybridRaygen.rgen"), res, trace_ray=hp)
    launch = it.m.builtin_globals()["LaunchIdKHR"]
    colors = []
    for (x, y) in pixels:
        it.cells[launch][0] = np.array([x, y, 0], U32)
        it.run_entry()
        colors.append(image.a[y, x].copy())
    out["final"] = np.array(colors, np.float32)
    np.savez_compressed(os.path.join(OUT, "spirv_variants.npz"), **out)
    print("variants: pixels", len(pixels), "shadow rays", len(pipe.rays), "chain rays", len(hp.rays), "secondary shadow rays", len(hp.shadow.rays))
    o.close(); sc.close()


# ------------------------------------------------------------------------------------------------------- a6-a8 on the other scenes
SCENES_DETAIL = (3, 2, 3, 14)      # 14 instances in the stress scene


def gen_scenes():
    """The resolve + hit + miss binaries on the Cornell scene (scene.cpp:162-272) and on the instanced stress scene (rotated and
    non-uniformly scaled instances: hybridHit.rchit:50 takes the normal matrix from inverse(model) in the shader, the oracle from
    the host-supplied inverse; materials cycle through the three shading models)."""
    from vulkanhybridraytracer_b200 import types as T
    out = {}
    for index, step in ((1, 1), (2, 1)):
        vk, sc, o, W, H = golden_scene(index, SCENES_DETAIL)
        cam = sc.camera(W, H)
        p = vk.make_params(W, H, frame=2)
        fr = o.hybrid_frame(cam, p)
        g, den = fr["g"], fr["denoised"]
        pipe = HybridPipeline(sc, o, T)
        image = Image2D(np.zeros((H, W, 4), F32))
        env = [unorm_image(sc.textures[sc.sky_tex]), unorm_image(sc.textures[sc.irradiance_tex])]
        simgs = []
        for l in range(len(sc.lights)):
            a = np.zeros((H, W, 4), F32); a[..., 0] = den[l].astype(F32) / F32(255); a[..., 3] = 1
            simgs.append(Image2D(a))
        res = {"cam": [mat4_cols(cam.inv_view), mat4_cols(cam.inv_proj), vec(2, 0, 0, 0)], "gbuffers": hybrid_gbuffer_images(g), "Lights": lights_block(sc),
               "MaterialBuffer": materials_block(sc), "environmentTexture": env, "shadowImage": simgs, "image": image, "LaunchSizeKHR": np.array([W, H, 1], U32)}
        it = Interpreter(module("hybridRaygen.rgen"), res, trace_ray=pipe)
        launch = it.m.builtin_globals()["LaunchIdKHR"]
        pixels, colors = [], []
        for y in range(0, H, step):
            for x in range(0, W, step):
                it.cells[launch][0] = np.array([x, y, 0], U32)
                it.run_entry()
                pixels.append((x, y)); colors.append(image.a[y, x].copy())
        out[f"s{index}_params"] = np.array([W, H, 2], np.int32); out[f"s{index}_pixels"] = np.array(pixels, np.int32); out[f"s{index}_final"] = np.array(colors, np.float32)
        print("scene", index, "pixels", len(pixels), "chain rays", len(pipe.rays), "secondary shadow rays", len(pipe.shadow.rays), "lights", len(sc.lights), "instances", len(sc.instances))
        o.close(); sc.close()
    np.savez_compressed(os.path.join(OUT, "spirv_scenes.npz"), **out)

GENERATORS = {"scenes": gen_scenes, "variants": gen_variants, "raytrace": gen_raytrace, "gbuffer": gen_gbuffer, "functions": gen_functions, "shadow": gen_shadow, "resolve": gen_resolve, "deferred": gen_deferred, "post": gen_post, "blur": gen_blur}

if __name__ == "__main__":
    which = sys.argv[1:] or list(GENERATORS)
    for w in which:
        GENERATORS[w]()
