"""The oracle against vectors produced by executing the REFERENCE'S OWN COMPILED SHADERS (CPU only).

`tests/golden/spirv_*.npz` come from `tests/golden/make_spirv_vectors.py`, which runs
/root/reference/data/shaders/output/*.spv in oracle/spirv_interp.py (the reference cannot be built or run, but it ships
its SPIR-V).  Integer results must be identical; float results are compared with the tolerance stated per test: the
interpreter evaluates sin / cos / pow / sqrt in float64 and rounds once, the oracle uses the numerics contract of
include/vkp_detmath.h, and Vulkan itself leaves those functions a few ulp of freedom.
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle_py
from vulkanhybridraytracer_b200 import types as T

from util import GOLDEN_DIR


def vectors(name):
    path = os.path.join(GOLDEN_DIR, f"spirv_{name}.npz")
    assert os.path.exists(path), f"{path} is missing: run tests/golden/make_spirv_vectors.py in the build container"
    return np.load(path)


@pytest.fixture(scope="module")
def L():
    return oracle_py.load()


def f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def test_tea_lcg_rnd_are_the_shader_binaries(L):
    """random.glsl:3-33 as compiled into shadowRaygen.rgen.spv: bit-exact, including the App. D rows."""
    z = vectors("functions")
    got = np.array([L.orc_tea(int(a), int(b)) for a, b in z["tea_in"]], dtype=np.uint32)
    assert np.array_equal(got, z["tea_out"])
    assert [hex(int(x)) for x in z["tea_out"][:6]] == ["0x741c187d", "0x8da6b311", "0x70d3aef1", "0x97cfe7dd", "0xcda8c132", "0xc908dec6"]
    for s, want_v, want_s, want_r in zip(z["seed_in"], z["lcg_out"], z["lcg_state"], z["rnd_out"]):
        st = C.c_uint32(int(s))
        assert L.orc_lcg(C.byref(st)) == int(want_v) and st.value == int(want_s)
        st = C.c_uint32(int(s))
        r = np.float32(L.orc_rnd(C.byref(st)))
        assert r.view(np.uint32) == np.float32(want_r).view(np.uint32) and st.value == int(want_s)


def test_sample_disk_is_the_shader_binary(L):
    """helpers.glsl:82-93 (sampleDisk of shadowRaygen.rgen.spv), including L == (0,1,0).  cos / sin / sqrt / normalize: 2e-6 absolute
    on a unit vector (a few ulp of the contract's sin / cos at angles up to 2 pi)."""
    z = vectors("functions")
    light = T.Light()
    worst = 0.0
    for k in range(len(z["disk_seed"])):
        row = z["disk_light"][k]
        for c in range(4):
            light.position[c] = float(row[c]); light.color[c] = float(row[4 + c])
        light.radius = float(row[8])
        out = (C.c_float * 3)()
        L.orc_sample_disk(C.byref(light), C.byref(f3(z["disk_pos"][k])), C.byref(f3(z["disk_L"][k])), int(z["disk_seed"][k]), C.byref(out))
        worst = max(worst, float(np.abs(np.array(out[:], np.float32) - z["disk_dir"][k]).max()))
    assert worst <= 2e-6, worst


def test_brdf_terms_are_the_shader_binaries(L):
    """DistributionGGX / GeometrySmith / FresnelSchlick (helpers.glsl:33-72) as compiled into hybridRaygen.rgen.spv (identical
    bits in hybridHit.rchit.spv, closesthit.rchit.spv and deferred.frag.spv: asserted when the vectors are made).  Only + - * /
    max and pow(x, 5): relative 1e-6 (times the conditioning of D with respect to dot(N, H), see below)."""
    z = vectors("functions")
    n = len(z["brdf_rough"])
    L.orc_distribution_ggx.restype = C.c_float
    L.orc_geometry_smith.restype = C.c_float
    for k in range(n):
        d = L.orc_distribution_ggx(C.byref(f3(z["brdf_N"][k])), C.byref(f3(z["brdf_H"][k])), C.c_float(float(z["brdf_rough"][k])))
        g = L.orc_geometry_smith(C.byref(f3(z["brdf_N"][k])), C.byref(f3(z["brdf_V"][k])), C.byref(f3(z["brdf_L"][k])), C.c_float(float(z["brdf_rough"][k])))
        fr = (C.c_float * 3)()
        L.orc_fresnel_schlick(C.c_float(float(z["brdf_cos"][k])), C.byref(f3(z["brdf_F0"][k])), C.byref(fr))
        wd, wg, wf = float(z["brdf_D"][k]), float(z["brdf_G"][k]), z["brdf_F"][k]
        if np.isfinite(wd):
            # D = a2 / (pi * (NdotH^2 * (a2 - 1) + 1)^2) amplifies the last bit of dot(N, H) by cond = 4 * NdotH^2 * (1 - a2) / (...):
            # the binary's dot is mul / add (as interpreted), the contract's is fused (contract v2), both are valid OpDot results
            Nd, Hd, a2 = z["brdf_N"][k].astype(np.float64), z["brdf_H"][k].astype(np.float64), float(z["brdf_rough"][k]) ** 4
            nh2 = max(float(Nd @ Hd), 0.0) ** 2
            cond = 4.0 * nh2 * abs(1.0 - a2) / max(abs(nh2 * (a2 - 1.0) + 1.0), 1e-30)
            assert abs(d - wd) <= 1e-6 * (1.0 + 0.25 * cond) * max(abs(wd), 1e-30) + 1e-30, (k, d, wd, cond)
        else:
            assert not np.isfinite(d) or d > 1e30, (k, d, wd)                 # roughness 0 on the lobe axis: a2 / 0
        assert abs(g - wg) <= 1e-6 * max(abs(wg), 1e-6), (k, g, wg)
        assert np.allclose(np.array(fr[:], np.float32), wf, rtol=2e-6, atol=1e-7), (k, fr[:], wf)


@pytest.fixture(scope="module")
def golden_setup():
    import vulkanhybridraytracer_b200 as vk
    from oracle.oracle_py import Oracle
    from util import GOLDEN_DETAIL
    sc = vk.HostScene(0, detail=GOLDEN_DETAIL)
    o = Oracle(sc)
    yield vk, sc, o
    o.close(); sc.close()


@pytest.mark.parametrize("tag", ["f0", "f3s4"])
def test_shadow_pass_is_shadowraygen_spv(L, golden_setup, tag):
    """SURVEY 8a-4.  shadowRaygen.rgen.spv main() was executed per pixel (traceRayEXT answered by the oracle's any-hit query on the
    ray the SHADER built, shadow.rmiss.spv on a miss).  The oracle's shadow pass must write the same R8 value, and the rays it
    would build (tea seed, sampleDisk by-value seed, origin + 0.1 dir, tmax = distance + 1) must be the shader's rays.  The
    second case (frame 3, 4 samples) shows shadowFactor staying 0 or 1 -- App. A #1 -- because the binary never advances the seed."""
    vk, sc, o = golden_setup
    z = vectors("shadow")
    W, H, frame, samples = [int(v) for v in z[f"{tag}_params"]]
    p = vk.make_params(W, H, frame=frame, shadow_samples=samples)
    cam = sc.camera(W, H)
    g = o.gbuffer(cam, p)
    sh = o.shadow(cam, p, g)
    px = z[f"{tag}_pixels"]
    want = z[f"{tag}_shadow"]
    got = np.stack([sh[l][px[:, 1], px[:, 0]] for l in range(len(sc.lights))], 1)
    same = (got == want).all(1)
    # a differing byte can only be an any-hit answer flipping on a silhouette: the shader's direction and the oracle's differ in
    # the last bits of sin / cos.  None is expected at this size; allow one pixel in a thousand.
    assert same.mean() >= 0.999, (same.mean(), np.argwhere(~same)[:5])
    assert set(np.unique(want)) <= {0, 255}                       # also with 4 samples: all four rays are the same ray
    assert (want == 255).any() and (want == 0).any()
    # the rays
    rays = z[f"{tag}_rays"]
    pos = g.position.astype(np.float32)
    light = T.Light()
    n_checked = 0
    for row in rays[:: max(1, len(rays) // 600)]:
        x, y = int(row[0]), int(row[1])
        P = pos[y, x, :3]
        seed = L.orc_tea(y * W + x, frame)
        best = np.inf
        for l in sc.lights:
            Lv = (l["position"][:3] - P).astype(np.float32)
            dist = np.sqrt(np.float32(Lv[0] * Lv[0] + Lv[1] * Lv[1] + Lv[2] * Lv[2]))
            Ln = (Lv / dist).astype(np.float32)
            for c in range(4):
                light.position[c] = float(l["position"][c]); light.color[c] = float(l["color"][c])
            light.radius = float(l["radius"])
            out = (C.c_float * 3)()
            L.orc_sample_disk(C.byref(light), C.byref(f3(P)), C.byref(f3(Ln)), seed, C.byref(out))
            d = np.array(out[:], np.float32)
            err = max(np.abs(d - row[5:8]).max(), np.abs((P + d * np.float32(0.1)) - row[2:5]).max() / 10.0, abs((dist + 1) - row[8]) / 50.0)
            best = min(best, float(err))
        assert best <= 3e-6, (x, y, best)
        n_checked += 1
    assert n_checked >= 100


def test_resolve_pass_is_hybridraygen_hybridhit_miss_spv(golden_setup):
    """SURVEY 8a-6 / 7 / 8.  hybridRaygen.rgen.spv main() was executed for every pixel of the golden frame; each traceRayEXT ran
    hybridHit.rchit.spv or miss.rmiss.spv main() in the same interpreter on the oracle's closest hit, and hybridHit's shadow rays
    went through the oracle's any-hit query + shadow.rmiss.spv.  Inputs are the oracle's own G-buffer and filtered shadow planes,
    so what is compared is exactly the deferred lighting resolve and the reflection / refraction chain.
    Tolerance: the binaries were evaluated with float64-then-rounded pow / atan / asin / inverse(), the oracle with the
    numerics contract: 2e-5 relative (+ 2e-6 absolute) per channel.  A chain whose bounce lands on another triangle because
    the two directions differ in the last bits would differ by more; none does at this size, one pixel in 500 is allowed."""
    vk, sc, o = golden_setup
    z = vectors("resolve")
    W, H, frame = [int(v) for v in z["params"]]
    p = vk.make_params(W, H, frame=frame)
    cam = sc.camera(W, H)
    o.reset_counters()
    fr = o.hybrid_frame(cam, p)
    px = z["pixels"]
    got = fr["final"][px[:, 1], px[:, 0], :3]
    want = z["final"][:, :3]
    err = np.abs(got - want)
    ok = (err <= 2e-5 * np.abs(want) + 2e-6).all(1)
    assert ok.mean() >= 0.998, (ok.mean(), px[~ok][:5], got[~ok][:5], want[~ok][:5])
    assert err.mean() <= 1e-5 and np.isfinite(want).all()
    # the binaries traced exactly as many rays as the oracle counts: reflection / refraction rays and secondary shadow rays
    assert int(z["ray_counts"][:, 0].sum()) == o.counters.rays[2] and int(z["ray_counts"][:, 1].sum()) == o.counters.rays[3]
    # every kind of pixel is in the frame: sky (albedo passthrough), diffuse, mirror and glass chains of several bounces
    ids = fr["g"].hit_ids[..., 0]
    assert (ids < 0).any() and z["ray_counts"][:, 0].max() >= 3 and (z["ray_counts"][:, 0] == 0).sum() > 1000
    modes = sc.materials["shadingMetallicRoughness"][:, 0][sc.instances["material_id"][np.maximum(ids, 0)]]
    assert {0.0, 3.0, 4.0} <= set(np.unique(modes[ids >= 0]).tolist())


@pytest.mark.parametrize("frame", [0, 5])
def test_denoise_pass_is_blur_comp_spv(frame):
    """SURVEY 8a-5.  blur.comp.spv main() at its compiled-in 1700 x 900 x 3 lights for 269 invocations (corners, borders, random),
    a reset frame (3x3 clamp + history fetch at pixel + motion.x * (W, H), out-of-range reads = 0) and an accumulation frame
    (mix(old, new, 1 / (frame + 1))).  The oracle's filter must store the same R8 value; a mix() landing within float rounding of
    x.5 / 255 may round the other way (the oracle evaluates the same expression: none seen)."""
    import vulkanhybridraytracer_b200 as vk
    from oracle.oracle_py import GBuffer, Oracle
    from util import blur_inputs
    W, H, Lc, shadow, hist, motion, pix = blur_inputs()
    z = vectors("blur")
    assert np.array_equal(z["pixels"], pix)
    sc = vk.HostScene(0, detail=(2, 2, 2, 2))                # the default scene has three lights: the filter only needs the count
    o = Oracle(sc)
    assert o.n_lights == Lc
    g = GBuffer(W, H); g.motion[...] = motion
    p = vk.make_params(W, H, frame=frame, denoise_frame=frame, denoise_cover_remainder=1)
    den = o.denoise(p, g, np.ascontiguousarray(shadow), np.ascontiguousarray(hist))
    got = np.stack([den[l][pix[:, 1], pix[:, 0]] for l in range(Lc)], 1)
    want = z[f"f{frame}_denoised"]
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 1 and (got == want).mean() >= 0.995, (got != want).sum()
    assert len(np.unique(want)) > 10
    o.close(); sc.close()


def test_post_pass_is_postfragment_spv(L):
    """SURVEY 8a-10.  postFragment.frag.spv main() on all 256 possible 8-bit inputs per channel: max(rgb, 0.001) ^ (1 / 2.2), alpha
    passed through.  The oracle's post pass stores unorm8 of the same: equal bytes, +-1 where pow() lands within an ulp of a
    rounding boundary."""
    import vulkanhybridraytracer_b200 as vk
    from oracle.oracle_py import Oracle
    z = vectors("post")
    tex = z["texels"][0]                                       # 256 x RGBA8
    want = np.rint(np.clip(z["frag_color"], 0, 1) * 255).astype(int)
    # the oracle's post pass takes the float frame: feed it the texel values so that its own 8-bit store reproduces them
    final = (tex.astype(np.float32) / np.float32(255)).reshape(1, 256, 4)
    sc = vk.HostScene(1, detail=(2, 2, 2, 2)); o = Oracle(sc)
    got = o.post(vk.make_params(256, 1), np.ascontiguousarray(final)).reshape(256, 4).astype(int)
    o.close(); sc.close()
    assert np.abs(got - want).max() <= 1 and (got == want).mean() >= 0.99, np.argwhere(got != want)[:5]
    assert (got[:, 3] == 255).all() and got[0, 0] == want[0, 0] > 0          # the 0.001 floor: black does not stay 0


@pytest.mark.parametrize("target", range(7))
def test_deferred_pass_is_deferred_frag_spv(golden_setup, target):
    """SURVEY 8a-12.  deferred.frag.spv main() on the oracle's G-buffer: the full shading path (target 0, every pixel) and the six
    debug views (deferred.frag:50-73).  2e-5 relative + 2e-6 absolute, as for the resolve pass."""
    vk, sc, o = golden_setup
    z = vectors("deferred")
    W, H = [int(v) for v in z["params"]]
    cam = sc.camera(W, H)
    g = o.gbuffer(cam, vk.make_params(W, H))
    d = o.deferred(cam, vk.make_params(W, H, mode=vk.MODE_DEFERRED, debug_target=target), g)
    px = z[f"t{target}_pixels"]; want = z[f"t{target}_color"]
    got = d[px[:, 1], px[:, 0]]
    nch = 4
    err = np.abs(got[:, :nch] - want[:, :nch])
    ok = (err <= 2e-5 * np.abs(want[:, :nch]) + 2e-6).all(1)
    assert ok.mean() >= 0.998, (target, ok.mean(), px[~ok][:5], got[~ok][:5], want[~ok][:5])


def test_gbuffer_pass_is_basic_vert_and_geometry_frag_spv(golden_setup):
    """SURVEY 8a-3.  basic.vert.spv main() ran on the three vertices of the triangle under each covered pixel, its outputs were
    interpolated with the hit's barycentrics and passed BY LOCATION to geometry_shader.frag.spv main() -- with a camera that moved
    since the previous frame, so the motion plane is exercised (and its previous - current orientation: basic.vert writes ndcPrev to
    location 4 and ndc to 5, the fragment shader reads 4 as inNdc and 5 as inNdcPrev).  The oracle's ray-cast G-buffer must store
    the same fp16 / unorm8 values: exact for most, +-1 ulp of the storage format where interpolate-then-transform and
    transform-then-interpolate round differently.  Normal-mapped pixels (dFdx / dFdy from neighbour rays) get 2e-3 on the normal."""
    import vulkanhybridraytracer_b200 as vk2
    from oracle.oracle_py import Oracle
    from util import GOLDEN_DETAIL
    sc = vk2.HostScene(0, detail=GOLDEN_DETAIL)            # own scene: the camera is moved
    o = Oracle(sc)
    z = vectors("gbuffer")
    W, H = [int(v) for v in z["params"]]
    cam0 = sc.camera(W, H)
    sc.move_camera(3, 6.0); sc.rotate_camera(12.0, 4.0)
    cam = sc.camera(W, H, prev=cam0)
    g = o.gbuffer(cam, vk2.make_params(W, H))
    px = z["pixels"]; X, Y = px[:, 0], px[:, 1]
    assert len(px) > 500

    def half_ulps(got_f16, want_f32):
        a = got_f16.view(np.int16).astype(np.int32); b = want_f32.astype(np.float16).view(np.int16).astype(np.int32)
        a = np.where(a < 0, -(a & 0x7fff), a); b = np.where(b < 0, -(b & 0x7fff), b)
        return np.abs(a - b)
    pos = half_ulps(g.position[Y, X], z["position"])
    assert pos.max() <= 1 and (pos == 0).mean() > 0.97, (pos.max(), (pos == 0).mean())
    plain = ~z["normal_mapped"]
    nrm = half_ulps(g.normal[Y, X], z["normal"])
    assert nrm[plain].max() <= 1 and (nrm[plain] == 0).mean() > 0.97
    if (~plain).any():
        assert np.abs(g.normal[Y, X][~plain].astype(np.float32) - z["normal"][~plain]).max() <= 2e-3
    mot = half_ulps(g.motion[Y, X], z["motion"])
    assert mot.max() <= 1 and (mot == 0).mean() > 0.95, (mot.max(), (mot == 0).mean())
    assert (np.abs(z["motion"] - 0.5) > 1e-3).mean() > 0.5                     # the camera did move
    for name in ("albedo", "material", "emissive"):
        want = np.rint(np.clip(z[name], 0, 1) * 255).astype(int)
        got = getattr(g, name)[Y, X][:, :3].astype(int)
        assert np.abs(got - want).max() <= 1 and (got == want).mean() > 0.995, name
    # the sign of the stored vector: re-projecting with it must land on the same surface in the previous frame
    g0 = o.gbuffer(cam0, vk2.make_params(W, H))
    mo = z["motion"] * 2 - 1
    lx = np.floor(X + 0.5 + 0.5 * mo[:, 0] * W).astype(int); ly = np.floor(Y + 0.5 + 0.5 * mo[:, 1] * H).astype(int)
    inside = (lx >= 0) & (ly >= 0) & (lx < W) & (ly < H)
    same = g0.hit_ids[ly[inside], lx[inside], 0] == g.hit_ids[Y[inside], X[inside], 0]
    wrong = np.floor(X + 0.5 - 0.5 * mo[:, 0] * W).astype(int), np.floor(Y + 0.5 - 0.5 * mo[:, 1] * H).astype(int)
    ins2 = (wrong[0] >= 0) & (wrong[1] >= 0) & (wrong[0] < W) & (wrong[1] < H)
    same2 = g0.hit_ids[wrong[1][ins2], wrong[0][ins2], 0] == g.hit_ids[Y[ins2], X[ins2], 0]
    # (most pixels show the floor, which is "the same surface" for either sign: the objects on it tell the two apart)
    obj = g.hit_ids[Y, X, 0] > 0
    same_o = g0.hit_ids[ly[inside & obj], lx[inside & obj], 0] == g.hit_ids[Y[inside & obj], X[inside & obj], 0]
    same2_o = g0.hit_ids[wrong[1][ins2 & obj], wrong[0][ins2 & obj], 0] == g.hit_ids[Y[ins2 & obj], X[ins2 & obj], 0]
    assert same.mean() > 0.95 and same.mean() > same2.mean() and same_o.mean() > 0.85 and same_o.mean() > same2_o.mean() + 0.1, \
        (same.mean(), same2.mean(), same_o.mean(), same2_o.mean())
    o.close(); sc.close()


@pytest.mark.parametrize("tag", ["f0s1", "f2s3"])
def test_raytraced_mode_is_raygen_closesthit_miss_spv(golden_setup, tag):
    """SURVEY 8a-11 (the Raytraced mode of BASELINE configs[2]).  raygen.rgen.spv main() per pixel -- NSAMPLES is 1 in the shipped
    binary -- with closesthit.rchit.spv / miss.rmiss.spv on the oracle's closest hit and shadow.rmiss.spv behind the any-hit
    shadow rays; second case: frame 2 (accumulation over a seeded previous image, a = 1 / 3) with 3 shadow samples per light.
    Same tolerance as the resolve pass; the ray counts must agree exactly."""
    vk, sc, o = golden_setup
    z = vectors("raytrace")
    W, H, frame, samples = [int(v) for v in z[f"{tag}_params"]]
    p = vk.make_params(W, H, mode=vk.MODE_RAYTRACING, spp=1, shadow_samples=samples, frame=frame, faithful_rng=1)
    cam = sc.camera(W, H)
    prev = z[f"{tag}_previous"] if frame else None
    o.reset_counters()
    a = o.raytrace(cam, p, None if prev is None else prev.copy())
    px = z[f"{tag}_pixels"]; want = z[f"{tag}_color"][:, :3]
    got = a[px[:, 1], px[:, 0], :3]
    err = np.abs(got - want)
    ok = (err <= 2e-5 * np.abs(want) + 2e-6).all(1)
    assert ok.mean() >= 0.998, (ok.mean(), px[~ok][:5], got[~ok][:5], want[~ok][:5])
    if len(px) == W * H:
        assert int(z[f"{tag}_ray_counts"][:, 0].sum()) == o.counters.rays[2] and int(z[f"{tag}_ray_counts"][:, 1].sum()) == o.counters.rays[3]


def test_spirv_interpreter_on_a_hand_assembled_module(tmp_path):
    """The interpreter itself, without the reference: a module assembled word by word here -- uint f(uint* a, uint* b) { uint s = 0;
    for (uint i = 0; i < *b; i++) s += *a * i; return s; } and float g(float* x) { return sqrt(*x) * 0.5 } -- must wrap like
    32-bit hardware and round like fp32."""
    import struct
    from oracle.spirv_interp import Interpreter, Module, U32, F32

    def ins(op, *ops):
        return [((len(ops) + 1) << 16) | op] + list(ops)

    def s(text):
        b = text.encode() + b"\0"
        b += b"\0" * (-len(b) % 4)
        return list(struct.unpack("<%dI" % (len(b) // 4), b))
    VOID, UINT, FLOAT, BOOL, PU, PF, FN_U, FN_F, C0, C1, CH, EXT = 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12
    F, A, B, L0, S, I, LH, LB, LC, LM, LCND = 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30
    t = list(range(40, 60))
    G, X, GL = 70, 71, 72
    w = [0x07230203, 0x00010500, 0, 100, 0]
    w += ins(17, 1)                                     # OpCapability Shader
    w += ins(11, EXT, *s("GLSL.std.450"))
    w += ins(14, 0, 1)                                  # OpMemoryModel Logical GLSL450
    w += ins(5, F, *s("f(u1;u1;")) + ins(5, G, *s("g(f1;"))
    w += ins(19, VOID) + ins(21, UINT, 32, 0) + ins(22, FLOAT, 32) + ins(20, BOOL)
    w += ins(32, PU, 7, UINT) + ins(32, PF, 7, FLOAT) + ins(33, FN_U, UINT, PU, PU) + ins(33, FN_F, FLOAT, PF)
    w += ins(43, UINT, C0, 0) + ins(43, UINT, C1, 1) + ins(43, FLOAT, CH, struct.unpack("<I", struct.pack("<f", 0.5))[0])
    # uint f(a, b)
    w += ins(54, UINT, F, 0, FN_U) + ins(55, PU, A) + ins(55, PU, B) + ins(248, L0)
    w += ins(59, PU, S, 7) + ins(59, PU, I, 7) + ins(62, S, C0) + ins(62, I, C0) + ins(249, LH)
    w += ins(248, LH) + ins(246, LM, LC, 0) + ins(249, LCND)
    w += ins(248, LCND) + ins(61, UINT, t[0], I) + ins(61, UINT, t[1], B) + ins(176, BOOL, t[2], t[0], t[1]) + ins(250, t[2], LB, LM)
    w += ins(248, LB) + ins(61, UINT, t[3], A) + ins(61, UINT, t[4], I) + ins(132, UINT, t[5], t[3], t[4]) + ins(61, UINT, t[6], S)
    w += ins(128, UINT, t[7], t[6], t[5]) + ins(62, S, t[7]) + ins(249, LC)
    w += ins(248, LC) + ins(61, UINT, t[8], I) + ins(128, UINT, t[9], t[8], C1) + ins(62, I, t[9]) + ins(249, LH)
    w += ins(248, LM) + ins(61, UINT, t[10], S) + ins(254, t[10]) + ins(56)
    # float g(x)
    w += ins(54, FLOAT, G, 0, FN_F) + ins(55, PF, X) + ins(248, GL) + ins(61, FLOAT, t[11], X) + ins(12, FLOAT, t[12], EXT, 31, t[11])
    w += ins(133, FLOAT, t[13], t[12], CH) + ins(254, t[13]) + ins(56)
    path = tmp_path / "m.spv"
    path.write_bytes(struct.pack("<%dI" % len(w), *w))
    it = Interpreter(Module(str(path)))
    for a, b in ((3, 5), (0xdeadbeef, 1000), (7, 0)):
        want = sum((a * i) & 0xffffffff for i in range(b)) & 0xffffffff
        assert int(it.call("f", U32(a), U32(b))[0]) == want
    for x in (2.0, 1e-20, 3.4e38, 0.0):
        got = it.call("g", F32(x))[0]
        assert np.float32(got).view(np.uint32) == (np.sqrt(np.float32(x)) * np.float32(0.5)).view(np.uint32)


def test_light_edge_cases_are_the_shader_binaries():
    """Branches the default lights never take, through shadowRaygen.rgen.spv and hybridRaygen.rgen.spv + hybridHit.rchit.spv:
    a directional light (pos.w < 0), a point light most of the scene lies beyond the maxDistance of (unshadowed in the shadow pass,
    shadowRaygen.rgen:81-83, but shadowed in the hit shader, hybridHit.rchit:115-125 -- App. A #4), and a zero-intensity light
    (attenuation 0 / the irradiance branch of the mirror and glass terms, App. A #8); frame 1, two shadow samples."""
    import vulkanhybridraytracer_b200 as vk2
    from oracle.oracle_py import Oracle
    from util import GOLDEN_DETAIL
    z = vectors("variants")
    W, H, frame, samples = [int(v) for v in z["params"]]
    sc = vk2.HostScene(0, detail=GOLDEN_DETAIL); o = Oracle(sc)
    lights = sc.lights.copy()
    for k in ("position", "color", "radius"):
        lights[k] = z["lights"][k]
    assert lights["position"][0][3] < 0 and lights["color"][2][3] == 0
    o.set_lights(lights)
    p = vk2.make_params(W, H, frame=frame, shadow_samples=samples)
    fr = o.hybrid_frame(sc.camera(W, H), p)
    px = z["pixels"]
    got = np.stack([fr["shadow"][l][px[:, 1], px[:, 0]] for l in range(3)], 1)
    assert (got == z["shadow"]).all(1).mean() >= 0.998
    assert (z["shadow"][:, 1] == 255).mean() > 0.25                # light 1: "beyond maxDistance" counts as lit in the shadow pass
    want = z["final"][:, :3]; have = fr["final"][px[:, 1], px[:, 0], :3]
    ok = (np.abs(have - want) <= 2e-5 * np.abs(want) + 2e-6).all(1)
    assert ok.mean() >= 0.996, (ok.mean(), px[~ok][:5], have[~ok][:5], want[~ok][:5])
    o.close(); sc.close()


@pytest.mark.parametrize("index", [1, 2])
def test_resolve_on_the_other_scenes_is_the_shader_binaries(index):
    """hybridRaygen.rgen.spv + hybridHit.rchit.spv + miss.rmiss.spv on the Cornell scene (scene.cpp:162-272, one light) and on the
    instanced stress scene (rotated, non-uniformly scaled instances: the shader derives the normal matrix from inverse(model)
    itself, hybridHit.rchit:50, the oracle takes the host-supplied inverse), frame 2."""
    import vulkanhybridraytracer_b200 as vk2
    from oracle.oracle_py import Oracle
    z = vectors("scenes")
    W, H, frame = [int(v) for v in z[f"s{index}_params"]]
    sc = vk2.HostScene(index, detail=(3, 2, 3, 14)); o = Oracle(sc)
    fr = o.hybrid_frame(sc.camera(W, H), vk2.make_params(W, H, frame=frame))
    px = z[f"s{index}_pixels"]; want = z[f"s{index}_final"][:, :3]
    have = fr["final"][px[:, 1], px[:, 0], :3]
    ok = (np.abs(have - want) <= 2e-5 * np.abs(want) + 2e-6).all(1)
    assert ok.mean() >= 0.998, (index, ok.mean(), px[~ok][:5], have[~ok][:5], want[~ok][:5])
    assert (fr["g"].hit_ids[..., 0] >= 0).mean() > 0.2
    o.close(); sc.close()
