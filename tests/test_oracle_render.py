"""CPU tests of the oracle itself: BVH == brute force, stage semantics of App. A, golden fixtures, libm twin."""
import os

import numpy as np
import pytest

import vulkanhybridraytracer_b200 as vk
from vulkanhybridraytracer_b200 import types as T
from oracle.oracle_py import Oracle

from util import random_rays, GOLDEN_DIR, golden_case

W, H = 96, 54


def test_bvh_equals_brute_force(small_scene, small_oracle):
    rays = random_rays(small_scene, 4000, seed=1)
    a = small_oracle.trace_rays(rays, brute_force=True)
    b = small_oracle.trace_rays(rays, brute_force=False)
    assert (a["instance"] >= 0).mean() > 0.3
    assert np.array_equal(a["instance"], b["instance"]) and np.array_equal(a["primitive"], b["primitive"])
    assert np.array_equal(a["t"].view(np.uint32), b["t"].view(np.uint32))
    sa = small_oracle.trace_rays(rays, any_hit=True, brute_force=True)
    sb = small_oracle.trace_rays(rays, any_hit=True, brute_force=False)
    assert np.array_equal(sa["t"] >= 0, sb["t"] >= 0)
    # any-hit agrees with closest-hit about occlusion
    assert np.array_equal(sa["t"] >= 0, a["t"] >= 0)


def test_closest_hit_tie_rule(small_scene):
    """Two coincident quads: the lower instance id wins (stated tie rule)."""
    sc = small_scene
    verts, idx = sc.meshes[0]          # the quad
    inst = np.zeros(2, T.instance_dtype)
    eye = np.eye(4, dtype=np.float32).T.reshape(-1)
    for k in range(2):
        inst[k]["mesh_id"] = 0; inst[k]["index_count"] = 6; inst[k]["material_id"] = 0
        inst[k]["model"] = eye; inst[k]["inv_model"] = eye
    o = Oracle(sc)
    o.set_instances(inst)
    rays = np.zeros(1, T.ray_dtype)
    rays["origin"] = [0.25, 0.1, 3.0]; rays["dir"] = [0, 0, -1]; rays["tmin"] = 0.001; rays["tmax"] = 100
    for brute in (True, False):
        h = o.trace_rays(rays, brute_force=brute)
        assert h["instance"][0] == 0 and h["t"][0] == 3.0
    # the diagonal is shared by both triangles of the quad: lower primitive id wins
    rays["origin"] = [0.5, 0.5, 3.0]
    h = o.trace_rays(rays, brute_force=True)
    assert h["instance"][0] == 0 and h["primitive"][0] == 0


def test_hybrid_frame_semantics(small_scene, small_oracle):
    sc, o = small_scene, small_oracle
    cam = sc.camera(W, H); p = vk.make_params(W, H)
    o.reset_counters()
    fr = o.hybrid_frame(cam, p)
    g = fr["g"]
    ids = g.hit_ids[..., 0]
    bg = ids < 0
    assert 0.2 < bg.mean() < 0.8
    # skybox.frag: background position / normal / material / emissive are all zero, motion = 0
    assert not g.position[bg].any() and not g.normal[bg].any() and not g.material[bg].any() and not g.motion[bg].any()
    # geometry: normal.w = 1, pos.w = material index of the instance hit
    assert (g.normal[~bg][:, 3] == 1).all()
    assert np.array_equal(g.position[..., 3][~bg].astype(np.int32), sc.instances["material_id"][ids[~bg]])
    # static camera -> motion exactly 0 -> stored 0.5
    assert (g.motion[~bg] == np.float16(0.5)).all()
    # S = 1 and faithful RNG: shadow factor is binary (App. A #1); glass pixels are fully lit (#5)
    assert set(np.unique(fr["shadow"])) <= {0, 255}
    glass = (ids == 2)
    assert glass.any() and (fr["shadow"][:, glass] == 255).all()
    # first frame (frame = 0): denoised = mix(clamp(hist=0), center, 0.4): lit pixels with a lit 3x3 -> 255*... check the formula on one plane
    sh = fr["shadow"][0].astype(np.float32) / 255.0
    pad = np.pad(sh, 1, constant_values=0.0)
    nb = np.stack([pad[1 + dy:1 + dy + H, 1 + dx:1 + dx + W] for dy in (-1, 0, 1) for dx in (-1, 0, 1)])
    old = np.clip(0.0, nb.min(0), nb.max(0))
    expect = np.rint((old * np.float32(0.6) + sh * np.float32(0.4)) * 255).astype(np.uint8)
    geo = ~bg     # background pixels reproject out of bounds as well, so the same formula holds everywhere
    assert np.abs(fr["denoised"][0].astype(int) - expect.astype(int)).max() <= 1
    # resolve: background pixels show the gamma-decoded sky (hybridRaygen.rgen:182-183)
    sky = (g.albedo[bg][:, :3].astype(np.float64) / 255.0) ** 2.2
    assert np.abs(fr["final"][bg][:, :3] - sky).max() < 1e-5
    assert np.isfinite(fr["final"]).all()
    # ray accounting: one primary per pixel, shadow rays only where NdotL > 0, reflections only from mirror / glass pixels
    c = list(o.counters.rays)
    assert c[0] == W * H and 0 < c[1] <= 3 * W * H and c[2] > 0 and c[3] > 0
    mirror_or_glass = np.isin(ids, [1, 2, 3])
    assert mirror_or_glass.sum() <= c[2] <= 8 * mirror_or_glass.sum()


def test_temporal_accumulation_and_remainder(small_scene, small_oracle):
    sc, o = small_scene, small_oracle
    Wr, Hr = 100, 60           # not multiples of 16
    cam = sc.camera(Wr, Hr)
    p0 = vk.make_params(Wr, Hr, frame=0)
    f0 = o.hybrid_frame(cam, p0)
    p1 = vk.make_params(Wr, Hr, frame=1)
    f1 = o.hybrid_frame(cam, p1, history=f0["denoised"])
    # blur.comp:41-48: out = mix(old, center, 1/(frame+1))
    exp = np.rint((f0["denoised"].astype(np.float32) / 255 * np.float32(0.5) + f1["shadow"].astype(np.float32) / 255 * np.float32(0.5)) * 255)
    assert np.abs(f1["denoised"].astype(int) - exp.astype(int)).max() <= 1
    # frame changes the seed -> a different penumbra sample somewhere
    assert (f0["shadow"] != f1["shadow"]).any()
    # faithful dispatch (renderer.cpp:2672): W/16 x H/16 groups, remainder pixels keep the history
    pf = vk.make_params(Wr, Hr, frame=1, denoise_cover_remainder=0)
    hist = np.full_like(f0["denoised"], 77)
    d = o.denoise(pf, f1["g"], f1["shadow"], hist)
    assert (d[:, :, 96:] == 77).all() and (d[:, 48:, :] == 77).all() and (d[:, :48, :96] != 77).any()


def test_shadow_samples_and_rng_modes(small_scene, small_oracle):
    sc, o = small_scene, small_oracle
    cam = sc.camera(W, H)
    g = o.gbuffer(cam, vk.make_params(W, H))
    s1 = o.shadow(cam, vk.make_params(W, H, shadow_samples=1), g)
    s4 = o.shadow(cam, vk.make_params(W, H, shadow_samples=4), g)
    assert np.array_equal(s1, s4)              # seed by value: all samples identical (App. A #1)
    a4 = o.shadow(cam, vk.make_params(W, H, shadow_samples=4, faithful_rng=0), g)
    assert set(np.unique(a4)) - {0, 64, 128, 191, 255} == set() and len(np.unique(a4)) > 2


def test_raytrace_and_deferred_modes(small_scene, small_oracle):
    sc, o = small_scene, small_oracle
    cam = sc.camera(W, H)
    p = vk.make_params(W, H, mode=vk.MODE_RAYTRACING, spp=2, shadow_samples=2, faithful_rng=0)
    a0 = o.raytrace(cam, p)
    assert np.isfinite(a0).all() and a0[..., :3].mean() > 0.05
    p1 = vk.make_params(W, H, mode=vk.MODE_RAYTRACING, spp=2, shadow_samples=2, faithful_rng=0, frame=1)
    a1 = o.raytrace(cam, p1, accum=a0)
    n1 = o.raytrace(cam, p1, accum=np.zeros_like(a0))
    assert np.allclose(a1[..., :3], 0.5 * a0[..., :3] + n1[..., :3], atol=1e-5)   # raygen.rgen:76-79 running average
    g = o.gbuffer(cam, vk.make_params(W, H))
    d = o.deferred(cam, vk.make_params(W, H, mode=vk.MODE_DEFERRED), g)
    assert np.isfinite(d).all()
    for target, plane in ((3, g.albedo), (5, g.material), (6, g.emissive)):
        dbg = o.deferred(cam, vk.make_params(W, H, mode=vk.MODE_DEFERRED, debug_target=target), g)
        assert np.allclose(dbg[..., :3], plane[..., :3].astype(np.float32) / 255.0, atol=1e-7)


@pytest.mark.parametrize("name", ["hybrid_f0", "hybrid_f1", "raytrace", "deferred"])
def test_golden_fixtures(name):
    """tests/golden/*.npz were produced by tests/golden/make_golden.py with this oracle; they pin it against silent drift
    (the reference itself ships no vectors)."""
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    assert os.path.exists(path), "run tests/golden/make_golden.py"
    ref = np.load(path)
    out = golden_case(name)
    for k in ref.files:
        assert np.array_equal(ref[k], out[k]), (name, k)


def test_libm_twin_is_within_north_star_tolerance(small_scene, small_oracle):
    """Independence check: the same oracle compiled against glibc's sinf/cosf/atan2f/asinf/powf instead of vkp_detmath.h."""
    sc = small_scene
    ol = Oracle(sc, libm=True)
    assert ol.L.orc_uses_libm() == 1 and small_oracle.L.orc_uses_libm() == 0
    cam = sc.camera(W, H); p = vk.make_params(W, H)
    a = small_oracle.hybrid_frame(cam, p); b = ol.hybrid_frame(cam, p)
    assert np.array_equal(a["g"].hit_ids, b["g"].hit_ids)
    assert np.array_equal(a["g"].position.view(np.uint16), b["g"].position.view(np.uint16))
    d = np.abs(a["final"][..., :3] - b["final"][..., :3])
    flips = (a["shadow"] != b["shadow"]).any(0)
    assert flips.mean() <= 1e-3                       # epsilon-tie penumbra flips only
    assert d.mean() <= 1e-3
    assert d[~flips].max() <= 1e-2


def test_normal_mapping_with_ray_differentials():
    """geometry_shader.frag:28-63: the helmet material carries a normal map; the cotangent frame comes from the position
    derivatives (neighbour camera rays on the hit triangle's plane).  Only that instance changes, the result is a unit
    vector near the interpolated normal, and a flat map (128, 128, 255) gives the interpolated normal back."""
    detail = (4, 3, 4, 6)
    w, h = 240, 136
    sc = vk.HostScene(0, detail=detail)
    cam = sc.camera(w, h); p = vk.make_params(w, h)
    assert sc.materials["textures"][4][1] >= 0
    g = Oracle(sc).gbuffer(cam, p)
    helmet = g.hit_ids[..., 0] == 5
    assert helmet.sum() > 100
    n = g.normal.astype(np.float32)[..., :3] * 2 - 1
    flat = vk.HostScene(0, detail=detail)
    tid = int(flat.materials["textures"][4][1])
    flat.textures[tid][...] = np.array([128, 128, 255, 255], np.uint8)
    gf = Oracle(flat).gbuffer(cam, p)
    off = vk.HostScene(0, detail=detail)
    off.materials["textures"][4][1] = -1
    g0 = Oracle(off).gbuffer(cam, p)
    n0 = g0.normal.astype(np.float32)[..., :3] * 2 - 1
    nf = gf.normal.astype(np.float32)[..., :3] * 2 - 1
    assert np.array_equal(g.normal[~helmet], g0.normal[~helmet])              # nobody else has a normal map
    for k in ("position", "albedo", "material", "emissive", "hit_ids"):
        assert np.array_equal(getattr(g, k), getattr(g0, k))
    assert (g.normal[helmet] != g0.normal[helmet]).any(axis=1).mean() > 0.9    # the map does perturb
    assert np.abs(np.linalg.norm(n[helmet], axis=1) - 1).max() < 2e-3
    cosang = (n[helmet] * n0[helmet]).sum(axis=1)
    assert cosang.min() > 0.5 and cosang.mean() > 0.9
    # flat map: normal_pixel = (1/127, 1/127, 1): within a degree of the interpolated normal
    assert ((nf[helmet] * n0[helmet]).sum(axis=1) > 0.998).mean() > 0.95          # (fp16 storage: +-5e-4 per component)


def test_gltf_viewer_scene_renders_through_the_oracle(tmp_path):
    """a multi-primitive, textured glTF prefab (host/gltf.cpp) goes through the same flat buffers as the
    built-in scenes: one TLAS instance per primitive with its own index range of the shared mesh, BVH == brute force,
    and the G-buffer shows the prefab's textures (albedo, emissive, metallic-roughness, normal map)."""
    from util import make_test_glb
    info = make_test_glb(tmp_path / "model.glb")
    sc = vk.HostScene.from_prefab(tmp_path / "model.glb")
    assert len(sc.instances) == 4 and sc.n_triangles == 2 + info["triangles"]
    assert sc.instances["first_index"][1:].tolist() == [6, 6 + int(sc.instances["index_count"][1]), 0]      # child (quad) was loaded first
    o = Oracle(sc)
    rays = random_rays(sc, 3000, seed=4)
    a = o.trace_rays(rays, brute_force=True); b = o.trace_rays(rays, brute_force=False)
    assert np.array_equal(a["instance"], b["instance"]) and np.array_equal(a["primitive"], b["primitive"])
    assert np.array_equal(a["t"].view(np.uint32), b["t"].view(np.uint32)) and set(np.unique(a["instance"])) >= {0, 1, 2, 3}
    w, h = 160, 90
    cam = sc.camera(w, h)
    fr = o.hybrid_frame(cam, vk.make_params(w, h))
    g = fr["g"]
    ids = g.hit_ids[..., 0]
    assert all((ids == k).sum() > 20 for k in range(4))
    textured = (ids == 1) | (ids == 3)                 # instance order is node_to_instance's: root's two primitives, then the child's quad
    plain = ids == 2
    base = info["textures"][0].reshape(-1, 3)
    texels = {tuple(t) for t in base.tolist()}
    seen = {tuple(t) for t in g.albedo[textured][:, :3].tolist()}
    assert len(seen) > 20 and seen <= texels           # baseColorFactor 1: albedo is exactly a texel of the base-colour map (NEAREST)
    assert len({tuple(t) for t in g.albedo[plain][:, :3].tolist()}) == 1          # untextured primitive: one flat colour
    assert (g.emissive[textured][:, :3] != 0).any() and (g.emissive[plain][:, :3] == 0).all()
    assert len(np.unique(g.material[textured][:, 1])) > 5 and len(np.unique(g.material[plain][:, 1])) == 1
    assert np.isfinite(fr["final"]).all() and fr["final"][..., :3].max() > 0.05
    o.close(); sc.close()


def test_reproject_fix_fetches_the_previous_frames_pixel(small_scene, small_oracle):
    """vkp_params.reproject_fix.  blur.comp:52-53 fetches the history at pixel + motion.x * (W, H) -- twice the
    right distance, x for both axes; the flag fetches floor(pixel + 0.5 + 0.5 * motion * (W, H)) instead (motion = previous - current).  After a camera move the
    re-projected history must (a) be the value the same surface point had in the previous frame and (b) bring the reset
    frame closer to the converged shadow than the reference's fetch does.  Default (0) stays the reference's behaviour."""
    sc = vk.HostScene(0, detail=(4, 3, 4, 6))
    o = small_oracle
    w, h = 160, 90
    cam0 = sc.camera(w, h)
    hist = None
    for f in range(8):                                          # converge the soft shadows at the first camera
        fr0 = o.hybrid_frame(cam0, vk.make_params(w, h, frame=f), history=hist); hist = fr0["denoised"]
    sc.move_camera(3, 6.0); sc.rotate_camera(12.0, 4.0)         # strafe + look: a few pixels of image motion
    cam1 = sc.camera(w, h, prev=cam0)
    truth = None
    for f in range(8):
        t = o.hybrid_frame(cam1, vk.make_params(w, h, frame=f), history=truth); truth = t["denoised"]
    p_ref = vk.make_params(w, h, frame=0)
    p_fix = vk.make_params(w, h, frame=0, reproject_fix=1)
    a = o.hybrid_frame(cam1, p_ref, history=hist)
    b = o.hybrid_frame(cam1, p_fix, history=hist)
    g = b["g"]
    assert np.array_equal(a["shadow"], b["shadow"]) and np.array_equal(a["g"].motion, g.motion)      # only the filter changes
    # (a) re-projection lands on the same surface: previous-frame instance id at the fetched pixel == current id
    mo = g.motion.astype(np.float32) * 2 - 1
    ys, xs = np.mgrid[0:h, 0:w]
    lx = np.floor((xs + np.float32(0.5)) + (mo[..., 0] * np.float32(0.5)) * np.float32(w)).astype(int)
    ly = np.floor((ys + np.float32(0.5)) + (mo[..., 1] * np.float32(0.5)) * np.float32(h)).astype(int)
    inside = (lx >= 0) & (ly >= 0) & (lx < w) & (ly < h) & (g.hit_ids[..., 0] >= 0)
    same = fr0["g"].hit_ids[ly[inside], lx[inside], 0] == g.hit_ids[..., 0][inside]
    assert inside.mean() > 0.3 and same.mean() > 0.97
    moved = (np.abs(lx - xs) + np.abs(ly - ys))[inside]
    assert moved.max() >= 2                                      # the move is not a no-op
    # (b) closer to the converged result, on the pixels whose history is available
    err_ref = np.abs(a["denoised"].astype(int) - truth.astype(int))[:, inside].mean()
    err_fix = np.abs(b["denoised"].astype(int) - truth.astype(int))[:, inside].mean()
    assert err_fix < err_ref, (err_fix, err_ref)
    # static camera: motion == 0 on geometry, the two fetches coincide there (sky pixels store motion 0 = decoded -1,
    # skybox.frag:19; their shadow value is never used, hybridRaygen.rgen:182-183) and the image is the same
    s0 = o.hybrid_frame(cam0, vk.make_params(w, h, frame=0), history=hist)
    s1 = o.hybrid_frame(cam0, vk.make_params(w, h, frame=0, reproject_fix=1), history=hist)
    geo = s0["g"].hit_ids[..., 0] >= 0
    assert np.array_equal(s0["denoised"][:, geo], s1["denoised"][:, geo]) and np.array_equal(s0["final"].view(np.uint32), s1["final"].view(np.uint32))
    sc.close()


def test_atrous_denoiser_smooths_noise_and_keeps_edges(small_scene, small_oracle):
    """vkp_params.denoiser = 1 (no reference counterpart): the a-trous iterations are a weighted mean with non-negative
    weights, so (a) uniform planes pass through unchanged, (b) every output lies between the minimum and the maximum of the input,
    (c) noise on a smooth surface shrinks, (d) background pixels are copied, and the default (0) is untouched by the new code."""
    sc, o = small_scene, small_oracle
    w, h = 128, 72
    cam = sc.camera(w, h)
    p0 = vk.make_params(w, h, frame=0)
    p1 = vk.make_params(w, h, frame=0, denoiser=1)
    g = o.gbuffer(cam, p0)
    L = len(sc.lights)
    nrm = g.normal.view(np.float16).reshape(h, w, 4)
    surface = nrm[..., 3] != 0
    assert surface.mean() > 0.3 and (~surface).any()
    zero_hist = np.zeros((L, h, w), np.uint8)
    for v in (0, 255, 77):
        flat = np.full((L, h, w), v, np.uint8)
        out = o.denoise(p1, g, flat, zero_hist)
        # reset frame: blur.comp's temporal stage with a clamped history gives mix(v, v, 0.4) = v on a uniform plane (borders read 0 outside the image)
        assert np.array_equal(out[:, 2:-2, 2:-2], flat[:, 2:-2, 2:-2])
    rng = np.random.default_rng(11)
    base = np.where(rng.random((L, h, w)) < 0.5, 255, 0).astype(np.uint8)          # 1-sample soft-shadow noise: binary
    acc = vk.make_params(w, h, frame=5, denoise_frame=5)                            # accumulation frame: out = mix(old, center, 1/6)
    acc1 = vk.make_params(w, h, frame=5, denoise_frame=5, denoiser=1)
    hist = np.full((L, h, w), 128, np.uint8)
    raw = o.denoise(acc, g, base, hist).astype(np.float32)
    fil = o.denoise(acc1, g, base, hist).astype(np.float32)
    s = np.broadcast_to(surface, raw.shape)
    assert fil[s].std() < 0.5 * raw[s].std()                                       # the filtered centre carries far less noise into the history
    assert np.array_equal(fil[~s], raw[~s])                                        # background: the raw value, as before
    assert np.array_equal(o.denoise(p0, g, base, zero_hist), o.denoise(vk.make_params(w, h, frame=0, denoiser=0), g, base, zero_hist))
