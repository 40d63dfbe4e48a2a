"""GPU parity tests: the CUDA path, called through the C-ABI, against the CPU oracle on the same seeded inputs.

Bar (north_star): closest-hit instance/primitive ids identical on >= 99.9 % of rays, linear radiance MAE <= 1e-3 and
per-pixel max <= 1e-2.  Because both sides implement the same numerics contract the tests also report -- and for the
integer / byte planes require -- bit equality.
"""
import ctypes as C

import numpy as np
import pytest

import vulkanhybridraytracer_b200 as vk
from vulkanhybridraytracer_b200 import PLANE, STAGE
from vulkanhybridraytracer_b200 import types as T
from oracle.oracle_py import Oracle

from util import bits, random_rays, upload_gbuffer, golden_case, GOLDEN_DETAIL, GOLDEN_W, GOLDEN_H

pytestmark = pytest.mark.gpu

MAE_TOL, MAX_TOL, ID_TOL = 1e-3, 1e-2, 0.999      # north_star tolerances


def check_radiance(gpu, ref):
    d = np.abs(gpu[..., :3].astype(np.float64) - ref[..., :3].astype(np.float64))
    assert d.mean() <= MAE_TOL and d.max() <= MAX_TOL, (d.mean(), d.max())
    return d


def check_gbuffer(ctx, g):
    ids = ctx.readback(PLANE.HIT_IDS)
    assert (ids == g.hit_ids).all(-1).mean() >= ID_TOL
    for plane, ref in ((PLANE.POSITION, g.position), (PLANE.NORMAL, g.normal), (PLANE.ALBEDO, g.albedo), (PLANE.MOTION, g.motion),
                       (PLANE.MATERIAL, g.material), (PLANE.EMISSIVE, g.emissive)):
        got = ctx.readback(plane)
        same = (bits(got) == bits(ref)).all(-1).mean()
        assert same >= ID_TOL, (plane, same)
    return ids


def test_device_math_is_bit_identical_to_host(gpu_ctx):
    """The numerics contract: include/vkp_detmath.h evaluates to the same bits on sm_100a and on the host."""
    import subprocess, os, tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(tempfile.mkdtemp(), "libprobe.so")
    subprocess.check_call(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so,
                           os.path.join(root, "tests/native/detmath_probe.c"), "-lm"])
    P = C.CDLL(so); P.probe_eval.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    rng = np.random.default_rng(0)
    n = 1 << 20
    cases = {
        0: (rng.uniform(-7, 7, n), None), 1: (rng.uniform(-7, 7, n), None), 2: (rng.normal(size=n), rng.normal(size=n)),
        3: (rng.uniform(-1, 1, n), None), 4: (rng.uniform(0, 2, n), rng.choice([2.2, 1 / 2.2, 5.0, 2.0], n)),
        5: (rng.uniform(1e-6, 100, n), None), 6: (rng.uniform(-30, 30, n), None), 7: (rng.normal(scale=50, size=n), None),
        8: (rng.uniform(-0.2, 1.2, n), None),
    }
    for func, (x, y) in cases.items():
        x = x.astype(np.float32); y = np.zeros_like(x) if y is None else y.astype(np.float32)
        host = np.empty_like(x)
        P.probe_eval(func, x.ctypes.data, y.ctypes.data, n, host.ctypes.data)
        dev = gpu_ctx.eval_math(func, x, y)
        assert np.array_equal(dev.view(np.uint32), host.view(np.uint32)), func


@pytest.mark.parametrize("any_hit", [False, True])
def test_trace_rays_against_brute_force(gpu_ctx, small_scene, small_oracle, any_hit):
    rays = random_rays(small_scene, 20000, seed=3)
    ref = small_oracle.trace_rays(rays, any_hit=any_hit, brute_force=True)
    got = gpu_ctx.trace_rays(rays, any_hit=any_hit)
    if any_hit:
        assert np.array_equal(got["t"] >= 0, ref["t"] >= 0)
    else:
        same = (got["instance"] == ref["instance"]) & (got["primitive"] == ref["primitive"])
        assert same.mean() >= ID_TOL
        assert np.array_equal(got["instance"], ref["instance"]) and np.array_equal(got["primitive"], ref["primitive"])
        assert np.array_equal(got["t"].view(np.uint32), ref["t"].view(np.uint32))
        assert np.array_equal(got["u"].view(np.uint32)[same & (ref["instance"] >= 0)], ref["u"].view(np.uint32)[same & (ref["instance"] >= 0)])
    assert (ref["t"] >= 0).mean() > 0.3


def test_trace_edge_cases(gpu_ctx, small_scene, small_oracle):
    assert len(gpu_ctx.trace_rays(np.zeros(0, T.ray_dtype))) == 0                      # empty input
    rays = np.zeros(6, T.ray_dtype)
    rays["tmin"] = 0.001; rays["tmax"] = 1000
    rays["origin"][0] = [0, 5, 10]; rays["dir"][0] = [0, 0, 0]                            # zero direction: misses everything
    rays["origin"][1] = [0, 5, 10]; rays["dir"][1] = [0, -1, 0]                           # straight down onto the floor (axis-aligned)
    rays["origin"][2] = [0, 5, 10]; rays["dir"][2] = [0, -1, 0]; rays["tmax"][2] = 4.0    # tmax before the floor
    rays["origin"][3] = [0, 5, 10]; rays["dir"][3] = [0, -1, 0]; rays["tmin"][3] = 6.0    # tmin behind the floor
    rays["origin"][4] = [0, 1e-3, -5]; rays["dir"][4] = [1, 0, 0]                         # grazing just above the floor plane
    rays["origin"][5] = [0, 0, -5]; rays["dir"][5] = [1, 0, 0]                            # exactly in the floor plane (det == 0)
    ref = small_oracle.trace_rays(rays, brute_force=True)
    got = gpu_ctx.trace_rays(rays)
    assert np.array_equal(got["instance"], ref["instance"]) and np.array_equal(got["t"].view(np.uint32), ref["t"].view(np.uint32))
    assert got["instance"][0] == -1 and got["instance"][1] == 0 and abs(got["t"][1] - 5.0) < 1e-5 and got["instance"][2] == -1 and got["instance"][3] == -1


@pytest.mark.parametrize("w,h,wavefront,fuse", [(320, 184, 0, 1), (250, 130, 0, 0), (320, 184, 1, 1), (250, 130, 1, 0), (320, 184, 2, 0), (250, 130, 2, 1), (320, 184, 4, 0), (250, 130, 4, 1), (320, 184, 5, 0), (250, 130, 5, 1), (320, 184, 6, 0), (250, 130, 6, 1), (320, 184, 7, 0), (250, 130, 7, 1), (320, 184, 8, 0), (250, 130, 8, 1)])
def test_hybrid_frames_stage_by_stage(gpu_ctx, small_scene, small_oracle, w, h, wavefront, fuse):
    ctx, sc, o = gpu_ctx, small_scene, small_oracle
    cam = sc.camera(w, h)
    ctx.set_camera(cam)
    ctx.set_params(vk.make_params(w, h)); ctx.reset_history()
    hist = None
    for frame in range(3):
        p = vk.make_params(w, h, frame=frame, wavefront=wavefront, fuse_denoise=fuse)
        ctx.set_params(p)
        o.reset_counters()
        ref = o.hybrid_frame(cam, p, history=hist)
        hist = ref["denoised"]
        ctx.render()
        c = ctx.counters()
        check_gbuffer(ctx, ref["g"])
        for l in range(len(sc.lights)):
            assert np.array_equal(ctx.readback(PLANE.SHADOW0 + l), ref["shadow"][l])
            assert np.array_equal(ctx.readback(PLANE.DENOISED0 + l), ref["denoised"][l])
        final = ctx.readback(PLANE.FINAL_F32)
        check_radiance(final, ref["final"])
        assert np.array_equal(final.view(np.uint32), ref["final"].view(np.uint32))          # stronger than the bar: same bits
        assert np.array_equal(ctx.readback(PLANE.FINAL_SRGB8), o.post(p, ref["final"]))
        assert [c.rays_primary, c.rays_shadow, c.rays_reflect, c.rays_secondary_shadow] == list(o.counters.rays)
        # gbuffer, shadow, [denoise], resolve; 1: + one k_bounce per bounce (8), never fused; 2: + queue clear, k_chain_scan, k_chain;
        # 4: + queue clear, k_chain_scan, k_chain_first, k_chain_shadow0, k_chain_shade0, k_chain_rest; 5: as 2 + k_resolve_paths;
        # + the two k_build_order launches of the longest-first schedule (side stream)
        # frames after the reset frame (denoise_frame > 0): k_shadow's threads apply the accumulation filter themselves, no k_denoise launch
        fused_in_resolve = bool(fuse and wavefront != 1)
        accumulate_in_shadow = frame > 0 and not fused_in_resolve
        assert c.kernel_launches == {0: 4, 1: 4 + 8, 2: 7, 4: 10, 5: 8, 6: 8, 7: 12, 8: 9}[wavefront] - (1 if fused_in_resolve else 0) - (1 if accumulate_in_shadow else 0) + 2


@pytest.mark.parametrize("wavefront", [0, 1, 2, 3, 4, 5, 6, 7, 8])
def test_chain_schedules_with_the_advancing_rng(gpu_ctx, small_scene, small_oracle, wavefront):
    """faithful_rng = 0: every sampleDisk call advances the path's seed, so the secondary shadow ray of light l depends on which of the
    lights before it were sampled (hybridHit.rchit:99-131).  The staged chain (4) replays that per (path, light) thread; all five
    schedules must give the oracle's bits and ray counts."""
    ctx, sc, o = gpu_ctx, small_scene, small_oracle
    w, h = 280, 160
    cam = sc.camera(w, h)
    ctx.set_camera(cam)
    ctx.set_params(vk.make_params(w, h)); ctx.reset_history()
    hist = None
    for frame in range(2):
        p = vk.make_params(w, h, frame=frame, wavefront=wavefront, faithful_rng=0, shadow_samples=2)
        ctx.set_params(p)
        o.reset_counters()
        ref = o.hybrid_frame(cam, p, history=hist)
        hist = ref["denoised"]
        ctx.render()
        c = ctx.counters()
        final = ctx.readback(PLANE.FINAL_F32)
        assert np.array_equal(final.view(np.uint32), ref["final"].view(np.uint32))
        assert [c.rays_primary, c.rays_shadow, c.rays_reflect, c.rays_secondary_shadow] == list(o.counters.rays)


def test_atrous_denoiser_matches_the_oracle(gpu_ctx, small_scene, small_oracle):
    """vkp_params.denoiser = 1: three a-trous iterations (k_atrous) in front of blur.comp's temporal stage.  Reset frame,
    two accumulation frames and the stage in isolation against the oracle's restatement, bit for bit; 16 x 16 remainder handling
    and the advancing RNG with soft shadows (the noise the filter is for) included.  A tile-sharded context refuses the flag."""
    ctx, sc, o = gpu_ctx, small_scene, small_oracle
    w, h = 278, 154
    cam = sc.camera(w, h)
    ctx.set_camera(cam)
    ctx.set_params(vk.make_params(w, h)); ctx.reset_history()
    hist = None
    for frame in range(3):
        p = vk.make_params(w, h, frame=frame, wavefront=2, denoiser=1, faithful_rng=0, shadow_samples=2)
        ctx.set_params(p)
        ref = o.hybrid_frame(cam, p, history=hist)
        hist = ref["denoised"]
        ctx.render()
        c = ctx.counters()
        for l in range(len(sc.lights)):
            assert np.array_equal(ctx.readback(PLANE.SHADOW0 + l), ref["shadow"][l]), (frame, l)
            assert np.array_equal(ctx.readback(PLANE.DENOISED0 + l), ref["denoised"][l]), (frame, l)
        assert np.array_equal(ctx.readback(PLANE.FINAL_F32).view(np.uint32), ref["final"].view(np.uint32)), frame
        assert c.kernel_launches == 7 + 3 + 2          # chain mode 2 + the three k_atrous launches + the two k_build_order (the a-trous path keeps its filter pass)
    # the filter changes the image (it is not a no-op on this scene) ...
    p_off = vk.make_params(w, h, frame=0, wavefront=2, denoiser=0, faithful_rng=0, shadow_samples=2)
    p_on = vk.make_params(w, h, frame=0, wavefront=2, denoiser=1, faithful_rng=0, shadow_samples=2)
    ctx.set_params(p_off); ctx.reset_history(); ctx.render(); a = ctx.readback(PLANE.DENOISED0).copy()
    ctx.set_params(p_on); ctx.reset_history(); ctx.render(); b = ctx.readback(PLANE.DENOISED0).copy()
    assert not np.array_equal(a, b)
    # ... and a tile-sharded context says no
    ctx.set_tiling(0, 2, 32, 16)
    with pytest.raises(vk.VkpError):
        ctx.render()
    ctx.set_tiling(0, 1, 16, 8)
    ctx.set_params(p_off); ctx.reset_history(); ctx.render()


def test_stage_isolation(gpu_ctx, small_scene, small_oracle):
    """Feed the ORACLE's planes into each CUDA stage so that errors cannot compound."""
    ctx, sc, o = gpu_ctx, small_scene, small_oracle
    w, h = 200, 120
    cam = sc.camera(w, h); p = vk.make_params(w, h, frame=2, shadow_samples=3, faithful_rng=0)
    ctx.set_camera(cam); ctx.set_params(p)
    g = o.gbuffer(cam, p)
    rng = np.random.default_rng(5)
    history = rng.integers(0, 256, size=(len(sc.lights), h, w), dtype=np.uint8)
    sh = o.shadow(cam, p, g)
    den = o.denoise(p, g, sh, history)
    fin = o.resolve(cam, p, g, den)
    upload_gbuffer(ctx, g)
    ctx.run_stage(STAGE.SHADOW)
    for l in range(len(sc.lights)):
        assert np.array_equal(ctx.readback(PLANE.SHADOW0 + l), sh[l])
    for l in range(len(sc.lights)):
        ctx.upload_plane(PLANE.SHADOW0 + l, sh[l]); ctx.upload_plane(PLANE.DENOISED0 + l, history[l])
    ctx.run_stage(STAGE.DENOISE)
    for l in range(len(sc.lights)):
        assert np.array_equal(ctx.readback(PLANE.DENOISED0 + l), den[l])
    ctx.run_stage(STAGE.RESOLVE)
    got = ctx.readback(PLANE.FINAL_F32)
    check_radiance(got, fin)
    assert np.array_equal(got.view(np.uint32), fin.view(np.uint32))
    ctx.upload_plane(PLANE.FINAL_F32, fin)
    ctx.run_stage(STAGE.POST)
    assert np.array_equal(ctx.readback(PLANE.FINAL_SRGB8), o.post(p, fin))


@pytest.mark.parametrize("cover", [1, 0])
def test_denoise_first_frame_and_remainder(gpu_ctx, small_scene, small_oracle, cover):
    ctx, sc, o = gpu_ctx, small_scene, small_oracle
    w, h = 150, 70          # remainder columns / rows
    cam = sc.camera(w, h)
    ctx.set_camera(cam)
    for frame in (0, -1, 3):
        p = vk.make_params(w, h, frame=max(frame, 0), denoise_frame=frame, denoise_cover_remainder=cover)
        ctx.set_params(p)
        g = o.gbuffer(cam, p); sh = o.shadow(cam, p, g)
        history = np.random.default_rng(frame + 7).integers(0, 256, size=sh.shape, dtype=np.uint8)
        ref = o.denoise(p, g, sh, history)
        upload_gbuffer(ctx, g)
        for l in range(sh.shape[0]):
            ctx.upload_plane(PLANE.SHADOW0 + l, sh[l]); ctx.upload_plane(PLANE.DENOISED0 + l, history[l])
        ctx.run_stage(STAGE.DENOISE)
        for l in range(sh.shape[0]):
            assert np.array_equal(ctx.readback(PLANE.DENOISED0 + l), ref[l]), (frame, l)


def test_soft_shadow_sampling_modes(gpu_ctx, small_scene, small_oracle):
    ctx, sc, o = gpu_ctx, small_scene, small_oracle
    w, h = 160, 96
    cam = sc.camera(w, h); ctx.set_camera(cam)
    for samples, faithful in ((1, 1), (8, 1), (8, 0), (64, 0)):
        p = vk.make_params(w, h, shadow_samples=samples, faithful_rng=faithful, frame=5)
        ctx.set_params(p)
        g = o.gbuffer(cam, p); ref = o.shadow(cam, p, g)
        upload_gbuffer(ctx, g); ctx.run_stage(STAGE.SHADOW)
        for l in range(ref.shape[0]):
            assert np.array_equal(ctx.readback(PLANE.SHADOW0 + l), ref[l]), (samples, faithful, l)


def test_raytraced_mode(gpu_ctx, small_scene, small_oracle):
    ctx, sc, o = gpu_ctx, small_scene, small_oracle
    w, h = 160, 96
    cam = sc.camera(w, h); ctx.set_camera(cam)
    ctx.set_params(vk.make_params(w, h, mode=vk.MODE_RAYTRACING)); ctx.reset_history()
    acc = None
    for frame in range(3):
        p = vk.make_params(w, h, mode=vk.MODE_RAYTRACING, spp=3, shadow_samples=2, faithful_rng=0, frame=frame)
        ctx.set_params(p)
        o.reset_counters()
        acc = o.raytrace(cam, p, accum=acc)
        ctx.render()
        got = ctx.readback(PLANE.FINAL_F32)
        check_radiance(got, acc)
        assert np.array_equal(got.view(np.uint32), acc.view(np.uint32))
        c = ctx.counters()
        assert [c.rays_reflect, c.rays_secondary_shadow] == list(o.counters.rays)[2:]
    # many shadow samples: the one-warp-per-pixel kernel (lanes split the samples), advancing and faithful RNG, 40 > 32 samples
    for faithful in (0, 1):
        p = vk.make_params(w, h, mode=vk.MODE_RAYTRACING, spp=2, shadow_samples=40, faithful_rng=faithful, frame=0)
        ctx.set_params(p); ctx.reset_history(); ctx.render()
        o.reset_counters()
        ref = o.raytrace(cam, p)
        assert np.array_equal(ctx.readback(PLANE.FINAL_F32).view(np.uint32), ref.view(np.uint32)), faithful
        c = ctx.counters()
        assert [c.rays_reflect, c.rays_secondary_shadow] == list(o.counters.rays)[2:]
    # faithful variant (seed by value, closesthit.rchit:107)
    p = vk.make_params(w, h, mode=vk.MODE_RAYTRACING, spp=1, shadow_samples=4, faithful_rng=1, frame=0)
    ctx.set_params(p); ctx.reset_history(); ctx.render()
    assert np.array_equal(ctx.readback(PLANE.FINAL_F32).view(np.uint32), o.raytrace(cam, p).view(np.uint32))


def test_deferred_mode_and_debug_targets(gpu_ctx, small_scene, small_oracle):
    ctx, sc, o = gpu_ctx, small_scene, small_oracle
    w, h = 160, 96
    cam = sc.camera(w, h); ctx.set_camera(cam)
    for target in range(7):
        p = vk.make_params(w, h, mode=vk.MODE_DEFERRED, debug_target=target)
        ctx.set_params(p); ctx.render()
        g = o.gbuffer(cam, p); ref = o.deferred(cam, p, g)
        got = ctx.readback(PLANE.FINAL_F32)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), target


def test_moving_camera_and_light_updates(gpu_ctx, small_scene, small_oracle):
    """Per-frame uniform path of VulkanEngine::update: new camera (motion vectors != 0), new lights, frame reset."""
    import vulkanhybridraytracer_b200 as vk2
    ctx, o = gpu_ctx, small_oracle
    sc = vk2.HostScene(0, detail=(4, 3, 4, 6))
    w, h = 192, 108
    cam0 = sc.camera(w, h)
    sc.move_camera(3, 8.0); sc.rotate_camera(40.0, -25.0)       # strafe right, look around
    cam1 = sc.camera(w, h, prev=cam0)
    lights = sc.lights.copy(); lights["position"][0][:3] = [6, 9, -1]; lights["color"][1] = [1, 0.2, 0.2, 400]
    ctx.set_lights(lights); o.set_lights(lights)
    try:
        p = vk.make_params(w, h, frame=0, denoise_frame=0)
        ctx.set_camera(cam1); ctx.set_params(p); ctx.reset_history(); ctx.render()
        ref = o.hybrid_frame(cam1, p)
        check_gbuffer(ctx, ref["g"])
        assert (ref["g"].motion != np.float16(0.5)).any()
        for l in range(3):
            assert np.array_equal(ctx.readback(PLANE.DENOISED0 + l), ref["denoised"][l])
        assert np.array_equal(ctx.readback(PLANE.FINAL_F32).view(np.uint32), ref["final"].view(np.uint32))
    finally:
        ctx.set_lights(small_scene.lights); o.set_lights(small_scene.lights)
        sc.close()


def test_golden_fixtures_on_gpu():
    """The committed fixtures (tests/golden, generated from the oracle) reproduced by the CUDA path."""
    import os
    sc = vk.HostScene(0, detail=GOLDEN_DETAIL)
    ctx = vk.Context(0); ctx.upload_scene(sc)
    W, H = GOLDEN_W, GOLDEN_H
    cam = sc.camera(W, H); ctx.set_camera(cam)
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    ctx.set_params(vk.make_params(W, H, frame=0)); ctx.reset_history(); ctx.render()
    f0 = np.load(os.path.join(root, "hybrid_f0.npz"))
    assert np.array_equal(ctx.readback(PLANE.HIT_IDS), f0["hit_ids"]) and np.array_equal(ctx.readback(PLANE.POSITION).view(np.uint16), f0["position"])
    assert np.array_equal(ctx.readback(PLANE.FINAL_F32).view(np.uint32), f0["final_bits"])
    ctx.set_params(vk.make_params(W, H, frame=1)); ctx.render()
    f1 = np.load(os.path.join(root, "hybrid_f1.npz"))
    for l in range(3):
        assert np.array_equal(ctx.readback(PLANE.DENOISED0 + l), f1["denoised"][l])
    assert np.array_equal(ctx.readback(PLANE.FINAL_F32).view(np.uint32), f1["final_bits"])
    p = vk.make_params(W, H, mode=vk.MODE_RAYTRACING, spp=2, shadow_samples=2, faithful_rng=0)
    ctx.set_params(p); ctx.reset_history(); ctx.render()
    assert np.array_equal(ctx.readback(PLANE.FINAL_F32).view(np.uint32), np.load(os.path.join(root, "raytrace.npz"))["final_bits"])
    ctx.set_params(vk.make_params(W, H, mode=vk.MODE_DEFERRED)); ctx.render()
    assert np.array_equal(ctx.readback(PLANE.FINAL_F32).view(np.uint32), np.load(os.path.join(root, "deferred.npz"))["final_bits"])
    ctx.close(); sc.close()
