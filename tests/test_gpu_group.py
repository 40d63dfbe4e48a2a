"""vkp_group: several GPUs behind one handle in ONE process (include/vkp.h "multi-GPU group"), the path the C++ host uses
(Renderer / vkpinut --gpus N) where the reference has its single device (vk_engine.cpp:296-302, Renderer::rasterize_hybrid
renderer.cpp:530-600).  The composited image must be bit-identical to one GPU.  On a one-GPU box the ranks share device 0 (same
code path: tiling, peer redirect of the sRGB8 stores, event ordering, reset-frame exchange); with >= 2 GPUs the real thing runs too."""
import os
import subprocess
import sys

import numpy as np
import pytest

import vulkanhybridraytracer_b200 as vk
from vulkanhybridraytracer_b200 import PLANE

pytestmark = pytest.mark.gpu


def _device_count():
    import torch
    return torch.cuda.device_count()


def _cameras(w, h):
    """Two views of a private copy of the small scene: frames 0-1 from the first, frame 2 after a jump (previous matrices = the first
    view, so motion vectors are non-zero and the history is reset mid-sequence)."""
    from conftest import SMALL_DETAIL
    sc = vk.HostScene(0, detail=SMALL_DETAIL)
    cam0 = sc.camera(w, h)
    sc.move_camera(3, 6.0); sc.rotate_camera(25.0, -10.0)
    cam1 = sc.camera(w, h, prev=cam0)
    sc.close()
    return cam0, cam1


def _frames(target, w, h, mode, spp=1, samples=1):
    """4 frames on `target` (Context or Group): reset frame, static frame, camera jump (reset frame with motion), static frame."""
    out = []
    cam0, cam1 = _cameras(w, h)
    target.set_params(vk.make_params(w, h, mode=mode)); target.reset_history()
    for cam, dframe in ((cam0, 0), (cam0, 1), (cam1, 0), (cam1, 1)):
        target.set_camera(cam)
        target.set_params(vk.make_params(w, h, mode=mode, spp=spp, shadow_samples=samples, frame=dframe, denoise_frame=dframe, wavefront=2))
        target.render()
        out.append((target.readback(PLANE.FINAL_SRGB8).copy(), target.readback(PLANE.FINAL_F32).copy()))
    return out


@pytest.mark.parametrize("ranks", [2, 3])
def test_group_on_one_device_is_bit_identical(small_scene, ranks):
    sc = small_scene
    w, h = 272, 150                                  # not a multiple of the 32 x 16 tiles
    solo = vk.Context(0); solo.upload_scene(sc)
    grp = vk.Group([0] * ranks); grp.upload_scene(sc)
    assert grp.size() == ranks
    for mode, kw in ((vk.MODE_HYBRID, {}), (vk.MODE_RAYTRACING, dict(spp=2, samples=2)), (vk.MODE_DEFERRED, {})):
        a = _frames(solo, w, h, mode, **kw)
        b = _frames(grp, w, h, mode, **kw)
        for f, ((s8a, f32a), (s8b, f32b)) in enumerate(zip(a, b)):
            assert np.array_equal(s8a, s8b), (mode, f)
            assert np.array_equal(f32a.view(np.uint32), f32b.view(np.uint32)), (mode, f)
    ca, cb = solo.counters(), grp.counters()
    assert [ca.rays_primary, ca.rays_shadow, ca.rays_reflect, ca.rays_secondary_shadow] == [cb.rays_primary, cb.rays_shadow, cb.rays_reflect, cb.rays_secondary_shadow]
    grp.close(); solo.close()


def test_group_on_several_gpus_is_bit_identical(small_scene):
    n = _device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    sc = small_scene
    w, h = 320, 184
    solo = vk.Context(0); solo.upload_scene(sc)
    grp = vk.Group(list(range(min(n, 8)))); grp.upload_scene(sc)
    for mode, kw in ((vk.MODE_HYBRID, {}), (vk.MODE_RAYTRACING, dict(spp=2, samples=2))):
        a = _frames(solo, w, h, mode, **kw)
        b = _frames(grp, w, h, mode, **kw)
        for f, ((s8a, f32a), (s8b, f32b)) in enumerate(zip(a, b)):
            assert np.array_equal(s8a, s8b), (mode, f)
            assert np.array_equal(f32a.view(np.uint32), f32b.view(np.uint32)), (mode, f)
    grp.close(); solo.close()


def test_vkpinut_gpus_writes_the_same_image(tmp_path):
    """bin/vkpinut --gpus 2 (the C++ host: Renderer over a vkp_group) against --gpus 1: same PNG bytes, Hybrid with a moving camera
    and Raytraced.  With one GPU in the box both ranks share it (VKP_GROUP_SAME_DEVICE)."""
    exe = os.path.join(os.path.dirname(vk.lib_dir()), "bin", "vkpinut")
    assert os.path.exists(exe), "make host"
    env = dict(os.environ)
    if _device_count() < 2:
        env["VKP_GROUP_SAME_DEVICE"] = "1"
    for mode, extra in (("hybrid", ["--walk", "forward:40", "--frames", "3"]), ("raytracing", ["--spp", "2", "--samples", "2", "--frames", "2"])):
        imgs = []
        for gpus in (1, 2):
            png = str(tmp_path / f"{mode}_{gpus}.png")
            r = subprocess.run([exe, "--width", "256", "--height", "144", "--detail", "4,3,4,9", "--mode", mode, "--gpus", str(gpus), "--png", png] + extra,
                               env=env, capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stdout + r.stderr
            imgs.append(open(png, "rb").read())
        assert imgs[0] == imgs[1], mode
