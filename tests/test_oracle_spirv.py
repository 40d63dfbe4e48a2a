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
    max and pow(x, 5): relative 1e-6."""
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
            assert abs(d - wd) <= 1e-6 * max(abs(wd), 1e-30) + 1e-30, (k, d, wd)
        else:
            assert not np.isfinite(d) or d > 1e30, (k, d, wd)                 # roughness 0 on the lobe axis: a2 / 0
        assert abs(g - wg) <= 1e-6 * max(abs(wg), 1e-6), (k, g, wg)
        assert np.allclose(np.array(fr[:], np.float32), wf, rtol=2e-6, atol=1e-7), (k, fr[:], wf)
