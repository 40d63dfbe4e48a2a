"""CPU tests: the numerics contract header (include/vkp_detmath.h) and the C-ABI surface (include/vkp.h)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import vulkanhybridraytracer_b200 as vk
from vulkanhybridraytracer_b200 import types as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    out = tmp_path_factory.mktemp("probe") / "libprobe.so"
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", str(out), os.path.join(ROOT, "tests/native/detmath_probe.c"), "-lm"])
    L = C.CDLL(str(out))
    L.probe_eval.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]

    def ev(func, x, y=None):
        x = np.ascontiguousarray(x, np.float32)
        y = np.zeros_like(x) if y is None else np.ascontiguousarray(y, np.float32)
        o = np.empty_like(x)
        L.probe_eval(func, x.ctypes.data, y.ctypes.data, x.size, o.ctypes.data)
        return o
    return ev


def ulp_err(got, ref64):
    ref32 = ref64.astype(np.float32)
    ulp = np.spacing(np.abs(ref32)).astype(np.float64)
    ulp = np.maximum(ulp, np.finfo(np.float32).tiny)
    return np.abs(got.astype(np.float64) - ref64) / ulp


def test_sincos_accuracy(probe):
    x = np.linspace(-7.0, 7.0, 200001).astype(np.float32)
    assert ulp_err(probe(0, x), np.sin(x.astype(np.float64))).max() < 2.5 or np.abs(probe(0, x) - np.sin(x.astype(np.float64))).max() < 2e-7
    assert np.abs(probe(0, x) - np.sin(x.astype(np.float64))).max() < 2e-7
    assert np.abs(probe(1, x) - np.cos(x.astype(np.float64))).max() < 2e-7
    # the hot path's range: angle = rnd * 2 * PI in [0, 2pi)
    x = (np.arange(0, 1 << 24, 257, dtype=np.float32) / np.float32(16777216.0)) * np.float32(2.0) * np.float32(3.14159265359)
    assert np.abs(probe(0, x) - np.sin(x.astype(np.float64))).max() < 2e-7
    assert np.abs(probe(1, x) - np.cos(x.astype(np.float64))).max() < 2e-7


def test_atan2_asin_accuracy(probe):
    rng = np.random.default_rng(1)
    d = rng.normal(size=(200000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True); d = d.astype(np.float32)
    assert np.abs(probe(2, d[:, 0], d[:, 2]) - np.arctan2(d[:, 0].astype(np.float64), d[:, 2].astype(np.float64))).max() < 1e-6
    assert np.abs(probe(3, d[:, 1]) - np.arcsin(d[:, 1].astype(np.float64))).max() < 1e-6
    # quadrants and axes
    xs = np.array([0, 0, 1, -1, 1, -1, 0.0], np.float32); ys = np.array([1, -1, 0, 0, 1, -1, 0.0], np.float32)
    assert np.abs(probe(2, xs, ys) - np.arctan2(xs.astype(np.float64), ys.astype(np.float64))).max() < 1e-6
    assert probe(3, np.array([1.0, -1.0, 2.0], np.float32)).tolist() == pytest.approx([np.pi / 2, -np.pi / 2, np.pi / 2], abs=1e-6)


def test_pow_accuracy(probe):
    x = np.linspace(0.0, 1.0, 100001).astype(np.float32)
    for e in (2.2, 1.0 / 2.2, 5.0, 2.0):
        got = probe(4, x, np.full_like(x, e))
        ref = np.power(x.astype(np.float64), np.float64(np.float32(e)))
        assert np.abs(got - ref).max() < 4e-6, e
        assert ((np.abs(got - ref) / np.maximum(ref, 1e-30))[x > 0.003]).max() < 4e-6, e
    # unorm8 albedo decode, the exact inputs of hybridRaygen.rgen:41
    x = (np.arange(256) / 255.0).astype(np.float32)
    got = probe(4, x, np.full_like(x, 2.2))
    assert got[0] == 0.0 and got[255] == 1.0
    assert np.abs(got - np.power(x.astype(np.float64), np.float64(np.float32(2.2)))).max() < 1e-6
    # pow(x, 0) = 1, pow(0, y) = 0, ior-like bases > 1
    assert probe(4, np.array([0.0, 3.0], np.float32), np.array([0.0, 0.0], np.float32)).tolist() == [1.0, 1.0]
    b = np.linspace(0.5, 3.0, 1001).astype(np.float32)
    assert (np.abs(probe(4, b, np.full_like(b, 2.0)) - b.astype(np.float64) ** 2) / b.astype(np.float64) ** 2).max() < 2e-6


def test_half_and_unorm_probe_matches_numpy(probe):
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.normal(scale=s, size=50000) for s in (1e-5, 1.0, 30.0, 1000.0)]).astype(np.float32)
    assert np.array_equal(probe(7, x), x.astype(np.float16).astype(np.float32))
    y = rng.uniform(-0.2, 1.2, size=100000).astype(np.float32)
    ref = np.rint(np.clip(y, 0, 1).astype(np.float32) * np.float32(255.0))
    assert np.array_equal(probe(8, y), ref.astype(np.float32))


# ---- C-ABI surface -------------------------------------------------------------------------------------
def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vkp.h")).read()
    return re.findall(r"VKP_API\s+[\w\s\*]+?\b(vkp_\w+)\s*\(", text)


def test_library_exports_every_declared_symbol():
    L = vk.load_vkp()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/vkp.h but not exported by libvkp.so"
    # and the binding declares a prototype for each of them
    assert set(names) <= set(L._signatures.keys())
    assert L.vkp_abi_version() == 1


def test_struct_sizes_match_the_reference_layouts():
    for cls, size in T.SIZES.items():
        assert C.sizeof(cls) == size, cls
    assert C.sizeof(T.Counters) == 32 + 32 + 8 + 8 + 24 + 8


def test_create_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = vk.load_vkp()
    h = C.c_void_p()
    r = L.vkp_create(C.byref(h), 0)
    assert r == -3 and not h.value
    assert b"no CPU fallback" in L.vkp_last_error(None)
    with pytest.raises(vk.VkpError):
        vk.Context(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vulkanhybridraytracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                # comments may name the oracle; nothing may include, import, link or load it
                assert "liboracle" not in text and "oracle_py" not in text and "oracle/" not in text and "orc_" not in text, os.path.join(dirpath, f)
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), os.path.join(dirpath, f)


def test_group_entry_points_validate_their_arguments_without_a_gpu():
    """vkp_group_* (include/vkp.h "multi-GPU group"): argument errors are reported before any CUDA call, so this runs on the CPU box."""
    import ctypes as C
    L = vk.load_vkp()
    assert L.vkp_group_size(None) == 0
    assert L.vkp_group_member(None, 0) is None
    g = C.c_void_p()
    devs = (C.c_int * 1)(0)
    assert L.vkp_group_create(C.byref(g), devs, 0, 32, 16) == -1 and not g.value            # VKP_ERR_INVALID: no devices
    assert L.vkp_group_create(C.byref(g), devs, 65, 32, 16) == -1 and not g.value           # more than 64 ranks
    assert L.vkp_group_create(None, devs, 1, 32, 16) == -1
    assert L.vkp_group_render(None) == -1 and L.vkp_group_sync(None) == -1
    L.vkp_group_destroy(None)                                                               # a no-op, like free(NULL)
