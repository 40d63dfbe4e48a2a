"""Known-answer and unit tests that pin the oracle's building blocks (CPU only).

The reference ships no vectors; SURVEY.md App. D derived TEA/LCG answers from random.glsl:3-33, the rest is
checked against float64 restatements of the GLSL formulas written here (helpers.glsl:33-93, miss.rmiss:16).
"""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import oracle_py
from vulkanhybridraytracer_b200 import types as T


@pytest.fixture(scope="module")
def L():
    return oracle_py.load()


# (val0, val1) -> tea, first rnd bits, second rnd bits     
TEA_KAT = [
    ((0, 0), 0x741C187D, 13828024, 81847),
    ((1, 0), 0x8DA6B311, 15192380, 7240555),
    ((0, 1), 0x70D3AEF1, 7068572, 10830411),
    ((1919, 0), 0x97CFE7DD, 16041880, 3909143),
    ((2073599, 0), 0xCDA8C132, 14462697, 1420084),
    ((1037760, 7), 0xC908DEC6, 16459629, 9069032),
]


@pytest.mark.parametrize("args,tea,b1,b2", TEA_KAT)
def test_tea_lcg_known_answers(L, args, tea, b1, b2):
    assert L.orc_tea(*args) == tea
    s = C.c_uint32(tea)
    assert L.orc_lcg(C.byref(s)) == b1
    assert L.orc_lcg(C.byref(s)) == b2
    s = C.c_uint32(tea)
    assert L.orc_rnd(C.byref(s)) == np.float32(b1) / np.float32(16777216.0)


def test_tea_against_python_restatement(L):
    def tea(v0, v1):
        s0 = 0
        for _ in range(16):
            s0 = (s0 + 0x9E3779B9) & 0xFFFFFFFF
            v0 = (v0 + ((((v1 << 4) & 0xFFFFFFFF) + 0xA341316C) ^ (v1 + s0) ^ ((v1 >> 5) + 0xC8013EA4))) & 0xFFFFFFFF
            v1 = (v1 + ((((v0 << 4) & 0xFFFFFFFF) + 0xAD90777D) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7E95761E))) & 0xFFFFFFFF
        return v0
    rng = np.random.default_rng(1234)
    for a, b in rng.integers(0, 2**32, size=(200, 2), dtype=np.uint64):
        assert L.orc_tea(int(a), int(b)) == tea(int(a), int(b))


def f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def test_sample_disk_geometry(L):
    """helpers.glsl:82-93: the sample lies on the unit-distance disk of radius light.radius around L, |dir| = 1,
    and the seed is taken by value (same seed -> same direction)."""
    rng = np.random.default_rng(7)
    light = T.Light()
    for _ in range(200):
        light.radius = float(rng.uniform(0.01, 0.5))
        Lv = rng.normal(size=3); Lv /= np.linalg.norm(Lv)
        pos = rng.uniform(-10, 10, size=3)
        seed = int(rng.integers(0, 2**32))
        out = (C.c_float * 3)(); out2 = (C.c_float * 3)()
        L.orc_sample_disk(C.byref(light), C.byref(f3(pos)), C.byref(f3(Lv)), seed, C.byref(out))
        L.orc_sample_disk(C.byref(light), C.byref(f3(pos)), C.byref(f3(Lv)), seed, C.byref(out2))
        d = np.array(out[:]); assert np.array_equal(d, np.array(out2[:]))
        assert abs(np.linalg.norm(d) - 1.0) < 1e-5
        # angle between dir and L is at most atan(radius)
        cosang = float(np.dot(d, Lv.astype(np.float32)))
        assert cosang >= math.cos(math.atan(light.radius)) - 1e-5
    # the documented special case L == (0,1,0): tangent = (0,0,1) (helpers.glsl:88-89), finite result
    out = (C.c_float * 3)()
    light.radius = 0.2
    L.orc_sample_disk(C.byref(light), C.byref(f3([0, 0, 0])), C.byref(f3([0, 1, 0])), 1, C.byref(out))
    assert np.isfinite(np.array(out[:])).all()


def test_sample_disk_matches_float64_formula(L):
    rng = np.random.default_rng(11)
    light = T.Light()
    for _ in range(100):
        light.radius = float(rng.uniform(0.05, 0.3))
        Lv = rng.normal(size=3); Lv /= np.linalg.norm(Lv); Lv = Lv.astype(np.float32)
        pos = rng.uniform(-5, 5, size=3).astype(np.float32)
        seed = int(rng.integers(0, 2**32))
        s = C.c_uint32(seed)
        r1 = L.orc_rnd(C.byref(s)); r2 = L.orc_rnd(C.byref(s))
        radius = light.radius * math.sqrt(r1); angle = r2 * 2.0 * 3.14159265359
        t = np.cross(Lv.astype(np.float64), [0, 1, 0]); t /= np.linalg.norm(t)
        b = np.cross(t, Lv.astype(np.float64)); b /= np.linalg.norm(b)
        target = Lv + radius * math.cos(angle) * t + radius * math.sin(angle) * b
        ref = target / np.linalg.norm(target)
        out = (C.c_float * 3)()
        L.orc_sample_disk(C.byref(light), C.byref(f3(pos)), C.byref(f3(Lv)), seed, C.byref(out))
        assert np.allclose(np.array(out[:]), ref, atol=2e-5)


def test_brdf_terms_match_float64(L):
    PI = 3.14159265359
    rng = np.random.default_rng(3)
    for _ in range(300):
        N = rng.normal(size=3); N /= np.linalg.norm(N)
        V = rng.normal(size=3); V /= np.linalg.norm(V)
        Lv = rng.normal(size=3); Lv /= np.linalg.norm(Lv)
        Hh = V + Lv; Hh /= np.linalg.norm(Hh)
        r = float(rng.uniform(0.05, 1.0))
        a2 = (r * r) ** 2
        ndh = max(float(N @ Hh), 0.0)
        denom = ndh * ndh * (a2 - 1.0) + 1.0
        ggx = a2 / (PI * denom * denom)
        got = L.orc_distribution_ggx(C.byref(f3(N)), C.byref(f3(Hh)), r)
        assert math.isclose(got, ggx, rel_tol=2e-4, abs_tol=1e-7)
        k = (r + 1.0) ** 2 / 8.0
        g1 = lambda x: x / (x * (1.0 - k) + k)
        smith = g1(max(float(N @ V), 0.0)) * g1(max(float(N @ Lv), 0.0))
        got = L.orc_geometry_smith(C.byref(f3(N)), C.byref(f3(V)), C.byref(f3(Lv)), r)
        assert math.isclose(got, smith, rel_tol=2e-4, abs_tol=1e-7)
        c = float(rng.uniform(0, 1)); F0 = rng.uniform(0, 1, size=3)
        out = (C.c_float * 3)()
        L.orc_fresnel_schlick(c, C.byref(f3(F0)), C.byref(out))
        assert np.allclose(np.array(out[:]), F0 + (1 - F0) * (1 - c) ** 5, rtol=2e-4, atol=1e-6)


def test_env_uv_matches_float64(L):
    rng = np.random.default_rng(5)
    for _ in range(500):
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        for use_miss, pi in ((0, 3.14159265359), (1, 3.141592)):
            out = (C.c_float * 2)()
            L.orc_env_uv(C.byref(f3(d)), use_miss, C.byref(out))
            u = 0.5 + math.atan2(d[0], d[2]) / (2 * pi); v = 0.5 - math.asin(d[1]) / pi
            assert abs(out[0] - u) < 2e-6 and abs(out[1] - v) < 2e-6
    # +z looks at u = 0.5, straight up is v = 0 (miss.rmiss:16)
    out = (C.c_float * 2)()
    L.orc_env_uv(C.byref(f3([0, 0, 1])), 0, C.byref(out)); assert abs(out[0] - 0.5) < 1e-7 and abs(out[1] - 0.5) < 1e-7
    L.orc_env_uv(C.byref(f3([0, 1, 0])), 0, C.byref(out)); assert abs(out[1]) < 1e-6


def test_half_conversion_is_numpy_float16_rne(L):
    rng = np.random.default_rng(9)
    xs = np.concatenate([rng.normal(scale=s, size=4000) for s in (1e-6, 1e-3, 1.0, 50.0, 3e4)]).astype(np.float32)
    xs = np.concatenate([xs, np.array([0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e9, -1e9, 5.96e-8, 2.98e-8, 2.99e-8, 6.1e-5], np.float32)])
    with np.errstate(over="ignore"):
        ref = xs.astype(np.float16)
    for x, r in zip(xs, ref):
        assert L.orc_to_half(float(x)) == int(r.view(np.uint16)), x
    # all 65536 half patterns decode like numpy
    allh = np.arange(65536, dtype=np.uint16)
    dec = allh.view(np.float16).astype(np.float32)
    for h in range(0, 65536, 7):
        got = L.orc_from_half(h)
        assert (math.isnan(got) and math.isnan(dec[h])) or got == dec[h]


def test_unorm8_round_half_even(L):
    assert L.orc_to_unorm8(-1.0) == 0 and L.orc_to_unorm8(float("nan")) == 0 and L.orc_to_unorm8(2.0) == 255
    assert L.orc_to_unorm8(1.0) == 255 and L.orc_to_unorm8(0.0) == 0
    for k in range(256):
        assert L.orc_to_unorm8(k / 255.0) == k
    assert L.orc_to_unorm8(0.5 / 255.0) == 0      # tie -> even
    assert L.orc_to_unorm8(1.5 / 255.0) == 2
    assert L.orc_to_unorm8(0.4) == 102
