/*
 * vkp_detmath.h -- deterministic single-precision elementary functions.
 *
 * The reference's GLSL leaves sin/cos/atan/asin/pow to the Vulkan driver
 * (helpers.glsl:86 cos/sin, miss.rmiss:16 atan/asin, hybridRaygen.rgen:41 pow,
 * helpers.glsl:71 pow); their rounding is implementation-defined.  Shadow
 * visibility is binary, so two "correct" libms that differ by one ulp flip
 * pixels on penumbra edges.  This header pins ONE implementation, built only
 * from IEEE-754 correctly-rounded +,-,*,/,sqrt, fused multiply-add and integer
 * bit operations in a fixed order, so that host (gcc, -ffp-contract=off) and
 * device (nvcc, -fmad=false) produce bit-identical results.
 *
 * Contract v2 (round 2): fused multiply-add is part of the contract, but only
 * where it is WRITTEN -- vkp_fmaf() is fmaf() on the host (correctly rounded by
 * C99, one vfmadd with -mfma) and __fmaf_rn() on the device; neither compiler
 * may contract anything on its own.  The polynomial kernels below are Horner
 * chains of vkp_fmaf, and the vector primitives at the end of the file (dot,
 * cross component, matrix row, mix) are what BOTH the CUDA kernels
 * (vkp_device.cuh) and the CPU oracle (oracle/orc_math.h) are built from: the
 * traversal and shading kernels issue half as many FP instructions as the
 * mul/add form of round 1 and stay bit-identical to the scalar evaluation.  It is the numerics contract of
 * the C-ABI, not part of the reference's algorithm; the CPU oracle may be
 * compiled against it (default) or against glibc (-DORC_USE_LIBM) and the
 * tests run both.
 *
 * Polynomial kernels follow the classic Cephes single-precision forms
 * (Moshier), accurate to ~1-2 ulp on the ranges the hot path uses.
 */
#ifndef VKP_DETMATH_H
#define VKP_DETMATH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define VKP_HD __host__ __device__ __forceinline__
#else
#define VKP_HD static inline
#endif

#ifdef __cplusplus
#include <cmath>
#include <cstring>
#else
#include <math.h>
#include <string.h>
#endif

VKP_HD uint32_t vkp_f2u(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
VKP_HD float vkp_u2f(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

/* fused multiply-add, one rounding: fmaf (C99) on the host, FFMA on the device */
VKP_HD float vkp_fmaf(float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}

/* IEEE sqrt: correctly rounded on both sides (nvcc default -prec-sqrt=true). */
VKP_HD float vkp_sqrtf(float x) { return sqrtf(x); }

#define VKP_PI_F      3.14159265359f   /* helpers.glsl:5  */
#define VKP_PI_MISS_F 3.141592f        /* miss.rmiss:11   */

/* round-to-nearest-even to integer for |x| < 2^22, without libm */
VKP_HD float vkp_rintf_small(float x)
{
    const float magic = 12582912.0f; /* 1.5 * 2^23 */
    float t = x + magic;
    return t - magic;
}

/* sin and cos of x (radians), |x| up to ~1e4 with graceful accuracy loss. */
VKP_HD void vkp_sincosf(float x, float* s, float* c)
{
    /* k = nearest integer to x * 2/pi ; r = x - k*pi/2 via 3-term Cody-Waite */
    float kf = vkp_rintf_small(x * 0.636619772367581343f);
    int   k  = (int)kf;
    float r  = vkp_fmaf(-kf, 1.5703125f, x);     /* pi/2 high bits (exact product for small k) */
    r = vkp_fmaf(-kf, 4.837512969970703125e-4f, r);
    r = vkp_fmaf(-kf, 7.54978995489188e-8f, r);
    float z = r * r;
    /* Cephes sinf / cosf kernels on [-pi/4, pi/4] */
    float ps = vkp_fmaf(vkp_fmaf(vkp_fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f) * z, r, r);
    float pc = vkp_fmaf(vkp_fmaf(vkp_fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f) * z, z,
                        vkp_fmaf(-0.5f, z, 1.0f));
    float sv, cv;
    switch (k & 3) {
    case 0:  sv =  ps; cv =  pc; break;
    case 1:  sv =  pc; cv = -ps; break;
    case 2:  sv = -ps; cv = -pc; break;
    default: sv = -pc; cv =  ps; break;
    }
    *s = sv; *c = cv;
}
VKP_HD float vkp_sinf(float x) { float s, c; vkp_sincosf(x, &s, &c); return s; }
VKP_HD float vkp_cosf(float x) { float s, c; vkp_sincosf(x, &s, &c); return c; }

/* atan(x), Cephes atanf */
VKP_HD float vkp_atanf(float xx)
{
    float x = xx < 0.0f ? -xx : xx;
    float y;
    if (x > 2.414213562373095f) {          /* tan(3pi/8) */
        y = 0.78539816339744830962f * 2.0f;
        x = -(1.0f / x);
    } else if (x > 0.4142135623730950f) {  /* tan(pi/8) */
        y = 0.78539816339744830962f;
        x = (x - 1.0f) / (x + 1.0f);
    } else {
        y = 0.0f;
    }
    float z = x * x;
    y = y + vkp_fmaf(vkp_fmaf(vkp_fmaf(vkp_fmaf(8.05374449538e-2f, z, -1.38776856032e-1f), z, 1.99777106478e-1f), z,
                              -3.33329491539e-1f) * z, x, x);
    return xx < 0.0f ? -y : y;
}

/* GLSL atan(y, x) == atan2(y, x) */
VKP_HD float vkp_atan2f(float y, float x)
{
    const float PI_F = 3.14159265358979323846f;
    if (x == 0.0f) {
        if (y == 0.0f) return 0.0f;
        return y > 0.0f ? 0.5f * PI_F : -0.5f * PI_F;
    }
    float a = vkp_atanf(y / x);
    if (x < 0.0f) a = (y < 0.0f) ? a - PI_F : a + PI_F;
    return a;
}

/* asin(x), Cephes asinf; |x| > 1 clamps (GLSL: undefined) */
VKP_HD float vkp_asinf(float xx)
{
    float a = xx < 0.0f ? -xx : xx;
    if (a > 1.0f) a = 1.0f;
    float z, x;
    int big = a > 0.5f;
    if (big) {
        z = 0.5f * (1.0f - a);
        x = vkp_sqrtf(z);
    } else {
        x = a;
        z = x * x;
    }
    float p = vkp_fmaf(vkp_fmaf(vkp_fmaf(vkp_fmaf(vkp_fmaf(4.2163199048e-2f, z, 2.4181311049e-2f), z, 4.5470025998e-2f), z,
                                         7.4953002686e-2f), z, 1.6666752422e-1f) * z, x, x);
    if (big) {
        p = p + p;
        p = 1.57079632679489661923f - p;
    }
    return xx < 0.0f ? -p : p;
}

/* log2(x) for finite x > 0 (normal numbers; denormals are treated as 0 -> -inf) */
VKP_HD float vkp_log2f(float x)
{
    uint32_t u = vkp_f2u(x);
    int e = (int)((u >> 23) & 0xffu) - 127;
    if (((u >> 23) & 0xffu) == 0u) return -INFINITY;
    float m = vkp_u2f((u & 0x007fffffu) | 0x3f800000u);  /* [1,2) */
    if (m > 1.41421356237f) { m = m * 0.5f; e += 1; }      /* [sqrt(.5), sqrt(2)) */
    float f = m - 1.0f;
    float z = f * f;
    /* Cephes logf polynomial: ln(1+f) = f - z/2 + f*z*P(f) */
    float p = 7.0376836292e-2f;
    p = vkp_fmaf(p, f, -1.1514610310e-1f); p = vkp_fmaf(p, f, 1.1676998740e-1f); p = vkp_fmaf(p, f, -1.2420140846e-1f);
    p = vkp_fmaf(p, f, 1.4249322787e-1f); p = vkp_fmaf(p, f, -1.6668057665e-1f); p = vkp_fmaf(p, f, 2.0000714765e-1f);
    p = vkp_fmaf(p, f, -2.4999993993e-1f); p = vkp_fmaf(p, f, 3.3333331174e-1f);
    float ln = (f * z) * p;
    ln = vkp_fmaf(-0.5f, z, ln);
    ln = ln + f;
    /* ln -> log2: split 1/ln2 for accuracy */
    float r = ln * 1.44269504088896341f;
    return r + (float)e;
}

/* 2^x */
VKP_HD float vkp_exp2f(float x)
{
    if (x >= 128.0f) return INFINITY;
    if (x < -126.0f) return 0.0f;
    float kf = vkp_rintf_small(x);
    int   k  = (int)kf;
    float f  = x - kf;                 /* [-0.5, 0.5] exact */
    /* Cephes exp2f polynomial */
    float p = 1.535336188319500e-4f;
    p = vkp_fmaf(p, f, 1.339887440266574e-3f); p = vkp_fmaf(p, f, 9.618437357674640e-3f); p = vkp_fmaf(p, f, 5.550332471162809e-2f);
    p = vkp_fmaf(p, f, 2.402264791363012e-1f); p = vkp_fmaf(p, f, 6.931472028550421e-1f); p = vkp_fmaf(p, f, 1.0f);
    /* scale by 2^k through the exponent field; k in [-126,128] */
    int e = k + 127;
    if (e <= 0) return 0.0f;
    if (e >= 255) return INFINITY;
    return p * vkp_u2f((uint32_t)e << 23);
}

/* GLSL pow(x, y) for x >= 0.  pow(0, y>0) = 0, pow(x, 0) = 1.  The two small integer exponents the shaders use
 * (helpers.glsl:71 pow(1-cos, 5.0); hybridRaygen.rgen:143 pow(eta, 2.0)) are plain products. */
VKP_HD float vkp_powf(float x, float y)
{
    if (y == 0.0f) return 1.0f;
    if (!(x > 0.0f)) return 0.0f;
    if (x == 1.0f) return 1.0f;
    if (y == 2.0f) return x * x;
    if (y == 5.0f) { float x2 = x * x; return (x2 * x2) * x; }
    return vkp_exp2f(y * vkp_log2f(x));
}

/* float -> unorm8 with round-half-even (Vulkan float->UNORM conversion) */
VKP_HD uint8_t vkp_unorm8(float x)
{
    if (!(x > 0.0f)) return 0;     /* also NaN -> 0 */
    if (x >= 1.0f) return 255;
    float v = vkp_rintf_small(x * 255.0f);
    return (uint8_t)(int)v;
}
VKP_HD float vkp_from_unorm8(uint8_t b) { return (float)b / 255.0f; }

/* float -> IEEE binary16 bits, round-to-nearest-even (matches __float2half_rn) */
VKP_HD uint16_t vkp_f32_to_f16(float f)
{
#if defined(__CUDA_ARCH__)
    /* cvt.rn.f16.f32: the same round-to-nearest-even conversion in one instruction (checked bit for bit by the GPU tests) */
    { unsigned short hw; asm("{ .reg .b16 t; cvt.rn.f16.f32 t, %1; mov.b16 %0, t; }" : "=h"(hw) : "f"(f)); return (uint16_t)hw; }
#endif
    uint32_t x = vkp_f2u(f);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) {                       /* inf / nan */
        return (uint16_t)(sign | 0x7c00u | ((ax > 0x7f800000u) ? 0x0200u : 0u));
    }
    if (ax >= 0x477ff000u) {                       /* >= 65520 rounds to inf */
        return (uint16_t)(sign | 0x7c00u);
    }
    if (ax < 0x38800000u) {                        /* subnormal half or zero */
        if (ax < 0x33000000u) return (uint16_t)sign;   /* < 2^-25 -> 0 (ties at 2^-25 -> 0 even) */
        uint32_t e = ax >> 23;                     /* biased exp, 102..112 */
        uint32_t m = (ax & 0x007fffffu) | 0x00800000u;
        uint32_t shift = 126u - e;                 /* 14..24 */
        uint32_t h = m >> shift;
        uint32_t rem = m & ((1u << shift) - 1u);
        uint32_t half = 1u << (shift - 1u);
        if (rem > half || (rem == half && (h & 1u))) h++;
        return (uint16_t)(sign | h);
    }
    uint32_t h = ((ax - 0x38000000u) >> 13);       /* rebias exponent, truncate mantissa */
    uint32_t rem = ax & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;
    return (uint16_t)(sign | h);
}

VKP_HD float vkp_f16_to_f32(uint16_t h)
{
#if defined(__CUDA_ARCH__)
    { float r; asm("{ .reg .b16 t; mov.b16 t, %1; cvt.f32.f16 %0, t; }" : "=f"(r) : "h"((unsigned short)h)); return r; }
#endif
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1fu;
    uint32_t m = h & 0x3ffu;
    if (e == 0u) {
        if (m == 0u) return vkp_u2f(sign);
        /* subnormal: value = m * 2^-24 (exact in fp32) */
        float v = (float)m * 5.9604644775390625e-08f;
        return sign ? -v : v;
    }
    if (e == 31u) return vkp_u2f(sign | 0x7f800000u | (m << 13));
    return vkp_u2f(sign | ((e + 112u) << 23) | (m << 13));
}

/* ---- vector primitives of contract v2 (scalars in, scalar out: float3 on the device and orc::vec3 on the host are built from them) ---- */
/* a . b */
VKP_HD float vkp_dot3(float ax, float ay, float az, float bx, float by, float bz)
{
    return vkp_fmaf(az, bz, vkp_fmaf(ay, by, ax * bx));
}
/* a*b - c*d: one component of a cross product */
VKP_HD float vkp_diff_of_products(float a, float b, float c, float d) { return vkp_fmaf(a, b, -(c * d)); }
/* m0*x + m1*y + m2*z + m3: one row of a 3x4 affine map applied to a point */
VKP_HD float vkp_row_point(float m0, float m1, float m2, float m3, float x, float y, float z)
{
    return vkp_fmaf(m2, z, vkp_fmaf(m1, y, vkp_fmaf(m0, x, m3)));
}
/* m0*x + m1*y + m2*z + m3*w: one row of a 4x4 matrix applied to a vec4 */
VKP_HD float vkp_row_vec4(float m0, float m1, float m2, float m3, float x, float y, float z, float w)
{
    return vkp_fmaf(m3, w, vkp_fmaf(m2, z, vkp_fmaf(m1, y, m0 * x)));
}
/* GLSL mix(x, y, a) = x*(1-a) + y*a */
VKP_HD float vkp_mixf(float x, float y, float a) { return vkp_fmaf(y, a, x * (1.0f - a)); }
/* c - a*b (reflect / refract tails, Cody-Waite steps) */
VKP_HD float vkp_nmsubf(float a, float b, float c) { return vkp_fmaf(-a, b, c); }

#endif /* VKP_DETMATH_H */
