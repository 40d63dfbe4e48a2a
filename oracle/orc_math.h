/*
 * orc_math.h -- TEST INFRASTRUCTURE (CPU oracle).  Not part of the shipped product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use oracle/.
 *
 * GLSL built-ins the reference shaders use, restated as scalar C++ with ONE fixed evaluation order
 * (left-to-right as written in the GLSL source, no FMA contraction by the compiler: compile with
 * -ffp-contract=off; the only fused operations are the ones written as vkp_fmaf in the contract-v2
 * primitives of include/vkp_detmath.h).  Elementary functions come from include/vkp_detmath.h (the numerics contract
 * of the C-ABI) unless ORC_USE_LIBM is defined, in which case glibc's are used (independence
 * check: the tests then allow epsilon-tie flips).
 */
#ifndef ORC_MATH_H
#define ORC_MATH_H

#include <cmath>
#include <cstdint>
#include "../include/vkp_detmath.h"

namespace orc {

#ifdef ORC_USE_LIBM
inline float e_sin(float x) { return sinf(x); }
inline float e_cos(float x) { return cosf(x); }
inline float e_atan2(float y, float x) { return atan2f(y, x); }
inline float e_asin(float x) { return asinf(x < -1.f ? -1.f : (x > 1.f ? 1.f : x)); }
inline float e_pow(float x, float y) { if (y == 0.f) return 1.f; if (!(x > 0.f)) return 0.f; return powf(x, y); }
#else
inline float e_sin(float x) { return vkp_sinf(x); }
inline float e_cos(float x) { return vkp_cosf(x); }
inline float e_atan2(float y, float x) { return vkp_atan2f(y, x); }
inline float e_asin(float x) { return vkp_asinf(x); }
inline float e_pow(float x, float y) { return vkp_powf(x, y); }
#endif

struct vec2 { float x, y; };
struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
};
struct vec4 { float x, y, z, w; };

inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(vec3 a, vec3 b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(vec3 a, vec3 b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, vec3 a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(vec3 a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator+(vec3 a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator-(vec3 a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }
inline vec3 operator-(float s, vec3 a) { return vec3(s - a.x, s - a.y, s - a.z); }
inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline vec3& operator*=(vec3& a, vec3 b) { a = a * b; return a; }
inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }
inline bool operator==(vec3 a, vec3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

/* contract v2 (include/vkp_detmath.h): dot, cross, matrix rows, mix, reflect / refract are FMA forms shared with the CUDA kernels */
inline float dot(vec3 a, vec3 b) { return vkp_dot3(a.x, a.y, a.z, b.x, b.y, b.z); }
inline vec3 cross(vec3 a, vec3 b)
{
    return vec3(vkp_diff_of_products(a.y, b.z, a.z, b.y), vkp_diff_of_products(a.z, b.x, a.x, b.z), vkp_diff_of_products(a.x, b.y, a.y, b.x));
}
inline float length(vec3 a) { return sqrtf(dot(a, a)); }
inline vec3 normalize(vec3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
inline float maxf(float a, float b) { return fmaxf(a, b); }
inline float minf(float a, float b) { return fminf(a, b); }
inline vec3 max3(vec3 a, float s) { return vec3(fmaxf(a.x, s), fmaxf(a.y, s), fmaxf(a.z, s)); }
inline float mixf(float x, float y, float a) { return vkp_mixf(x, y, a); }
inline vec3 mix(vec3 x, vec3 y, float a) { return vec3(vkp_mixf(x.x, y.x, a), vkp_mixf(x.y, y.y, a), vkp_mixf(x.z, y.z, a)); }
inline vec3 pow3(vec3 a, float e) { return vec3(e_pow(a.x, e), e_pow(a.y, e), e_pow(a.z, e)); }
inline vec3 reflect(vec3 I, vec3 N) { const float k = 2.0f * dot(N, I); return vec3(vkp_nmsubf(N.x, k, I.x), vkp_nmsubf(N.y, k, I.y), vkp_nmsubf(N.z, k, I.z)); }
inline vec3 refract(vec3 I, vec3 N, float eta)
{
    float d = dot(N, I);
    float k = vkp_nmsubf(eta * eta, vkp_nmsubf(d, d, 1.0f), 1.0f);
    if (k < 0.0f) return vec3(0.0f);
    const float s = vkp_fmaf(eta, d, sqrtf(k));
    return vec3(vkp_nmsubf(N.x, s, I.x * eta), vkp_nmsubf(N.y, s, I.y * eta), vkp_nmsubf(N.z, s, I.z * eta));
}

/* column-major mat4 (GLM): m[c*4+r] */
inline vec3 mul_point(const float* m, vec3 p)   /* (M * vec4(p,1)).xyz */
{
    return vec3(vkp_row_point(m[0], m[4], m[8], m[12], p.x, p.y, p.z),
                vkp_row_point(m[1], m[5], m[9], m[13], p.x, p.y, p.z),
                vkp_row_point(m[2], m[6], m[10], m[14], p.x, p.y, p.z));
}
inline vec3 mul_dir(const float* m, vec3 d)     /* (M * vec4(d,0)).xyz */
{
    return vec3(vkp_dot3(m[0], m[4], m[8], d.x, d.y, d.z),
                vkp_dot3(m[1], m[5], m[9], d.x, d.y, d.z),
                vkp_dot3(m[2], m[6], m[10], d.x, d.y, d.z));
}
inline vec4 mul_vec4(const float* m, vec4 v)
{
    vec4 r;
    r.x = vkp_row_vec4(m[0], m[4], m[8], m[12], v.x, v.y, v.z, v.w);
    r.y = vkp_row_vec4(m[1], m[5], m[9], m[13], v.x, v.y, v.z, v.w);
    r.z = vkp_row_vec4(m[2], m[6], m[10], m[14], v.x, v.y, v.z, v.w);
    r.w = vkp_row_vec4(m[3], m[7], m[11], m[15], v.x, v.y, v.z, v.w);
    return r;
}
/* mat3(transpose(M)) * n : row i of the result = dot(column i of M, n) */
inline vec3 mul_transpose3(const float* m, vec3 n)
{
    return vec3(vkp_dot3(m[0], m[1], m[2], n.x, n.y, n.z),
                vkp_dot3(m[4], m[5], m[6], n.x, n.y, n.z),
                vkp_dot3(m[8], m[9], m[10], n.x, n.y, n.z));
}

/* random.glsl:3-17 */
inline uint32_t tea(uint32_t val0, uint32_t val1)
{
    uint32_t v0 = val0, v1 = val1, s0 = 0;
    for (uint32_t n = 0; n < 16; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
/* random.glsl:21-27 */
inline uint32_t lcg(uint32_t& prev)
{
    const uint32_t LCG_A = 1664525u, LCG_C = 1013904223u;
    prev = LCG_A * prev + LCG_C;
    return prev & 0x00FFFFFFu;
}
/* random.glsl:30-33 */
inline float rnd(uint32_t& prev) { return (float)lcg(prev) / (float)0x01000000; }

/* storage conversions (renderer.cpp:255-261: RGBA16F / RGBA8 attachments) */
inline uint16_t to_half(float f) { return vkp_f32_to_f16(f); }
inline float from_half(uint16_t h) { return vkp_f16_to_f32(h); }
inline uint8_t to_unorm8(float f) { return vkp_unorm8(f); }
inline float from_unorm8(uint8_t b) { return (float)b / 255.0f; }

} // namespace orc
#endif
