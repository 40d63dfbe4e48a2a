// vkm.h -- the subset of GLM the reference's host code uses (glm::vec3/vec4/mat4, translate, rotate,
// scale, lookAt, perspective, inverse, radians), re-implemented because GLM is un-vendored upstream
// (.gitignore:7) and absent here.  Column-major, right-handed, clip z in [-1,1] (GLM defaults; the
// reference defines no GLM_FORCE_* macro, vk_types.h:16).
#pragma once
#include <cmath>
#include <cstring>

namespace vkm {

struct vec2 { float x = 0, y = 0; vec2() = default; vec2(float a, float b) : x(a), y(b) {} };

struct vec3 {
    float x = 0, y = 0, z = 0;
    vec3() = default;
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, vec3 a) { return a * s; }
inline vec3 operator/(vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline vec3& operator-=(vec3& a, vec3 b) { a = a - b; return a; }
inline bool operator==(vec3 a, vec3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
inline vec3 normalize(vec3 a) { return a / length(a); }

struct vec4 {
    float x = 0, y = 0, z = 0, w = 0;
    vec4() = default;
    explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(vec3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
    explicit operator vec3() const { return {x, y, z}; }
};
inline bool operator==(vec4 a, vec4 b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }
inline vec4 operator*(vec4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline vec4 operator+(vec4 a, vec4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }

struct mat4 {
    vec4 c[4];
    mat4() = default;
    explicit mat4(float d) { c[0].x = d; c[1].y = d; c[2].z = d; c[3].w = d; }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
    const float* data() const { return &c[0].x; }
    float* data() { return &c[0].x; }
};
inline vec4 operator*(const mat4& m, vec4 v) { return m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * v.w; }
inline mat4 operator*(const mat4& a, const mat4& b)
{
    mat4 r;
    for (int j = 0; j < 4; j++) r[j] = a * b[j];
    return r;
}
inline bool operator==(const mat4& a, const mat4& b) { return std::memcmp(a.data(), b.data(), 64) == 0; }
inline bool operator!=(const mat4& a, const mat4& b) { return !(a == b); }

inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }

inline mat4 translate(const mat4& m, vec3 v)
{
    mat4 r = m;
    r[3] = m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3];
    return r;
}
inline mat4 scale(const mat4& m, vec3 v)
{
    mat4 r;
    r[0] = m[0] * v.x; r[1] = m[1] * v.y; r[2] = m[2] * v.z; r[3] = m[3];
    return r;
}
inline mat4 rotate(const mat4& m, float angle, vec3 v)
{
    const float c = std::cos(angle), s = std::sin(angle);
    vec3 axis = normalize(v);
    vec3 temp = axis * (1.0f - c);
    mat4 rot(1.0f);
    rot[0].x = c + temp.x * axis.x;          rot[0].y = temp.x * axis.y + s * axis.z; rot[0].z = temp.x * axis.z - s * axis.y;
    rot[1].x = temp.y * axis.x - s * axis.z; rot[1].y = c + temp.y * axis.y;          rot[1].z = temp.y * axis.z + s * axis.x;
    rot[2].x = temp.z * axis.x + s * axis.y; rot[2].y = temp.z * axis.y - s * axis.x; rot[2].z = c + temp.z * axis.z;
    mat4 r;
    r[0] = m[0] * rot[0].x + m[1] * rot[0].y + m[2] * rot[0].z;
    r[1] = m[0] * rot[1].x + m[1] * rot[1].y + m[2] * rot[1].z;
    r[2] = m[0] * rot[2].x + m[1] * rot[2].y + m[2] * rot[2].z;
    r[3] = m[3];
    return r;
}
inline mat4 lookAt(vec3 eye, vec3 center, vec3 up)
{
    const vec3 f = normalize(center - eye);
    const vec3 s = normalize(cross(f, up));
    const vec3 u = cross(s, f);
    mat4 r(1.0f);
    r[0].x = s.x; r[1].x = s.y; r[2].x = s.z;
    r[0].y = u.x; r[1].y = u.y; r[2].y = u.z;
    r[0].z = -f.x; r[1].z = -f.y; r[2].z = -f.z;
    r[3].x = -dot(s, eye); r[3].y = -dot(u, eye); r[3].z = dot(f, eye);
    return r;
}
inline mat4 perspective(float fovy, float aspect, float zNear, float zFar)
{
    const float t = std::tan(fovy / 2.0f);
    mat4 r;   // zero
    r[0].x = 1.0f / (aspect * t);
    r[1].y = 1.0f / t;
    r[2].z = -(zFar + zNear) / (zFar - zNear);
    r[2].w = -1.0f;
    r[3].z = -(2.0f * zFar * zNear) / (zFar - zNear);
    return r;
}
// general 4x4 inverse by cofactors, accumulated in double then rounded once
inline mat4 inverse(const mat4& mm)
{
    double m[16], inv[16];
    for (int i = 0; i < 16; i++) m[i] = mm.data()[i];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    mat4 r;
    double id = 1.0 / det;
    for (int i = 0; i < 16; i++) r.data()[i] = (float)(inv[i] * id);
    return r;
}

} // namespace vkm
