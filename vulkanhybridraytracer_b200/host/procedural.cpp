// procedural.cpp -- deterministic stand-ins for the reference's git-ignored assets
// (.gitignore:1-11: data/meshes, data/textures).  SURVEY.md App. C lists them:
//   sphere.obj          -> UV sphere, radius 1, 64 x 32
//   lucy.obj            -> statue-like displaced icosphere, 7 subdivisions (327 680 triangles), pre-scaled x100
//   DamagedHelmet.gltf  -> icosphere, 5 subdivisions (20 480 triangles) + procedural RGBA8 texture set
//   sky / irradiance    -> analytic gradient RGBA8 textures (2048x1024, 64x32)
// Everything is integer-hash noise and +,-,*,/,sqrt so the arrays are identical on every machine.
#include "procedural.h"

#include <cmath>
#include <cstdint>
#include <map>
#include <utility>

namespace proc {

static inline uint32_t hash_u32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
static inline float hash3(int x, int y, int z, uint32_t seed)
{
    uint32_t h = hash_u32((uint32_t)x * 0x8da6b343u ^ hash_u32((uint32_t)y * 0xd8163841u ^ hash_u32((uint32_t)z * 0xcb1ab31fu ^ seed)));
    return (float)(h & 0xffffffu) / 16777216.0f;   // [0,1)
}
static inline float smooth(float t) { return t * t * (3.0f - 2.0f * t); }

// trilinear value noise in [0,1)
static float vnoise(float x, float y, float z, uint32_t seed)
{
    float fx = std::floor(x), fy = std::floor(y), fz = std::floor(z);
    int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    float tx = smooth(x - fx), ty = smooth(y - fy), tz = smooth(z - fz);
    float c[2][2][2];
    for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) for (int d = 0; d < 2; d++) c[a][b][d] = hash3(ix + a, iy + b, iz + d, seed);
    float x00 = c[0][0][0] + (c[1][0][0] - c[0][0][0]) * tx, x10 = c[0][1][0] + (c[1][1][0] - c[0][1][0]) * tx;
    float x01 = c[0][0][1] + (c[1][0][1] - c[0][0][1]) * tx, x11 = c[0][1][1] + (c[1][1][1] - c[0][1][1]) * tx;
    float y0 = x00 + (x10 - x00) * ty, y1 = x01 + (x11 - x01) * ty;
    return y0 + (y1 - y0) * tz;
}
static float fbm(float x, float y, float z, int octaves, uint32_t seed)
{
    float s = 0.0f, a = 0.5f, f = 1.0f;
    for (int o = 0; o < octaves; o++) { s += a * (vnoise(x * f, y * f, z * f, seed + 101u * (uint32_t)o) - 0.5f); a *= 0.5f; f *= 2.0f; }
    return s;   // about [-0.5, 0.5]
}

// polynomial atan2 / asin are not needed for geometry: uv only has to be deterministic, so use an
// octahedral-style mapping built from +,-,*,/ only.
static vkm::vec2 dir_to_uv(vkm::vec3 d)
{
    float ax = std::fabs(d.x), ay = std::fabs(d.y), az = std::fabs(d.z);
    float s = ax + ay + az; if (s == 0.0f) s = 1.0f;
    float px = d.x / s, pz = d.z / s;
    if (d.y < 0.0f) {
        float ox = (1.0f - std::fabs(pz)) * (px >= 0.0f ? 1.0f : -1.0f);
        float oz = (1.0f - std::fabs(px)) * (pz >= 0.0f ? 1.0f : -1.0f);
        px = ox; pz = oz;
    }
    return vkm::vec2(px * 0.5f + 0.5f, pz * 0.5f + 0.5f);
}

void uv_sphere(int segments, int rings, std::vector<Vertex>& v, std::vector<uint32_t>& idx)
{
    v.clear(); idx.clear();
    const double PI = 3.14159265358979323846;
    for (int r = 0; r <= rings; r++) {
        double theta = PI * (double)r / (double)rings;
        double st = std::sin(theta), ct = std::cos(theta);
        if (r == 0) { st = 0.0; ct = 1.0; }
        if (r == rings) { st = 0.0; ct = -1.0; }
        for (int s = 0; s <= segments; s++) {
            double phi = 2.0 * PI * (double)s / (double)segments;
            double sp = std::sin(phi), cp = std::cos(phi);
            if (s == segments) { sp = 0.0; cp = 1.0; }
            if (s == 0) { sp = 0.0; cp = 1.0; }
            Vertex vx;
            vx.position = vkm::vec3((float)(st * cp), (float)ct, (float)(st * sp));
            vx.normal = vx.position;
            vx.color = vkm::vec3(1.0f);
            vx.uv = vkm::vec2((float)s / (float)segments, (float)r / (float)rings);
            v.push_back(vx);
        }
    }
    const uint32_t stride = (uint32_t)segments + 1u;
    for (int r = 0; r < rings; r++)
        for (int s = 0; s < segments; s++) {
            uint32_t a = (uint32_t)r * stride + (uint32_t)s, b = a + 1u, c = a + stride, d = c + 1u;
            if (r != 0) { idx.push_back(a); idx.push_back(b); idx.push_back(c); }           // skip the degenerate pole triangles
            if (r != rings - 1) { idx.push_back(b); idx.push_back(d); idx.push_back(c); }
        }
}

static void icosphere_unit(int subdiv, std::vector<vkm::vec3>& pos, std::vector<uint32_t>& idx)
{
    const float t = (1.0f + std::sqrt(5.0f)) / 2.0f;
    const float base[12][3] = {{-1, t, 0}, {1, t, 0}, {-1, -t, 0}, {1, -t, 0}, {0, -1, t}, {0, 1, t},
                               {0, -1, -t}, {0, 1, -t}, {t, 0, -1}, {t, 0, 1}, {-t, 0, -1}, {-t, 0, 1}};
    static const uint32_t faces[20][3] = {{0, 11, 5}, {0, 5, 1}, {0, 1, 7}, {0, 7, 10}, {0, 10, 11}, {1, 5, 9}, {5, 11, 4},
                                          {11, 10, 2}, {10, 7, 6}, {7, 1, 8}, {3, 9, 4}, {3, 4, 2}, {3, 2, 6}, {3, 6, 8},
                                          {3, 8, 9}, {4, 9, 5}, {2, 4, 11}, {6, 2, 10}, {8, 6, 7}, {9, 8, 1}};
    pos.clear(); idx.clear();
    for (auto& b : base) pos.push_back(vkm::normalize(vkm::vec3(b[0], b[1], b[2])));
    for (auto& f : faces) { idx.push_back(f[0]); idx.push_back(f[1]); idx.push_back(f[2]); }
    for (int s = 0; s < subdiv; s++) {
        std::map<std::pair<uint32_t, uint32_t>, uint32_t> cache;
        auto midpoint = [&](uint32_t a, uint32_t b) -> uint32_t {
            std::pair<uint32_t, uint32_t> key = a < b ? std::make_pair(a, b) : std::make_pair(b, a);
            auto it = cache.find(key);
            if (it != cache.end()) return it->second;
            vkm::vec3 m = vkm::normalize((pos[key.first] + pos[key.second]) * 0.5f);
            pos.push_back(m);
            uint32_t id = (uint32_t)pos.size() - 1u;
            cache[key] = id;
            return id;
        };
        std::vector<uint32_t> next; next.reserve(idx.size() * 4);
        for (size_t i = 0; i < idx.size(); i += 3) {
            uint32_t a = idx[i], b = idx[i + 1], c = idx[i + 2];
            uint32_t ab = midpoint(a, b), bc = midpoint(b, c), ca = midpoint(c, a);
            const uint32_t tri[12] = {a, ab, ca, b, bc, ab, c, ca, bc, ab, bc, ca};
            next.insert(next.end(), tri, tri + 12);
        }
        idx.swap(next);
    }
}

static void finish_mesh(const std::vector<vkm::vec3>& unit, const std::vector<vkm::vec3>& pos, const std::vector<uint32_t>& idx,
                        std::vector<Vertex>& v)
{
    std::vector<vkm::vec3> nrm(pos.size());
    for (size_t i = 0; i < idx.size(); i += 3) {
        vkm::vec3 a = pos[idx[i]], b = pos[idx[i + 1]], c = pos[idx[i + 2]];
        vkm::vec3 n = vkm::cross(b - a, c - a);          // area-weighted
        nrm[idx[i]] += n; nrm[idx[i + 1]] += n; nrm[idx[i + 2]] += n;
    }
    v.resize(pos.size());
    for (size_t i = 0; i < pos.size(); i++) {
        v[i].position = pos[i];
        float l = vkm::length(nrm[i]);
        v[i].normal = l > 0.0f ? nrm[i] / l : unit[i];
        v[i].color = vkm::vec3(1.0f);
        v[i].uv = dir_to_uv(unit[i]);
    }
}

void lucy_standin(std::vector<Vertex>& v, std::vector<uint32_t>& idx, int subdiv)
{
    std::vector<vkm::vec3> unit; icosphere_unit(subdiv, unit, idx);
    std::vector<vkm::vec3> pos(unit.size());
    for (size_t i = 0; i < unit.size(); i++) {
        vkm::vec3 p = unit[i];
        // a tapered, waisted column with folds: tall in y, displaced by multi-octave noise
        float h = p.y * 0.5f + 0.5f;                                  // 0 bottom .. 1 top
        float profile = 0.55f + 0.35f * (1.0f - h) * (1.0f - h) - 0.25f * h * (1.0f - h) * 4.0f * 0.3f + 0.18f * (h > 0.8f ? (h - 0.8f) * 5.0f : 0.0f);
        float n = fbm(p.x * 2.5f + 7.1f, p.y * 2.5f - 3.3f, p.z * 2.5f + 1.7f, 5, 1234u);
        float folds = 0.06f * vnoise(p.x * 9.0f, p.y * 3.0f, p.z * 9.0f, 77u);
        float r = profile * (1.0f + 0.45f * n) + folds;
        vkm::vec3 q(p.x * r * 0.62f, (p.y * (1.0f + 0.12f * n)) * 1.25f + 1.27f, p.z * r * 0.62f);
        if (q.y < 0.0f) q.y = 0.0f;
        pos[i] = q * 100.0f;                                         // the scene applies S(0.01) (scene.cpp:143)
    }
    finish_mesh(unit, pos, idx, v);
}

void helmet_standin(std::vector<Vertex>& v, std::vector<uint32_t>& idx, int subdiv)
{
    std::vector<vkm::vec3> unit; icosphere_unit(subdiv, unit, idx);
    std::vector<vkm::vec3> pos(unit.size());
    for (size_t i = 0; i < unit.size(); i++) {
        vkm::vec3 p = unit[i];
        float n = fbm(p.x * 3.0f + 11.0f, p.y * 3.0f + 5.0f, p.z * 3.0f - 2.0f, 4, 4321u);
        float visor = (p.z > 0.55f && std::fabs(p.y) < 0.35f) ? -0.08f : 0.0f;   // a dented "visor"
        float r = 0.92f + 0.10f * n + visor;
        pos[i] = vkm::vec3(p.x * r, p.y * r * 1.05f, p.z * r);
    }
    finish_mesh(unit, pos, idx, v);
}

void displaced_icosphere(std::vector<Vertex>& v, std::vector<uint32_t>& idx, int subdiv, uint32_t seed, float amount)
{
    std::vector<vkm::vec3> unit; icosphere_unit(subdiv, unit, idx);
    std::vector<vkm::vec3> pos(unit.size());
    for (size_t i = 0; i < unit.size(); i++) {
        vkm::vec3 p = unit[i];
        float n = fbm(p.x * 2.0f + 3.0f, p.y * 2.0f + 1.0f, p.z * 2.0f - 5.0f, 5, seed);
        pos[i] = p * (1.0f + amount * n);
    }
    finish_mesh(unit, pos, idx, v);
}

static inline uint8_t u8(float x) { x = x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x); return (uint8_t)(x * 255.0f + 0.5f); }

// analytic sky: horizon haze -> zenith blue, a warm sun blob, dark ground; v = 0 is straight up (miss.rmiss:16)
void sky_texture(int w, int h, std::vector<uint8_t>& rgba)
{
    rgba.resize(4 * (size_t)w * h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float v = ((float)y + 0.5f) / (float)h, u = ((float)x + 0.5f) / (float)w;
            float up = 1.0f - 2.0f * v;                       // ~ +1 zenith, -1 nadir (linear in v, not sin)
            float r, g, b;
            if (up >= 0.0f) {
                float t = std::sqrt(up);
                r = 0.78f + (0.18f - 0.78f) * t; g = 0.86f + (0.38f - 0.86f) * t; b = 0.95f + (0.80f - 0.95f) * t;
            } else {
                float t = std::sqrt(-up);
                r = 0.55f + (0.16f - 0.55f) * t; g = 0.52f + (0.14f - 0.52f) * t; b = 0.48f + (0.12f - 0.48f) * t;
            }
            float du = u - 0.70f, dv = v - 0.30f;
            float d2 = du * du * 4.0f + dv * dv;
            float sun = d2 < 0.0004f ? 1.0f : 0.0004f / d2;
            sun = sun > 1.0f ? 1.0f : sun;
            r += sun * 0.9f; g += sun * 0.8f; b += sun * 0.5f;
            float cloud = vnoise(u * 24.0f, v * 12.0f, 0.5f, 99u);
            if (up > 0.05f && cloud > 0.62f) { float c = (cloud - 0.62f) * 1.6f; r += c; g += c; b += c; }
            uint8_t* p = &rgba[4 * ((size_t)y * w + x)];
            p[0] = u8(r); p[1] = u8(g); p[2] = u8(b); p[3] = 255;
        }
}

void irradiance_texture(int w, int h, std::vector<uint8_t>& rgba)
{
    rgba.resize(4 * (size_t)w * h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float v = ((float)y + 0.5f) / (float)h;
            float up = 1.0f - 2.0f * v;
            float t = up * 0.5f + 0.5f;
            uint8_t* p = &rgba[4 * ((size_t)y * w + x)];
            p[0] = u8(0.10f + 0.12f * t); p[1] = u8(0.10f + 0.16f * t); p[2] = u8(0.09f + 0.24f * t); p[3] = 255;
        }
}

void helmet_textures(int w, int h, std::vector<uint8_t>& albedo, std::vector<uint8_t>& emissive, std::vector<uint8_t>& metal_rough, std::vector<uint8_t>& normal)
{
    albedo.resize(4 * (size_t)w * h); emissive.resize(4 * (size_t)w * h); metal_rough.resize(4 * (size_t)w * h); normal.resize(4 * (size_t)w * h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float u = ((float)x + 0.5f) / (float)w, v = ((float)y + 0.5f) / (float)h;
            float n = vnoise(u * 16.0f, v * 16.0f, 0.0f, 555u);
            float scratches = vnoise(u * 90.0f, v * 6.0f, 3.0f, 777u);
            int panel = ((int)(u * 8.0f) + (int)(v * 8.0f)) & 1;
            size_t o = 4 * ((size_t)y * w + x);
            float base = panel ? 0.55f : 0.38f;
            albedo[o + 0] = u8(base + 0.25f * n); albedo[o + 1] = u8(base * 0.9f + 0.2f * n); albedo[o + 2] = u8(base * 0.8f + 0.15f * n); albedo[o + 3] = 255;
            float stripe = (std::fabs(v - 0.5f) < 0.015f || std::fabs(u - 0.25f) < 0.008f) ? 1.0f : 0.0f;
            emissive[o + 0] = u8(0.05f * stripe); emissive[o + 1] = u8(0.6f * stripe); emissive[o + 2] = u8(0.9f * stripe); emissive[o + 3] = 255;
            // glTF packing read by the shaders as (x, roughness = .y, metallic = .z)
            metal_rough[o + 0] = 0; metal_rough[o + 1] = u8(0.25f + 0.5f * scratches); metal_rough[o + 2] = u8(panel ? 0.9f : 0.6f); metal_rough[o + 3] = 255;
            // tangent-space normal map (the glTF sample has one, vk_mesh.cpp:783-786): dents from the gradient of a noise field
            const float e = 1.0f / 512.0f;
            float hx = vnoise((u + e) * 24.0f, v * 24.0f, 7.0f, 999u) - vnoise((u - e) * 24.0f, v * 24.0f, 7.0f, 999u);
            float hy = vnoise(u * 24.0f, (v + e) * 24.0f, 7.0f, 999u) - vnoise(u * 24.0f, (v - e) * 24.0f, 7.0f, 999u);
            float nx = -6.0f * hx, ny = -6.0f * hy, nz = 1.0f;
            float inv = 1.0f / std::sqrt(nx * nx + ny * ny + nz * nz);
            normal[o + 0] = u8(nx * inv * 0.5f + 0.5f); normal[o + 1] = u8(ny * inv * 0.5f + 0.5f); normal[o + 2] = u8(nz * inv * 0.5f + 0.5f); normal[o + 3] = 255;
        }
}

} // namespace proc
