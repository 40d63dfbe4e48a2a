/*
 * oracle.cpp -- TEST INFRASTRUCTURE: scalar CPU restatement of the reference's hot path.
 *
 * NOT shipped, NOT on the product path.  Only tests/, __graft_entry__.smoke() and bench.py
 * (cpu_baseline / --impl reference) may load liboracle.so.
 *
 * PINNING.  The reference ships no tests or golden vectors and cannot be built or run here, but it
 * ships its COMPILED shaders (data/shaders/output/*.spv).  oracle/spirv_interp.py executes them and
 * tests/test_oracle_spirv.py holds every shading stage below to their outputs (committed under
 * tests/golden/spirv_*.npz with the generating script): G-buffer, shadow ray generation, blur.comp,
 * the resolve + hit + miss shaders, Raytraced mode, deferred.frag, postFragment.frag.  What stays
 * UNPINNED is what the reference does not contain: BVH build, traversal and the ray-triangle test
 * live in the Vulkan driver / RT hardware -- see the intersection contract in oracle.h.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * The code is deliberately plain: one pixel at a time, GLSL evaluation order, OpenMP over rows.
 */
#include "oracle.h"
#include "orc_math.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

/* timing aid for bench.py: process only rows y with y % stride == offset (a representative sample of the frame) */
static int g_row_stride = 1, g_row_offset = 0;

/* ------------------------------------------------------------------------------------------------
 * scene storage + a simple binned-SAH BVH per (mesh, firstIndex, indexCount) range (the reference
 * builds one BLAS per primitive: renderer.cpp:1643-1666, vk_mesh.cpp:424-470)
 * ---------------------------------------------------------------------------------------------- */
namespace {

struct Box {
    float lo[3], hi[3];
    void reset() { for (int k = 0; k < 3; k++) { lo[k] = INFINITY; hi[k] = -INFINITY; } }
    void grow(const float* p) { for (int k = 0; k < 3; k++) { lo[k] = fminf(lo[k], p[k]); hi[k] = fmaxf(hi[k], p[k]); } }
    void grow(const Box& b) { for (int k = 0; k < 3; k++) { lo[k] = fminf(lo[k], b.lo[k]); hi[k] = fmaxf(hi[k], b.hi[k]); } }
    float area() const { float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2]; return 2.f * (dx * dy + dy * dz + dz * dx); }
};

struct BNode { Box box; int32_t left, right; uint32_t first, count; };

struct OMesh { std::vector<vkp_vertex> v; std::vector<uint32_t> idx; };
struct OTex { std::vector<uint8_t> rgba; int w, h; };

struct OBlas {
    int mesh; uint32_t first_index, index_count;
    std::vector<BNode> nodes;
    std::vector<uint32_t> prims;     /* gl_PrimitiveID values in leaf order */
};

struct OInst { vkp_instance rec; int blas; Box world; };

} // namespace

struct orc_scene {
    std::vector<OMesh> meshes;
    std::vector<OTex> textures;
    int sky_tex, irr_tex;
    std::vector<vkp_material> materials;
    std::vector<vkp_light> lights;
    std::vector<OInst> instances;
    std::vector<OBlas> blas;
};

namespace {

inline void tri_vertices(const OMesh& m, uint32_t first_index, uint32_t prim, vec3& a, vec3& b, vec3& c)
{
    const uint32_t* ix = &m.idx[first_index + 3u * prim];
    const float* p0 = m.v[ix[0]].position; const float* p1 = m.v[ix[1]].position; const float* p2 = m.v[ix[2]].position;
    a = vec3(p0[0], p0[1], p0[2]); b = vec3(p1[0], p1[1], p1[2]); c = vec3(p2[0], p2[1], p2[2]);
}

void build_blas(const OMesh& m, OBlas& B)
{
    const uint32_t n = B.index_count / 3u;
    B.prims.resize(n);
    std::vector<Box> pb(n);
    std::vector<float> cen(3 * (size_t)n);
    for (uint32_t i = 0; i < n; i++) {
        B.prims[i] = i;
        vec3 a, b, c; tri_vertices(m, B.first_index, i, a, b, c);
        pb[i].reset(); pb[i].grow(&a.x); pb[i].grow(&b.x); pb[i].grow(&c.x);
        for (int k = 0; k < 3; k++) cen[3 * (size_t)i + k] = 0.5f * (pb[i].lo[k] + pb[i].hi[k]);
    }
    B.nodes.clear(); B.nodes.reserve(2 * (size_t)n / 2 + 2);
    struct Task { int node; uint32_t lo, hi; };
    std::vector<Task> stack;
    BNode root; root.left = root.right = -1; root.first = 0; root.count = n; root.box.reset();
    for (uint32_t i = 0; i < n; i++) root.box.grow(pb[i]);
    B.nodes.push_back(root);
    if (n == 0) return;
    stack.push_back({0, 0, n});
    const int NB = 16;
    while (!stack.empty()) {
        Task t = stack.back(); stack.pop_back();
        uint32_t cnt = t.hi - t.lo;
        if (cnt <= 4) { B.nodes[t.node].first = t.lo; B.nodes[t.node].count = cnt; continue; }
        Box cb; cb.reset();
        for (uint32_t i = t.lo; i < t.hi; i++) cb.grow(&cen[3 * (size_t)B.prims[i]]);
        int bestAxis = -1; int bestSplit = -1; float bestCost = INFINITY;
        for (int ax = 0; ax < 3; ax++) {
            float ext = cb.hi[ax] - cb.lo[ax];
            if (!(ext > 0.f)) continue;
            Box bb[NB]; uint32_t bc[NB];
            for (int k = 0; k < NB; k++) { bb[k].reset(); bc[k] = 0; }
            float scale = (float)NB / ext;
            for (uint32_t i = t.lo; i < t.hi; i++) {
                uint32_t p = B.prims[i];
                int k = (int)((cen[3 * (size_t)p + ax] - cb.lo[ax]) * scale); if (k >= NB) k = NB - 1; if (k < 0) k = 0;
                bb[k].grow(pb[p]); bc[k]++;
            }
            float la[NB], ra[NB]; uint32_t lc[NB], rc[NB];
            Box acc; acc.reset(); uint32_t c = 0;
            for (int k = 0; k < NB - 1; k++) { acc.grow(bb[k]); c += bc[k]; la[k] = c ? acc.area() : 0.f; lc[k] = c; }
            acc.reset(); c = 0;
            for (int k = NB - 1; k > 0; k--) { acc.grow(bb[k]); c += bc[k]; ra[k - 1] = c ? acc.area() : 0.f; rc[k - 1] = c; }
            for (int k = 0; k < NB - 1; k++) {
                if (lc[k] == 0 || rc[k] == 0) continue;
                float cost = la[k] * (float)lc[k] + ra[k] * (float)rc[k];
                if (cost < bestCost) { bestCost = cost; bestAxis = ax; bestSplit = k; }
            }
        }
        uint32_t mid;
        if (bestAxis < 0) {
            mid = t.lo + cnt / 2;   /* all centroids coincide: split in the middle */
        } else {
            float ext = cb.hi[bestAxis] - cb.lo[bestAxis];
            float scale = (float)NB / ext;
            auto it = std::partition(B.prims.begin() + t.lo, B.prims.begin() + t.hi, [&](uint32_t p) {
                int k = (int)((cen[3 * (size_t)p + bestAxis] - cb.lo[bestAxis]) * scale); if (k >= NB) k = NB - 1; if (k < 0) k = 0;
                return k <= bestSplit;
            });
            mid = (uint32_t)(it - B.prims.begin());
            if (mid == t.lo || mid == t.hi) mid = t.lo + cnt / 2;
        }
        BNode l, r; l.left = l.right = r.left = r.right = -1; l.count = r.count = 0; l.first = r.first = 0;
        l.box.reset(); r.box.reset();
        for (uint32_t i = t.lo; i < mid; i++) l.box.grow(pb[B.prims[i]]);
        for (uint32_t i = mid; i < t.hi; i++) r.box.grow(pb[B.prims[i]]);
        int li = (int)B.nodes.size(); B.nodes.push_back(l);
        int ri = (int)B.nodes.size(); B.nodes.push_back(r);
        B.nodes[t.node].left = li; B.nodes[t.node].right = ri; B.nodes[t.node].count = 0;
        stack.push_back({li, t.lo, mid}); stack.push_back({ri, mid, t.hi});
    }
}

/* conservative slab test: slack covers the rounding of (b-o)*inv on both ends */
inline bool box_hit(const Box& b, vec3 o, vec3 inv, float tmin, float tmax)
{
    float t0 = tmin, t1 = tmax;
    const float oo[3] = {o.x, o.y, o.z}, ii[3] = {inv.x, inv.y, inv.z};
    for (int k = 0; k < 3; k++) {
        float a = (b.lo[k] - oo[k]) * ii[k], c = (b.hi[k] - oo[k]) * ii[k];
        if (a != a || c != c) continue;           /* 0 * inf: ray lies in the slab plane -> no constraint */
        float n = fminf(a, c), f = fmaxf(a, c);
        n = n - 1e-5f * fabsf(n) - 1e-30f; f = f + 1e-5f * fabsf(f) + 1e-30f;
        t0 = fmaxf(t0, n); t1 = fminf(t1, f);
    }
    return t0 <= t1;
}

struct Hit { float t, u, v; int inst, prim; };

/* the intersection contract of oracle.h / DESIGN.md */
inline bool tri_hit(vec3 o, vec3 d, vec3 v0, vec3 v1, vec3 v2, float tmin, float tmax, float& t, float& u, float& v)
{
    vec3 e1 = v1 - v0, e2 = v2 - v0;
    vec3 p = cross(d, e2);
    float det = dot(e1, p);
    if (det == 0.0f) return false;
    float inv = 1.0f / det;
    vec3 s = o - v0;
    u = dot(s, p) * inv;
    if (!(u >= 0.0f) || u > 1.0f) return false;
    vec3 q = cross(s, e1);
    v = dot(d, q) * inv;
    if (!(v >= 0.0f) || u + v > 1.0f) return false;
    t = dot(e2, q) * inv;
    return t > tmin && t < tmax;
}

inline bool better(float t, int inst, int prim, const Hit& h)
{
    if (h.inst < 0) return true;
    if (t < h.t) return true;
    if (t > h.t) return false;
    if (inst != h.inst) return inst < h.inst;
    return prim < h.prim;
}

bool trace(const orc_scene& S, vec3 o, vec3 d, float tmin, float tmax, bool any_hit, bool brute, Hit& best)
{
    best.t = -1.f; best.u = best.v = 0.f; best.inst = -1; best.prim = -1;
    vec3 winv(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    for (int ii = 0; ii < (int)S.instances.size(); ii++) {
        const OInst& I = S.instances[ii];
        const float limit = best.inst >= 0 ? best.t : tmax;
        if (!brute && !box_hit(I.world, o, winv, tmin, limit)) continue;
        const OBlas& B = S.blas[I.blas];
        const OMesh& M = S.meshes[B.mesh];
        vec3 oo = mul_point(I.rec.inv_model, o);
        vec3 od = mul_dir(I.rec.inv_model, d);
        if (brute) {
            const uint32_t n = B.index_count / 3u;
            for (uint32_t p = 0; p < n; p++) {
                vec3 a, b, c; tri_vertices(M, B.first_index, p, a, b, c);
                float t, u, v;
                if (tri_hit(oo, od, a, b, c, tmin, tmax, t, u, v)) {
                    if (any_hit) { best.t = t; best.u = u; best.v = v; best.inst = ii; best.prim = (int)p; return true; }
                    if (better(t, ii, (int)p, best)) { best.t = t; best.u = u; best.v = v; best.inst = ii; best.prim = (int)p; }
                }
            }
            continue;
        }
        vec3 inv(1.0f / od.x, 1.0f / od.y, 1.0f / od.z);
        int stack[128]; int sp = 0; stack[sp++] = 0;
        while (sp) {
            const BNode& N = B.nodes[stack[--sp]];
            const float lim = best.inst >= 0 ? best.t : tmax;
            if (!box_hit(N.box, oo, inv, tmin, lim)) continue;
            if (N.left < 0) {
                for (uint32_t k = 0; k < N.count; k++) {
                    uint32_t p = B.prims[N.first + k];
                    vec3 a, b, c; tri_vertices(M, B.first_index, p, a, b, c);
                    float t, u, v;
                    if (tri_hit(oo, od, a, b, c, tmin, tmax, t, u, v)) {
                        if (any_hit) { best.t = t; best.u = u; best.v = v; best.inst = ii; best.prim = (int)p; return true; }
                        if (better(t, ii, (int)p, best)) { best.t = t; best.u = u; best.v = v; best.inst = ii; best.prim = (int)p; }
                    }
                }
            } else {
                if (sp + 2 > 128) { fprintf(stderr, "oracle: BVH stack overflow\n"); abort(); }
                stack[sp++] = N.left; stack[sp++] = N.right;
            }
        }
    }
    return best.inst >= 0;
}

void update_instance_boxes(orc_scene& S)
{
    for (OInst& I : S.instances) {
        const Box& r = S.blas[I.blas].nodes[0].box;
        I.world.reset();
        for (int c = 0; c < 8; c++) {
            vec3 p((c & 1) ? r.hi[0] : r.lo[0], (c & 2) ? r.hi[1] : r.lo[1], (c & 4) ? r.hi[2] : r.lo[2]);
            vec3 w = mul_point(I.rec.model, p);
            I.world.grow(&w.x);
        }
        /* pad: the object-space hit is exact inside the object box, the world box only approximately */
        for (int k = 0; k < 3; k++) {
            float e = 1e-4f * fmaxf(fmaxf(fabsf(I.world.lo[k]), fabsf(I.world.hi[k])), 1e-3f);
            I.world.lo[k] -= e; I.world.hi[k] += e;
        }
    }
}

void set_instances(orc_scene& S, const vkp_instance* inst, int n)
{
    S.instances.clear();
    for (int i = 0; i < n; i++) {
        OInst I; I.rec = inst[i]; I.blas = -1;
        for (size_t b = 0; b < S.blas.size(); b++)
            if (S.blas[b].mesh == inst[i].mesh_id && S.blas[b].first_index == inst[i].first_index && S.blas[b].index_count == inst[i].index_count) { I.blas = (int)b; break; }
        if (I.blas < 0) {
            OBlas B; B.mesh = inst[i].mesh_id; B.first_index = inst[i].first_index; B.index_count = inst[i].index_count;
            build_blas(S.meshes[B.mesh], B);
            S.blas.push_back(std::move(B)); I.blas = (int)S.blas.size() - 1;
        }
        S.instances.push_back(I);
    }
    update_instance_boxes(S);
}

/* ---- texture / G-buffer fetch helpers ---------------------------------------------------------- */

/* texture() with VK_FILTER_NEAREST + REPEAT, 8-bit RGBA (vk_textures.cpp:15,26; vk_initializers.h:86) */
inline vec4 tex_fetch(const OTex& T, float u, float v)
{
    float fx = u * (float)T.w, fy = v * (float)T.h;
    int ix = (fx == fx) ? (int)floorf(fx) : 0, iy = (fy == fy) ? (int)floorf(fy) : 0;
    ix %= T.w; if (ix < 0) ix += T.w;
    iy %= T.h; if (iy < 0) iy += T.h;
    const uint8_t* p = &T.rgba[4 * ((size_t)iy * T.w + ix)];
    vec4 r; r.x = from_unorm8(p[0]); r.y = from_unorm8(p[1]); r.z = from_unorm8(p[2]); r.w = from_unorm8(p[3]);
    return r;
}
inline vec3 xyz(vec4 v) { return vec3(v.x, v.y, v.z); }

/* miss.rmiss:16 / hybridRaygen.rgen:59 equirect mapping; pi differs between the two files */
inline vec2 env_uv(vec3 d, float pi)
{
    vec2 uv;
    uv.x = 0.5f + e_atan2(d.x, d.z) / (2.0f * pi);
    uv.y = 0.5f - e_asin(d.y) / pi;
    return uv;
}

inline const vkp_material& material_at(const orc_scene& S, float matIdx)
{
    int i = (int)matIdx; if (i < 0) i = 0; if (i >= (int)S.materials.size()) i = (int)S.materials.size() - 1;
    return S.materials[i];
}

inline vec4 load_half4(const uint16_t* plane, size_t px)
{
    vec4 r; r.x = from_half(plane[4 * px + 0]); r.y = from_half(plane[4 * px + 1]); r.z = from_half(plane[4 * px + 2]); r.w = from_half(plane[4 * px + 3]);
    return r;
}
inline vec3 load_u8x3(const uint8_t* plane, size_t px)
{
    return vec3(from_unorm8(plane[4 * px + 0]), from_unorm8(plane[4 * px + 1]), from_unorm8(plane[4 * px + 2]));
}

/* ---- helpers.glsl ------------------------------------------------------------------------------ */
const float PI = 3.14159265359f;        /* helpers.glsl:5 */
const float PI_MISS = 3.141592f;        /* miss.rmiss:11  */

/* helpers.glsl:33-44 */
float DistributionGGX(vec3 N, vec3 H, float roughness)
{
    float a = roughness * roughness;
    float a2 = a * a;
    float NdotH = maxf(dot(N, H), 0.0f);
    float NdotH2 = NdotH * NdotH;
    float denom = (NdotH2 * (a2 - 1.0f) + 1.0f);
    denom = PI * denom * denom;
    return a2 / denom;
}
/* helpers.glsl:47-56 */
float GeometrySchlickGGX(float NdotV, float roughness)
{
    float r = roughness + 1.0f;
    float k = (r * r) / 8.0f;
    float nom = NdotV;
    float denom = NdotV * (1.0f - k) + k;
    return nom / denom;
}
/* helpers.glsl:58-66 */
float GeometrySmith(vec3 N, vec3 V, vec3 L, float roughness)
{
    float NdotV = maxf(dot(N, V), 0.0f);
    float NdotL = maxf(dot(N, L), 0.0f);
    float ggx1 = GeometrySchlickGGX(NdotV, roughness);
    float ggx2 = GeometrySchlickGGX(NdotL, roughness);
    return ggx1 * ggx2;
}
/* helpers.glsl:69-72 */
vec3 FresnelSchlick(float cosTheta, vec3 F0)
{
    return F0 + (1.0f - F0) * e_pow(1.0f - cosTheta, 5.0f);
}
/* helpers.glsl:82-93.  `seed` is taken BY VALUE in the reference (App. A #1); `advance` selects the
 * non-faithful variant that writes the advanced state back. */
vec3 sampleDisk(const vkp_light& light, vec3 position, vec3 L, uint32_t& seed_io, bool advance)
{
    uint32_t seed = seed_io;
    float radius = light.radius * sqrtf(rnd(seed));
    float angle = rnd(seed) * 2.0f * PI;
    vec2 point; point.x = radius * e_cos(angle); point.y = radius * e_sin(angle);
    vec3 tangent = normalize(cross(L, vec3(0, 1, 0)));
    if (L == vec3(0, 1, 0)) tangent = vec3(0, 0, 1);
    vec3 bitangent = normalize(cross(tangent, L));
    vec3 target = position + L + point.x * tangent + point.y * bitangent;
    if (advance) seed_io = seed;
    return normalize(target - position);
}

/* camera ray of raygen.rgen:32-40 for pixel (x,y) with sub-pixel offset (ox,oy) */
inline void camera_ray(const vkp_camera& cam, int x, int y, float ox, float oy, int W, int H, vec3& origin, vec3& dir)
{
    float pcx = (float)x + ox, pcy = (float)y + oy;
    float u = pcx / (float)W, v = pcy / (float)H;
    float dx = u * 2.0f - 1.0f, dy = v * 2.0f - 1.0f;
    origin = vec3(cam.inv_view[12], cam.inv_view[13], cam.inv_view[14]);       /* viewInverse * (0,0,0,1) */
    vec4 tv; tv.x = dx; tv.y = dy; tv.z = 1.0f; tv.w = 1.0f;
    vec4 target = mul_vec4(cam.inv_proj, tv);
    vec3 tn = normalize(xyz(target));
    dir = mul_dir(cam.inv_view, tn);
}

struct Payload { vec3 color; float dist; vec3 direction; float scatter; vec3 origin; uint32_t seed; };

/* miss.rmiss:13-19 */
void shade_miss(const orc_scene& S, vec3 rayDir, Payload& prd)
{
    vec3 dir = normalize(rayDir);
    vec2 uv = env_uv(dir, PI_MISS);
    vec3 color = pow3(xyz(tex_fetch(S.textures[S.sky_tex], uv.x, uv.y)), 2.2f);
    prd.color = color; prd.dist = -1.0f;
}

/* hybridHit.rchit:24-184 (rt_mode=false) and closesthit.rchit:26-186 (rt_mode=true) */
void shade_hit(const orc_scene& S, const vkp_params& P, bool rt_mode, vec3 rayO, vec3 rayD, const Hit& h, Payload& prd, orc_counters& C)
{
    const OInst& I = S.instances[h.inst];
    const OBlas& B = S.blas[I.blas];
    const OMesh& M = S.meshes[B.mesh];
    const vec3 bary(1.0f - h.u - h.v, h.u, h.v);
    const uint32_t* ind = &M.idx[3u * (uint32_t)h.prim + I.rec.first_index];
    const vkp_vertex& v0 = M.v[ind[0]]; const vkp_vertex& v1 = M.v[ind[1]]; const vkp_vertex& v2 = M.v[ind[2]];
    const vec3 n0(v0.normal[0], v0.normal[1], v0.normal[2]), n1(v1.normal[0], v1.normal[1], v1.normal[2]), n2(v2.normal[0], v2.normal[1], v2.normal[2]);
    const vec3 normal = n0 * bary.x + n1 * bary.y + n2 * bary.z;
    vec2 uv; uv.x = v0.uv[0] * bary.x + v1.uv[0] * bary.y + v2.uv[0] * bary.z;
             uv.y = v0.uv[1] * bary.x + v1.uv[1] * bary.y + v2.uv[1] * bary.z;
    /* N = normalize(mat3(transpose(inverse(model))) * normal): the host supplies inverse(model) */
    const vec3 N = normalize(mul_transpose3(I.rec.inv_model, normal));
    const vec3 V = normalize(-rayD);
    const float NdotV = clampf(dot(N, V), 0.0f, 1.0f);
    const vec3 worldPos = rayO + rayD * h.t;

    vec3 Lo(0.0f);
    float attenuation = 1.0f;

    const vkp_material& mat = S.materials[std::min<int>(std::max<int>(I.rec.material_id, 0), (int)S.materials.size() - 1)];
    const int shadingMode = (int)mat.shadingMetallicRoughness[0];
    vec3 albedo = mat.textures[0] > -1 ? xyz(tex_fetch(S.textures[(int)mat.textures[0]], uv.x, uv.y)) : vec3(mat.diffuse[0], mat.diffuse[1], mat.diffuse[2]);
    const vec3 emissive = mat.textures[2] > -1 ? xyz(tex_fetch(S.textures[(int)mat.textures[2]], uv.x, uv.y)) : vec3(0.0f);
    const vec3 roughnessMetallic = mat.textures[3] > -1 ? xyz(tex_fetch(S.textures[(int)mat.textures[3]], uv.x, uv.y))
                                                        : vec3(0.0f, mat.shadingMetallicRoughness[2], mat.shadingMetallicRoughness[1]);
    albedo = pow3(albedo, 2.2f);
    const float roughness = roughnessMetallic.y;
    const float metallic = roughnessMetallic.z;
    const vec3 F0 = mix(vec3(0.04f), albedo, metallic);

    const vec2 environmentUV = env_uv(N, PI);
    const vec3 irradiance = xyz(tex_fetch(S.textures[S.irr_tex], environmentUV.x, environmentUV.y));

    vec3 direction(1, 1, 1); float scatter = 0.0f;
    const vec3 origin = worldPos;
    const int shadowSamples = P.shadow_samples;
    const bool advance = !P.faithful_rng;

    for (size_t i = 0; i < S.lights.size(); i++) {
        const vkp_light& light = S.lights[i];
        const vec3 lpos(light.position[0], light.position[1], light.position[2]);
        const vec3 lcol(light.color[0], light.color[1], light.color[2]);
        const bool isDirectional = light.position[3] < 0;
        vec3 L = isDirectional ? lpos : (lpos - worldPos);
        const float light_max_distance = light.position[3];
        const float light_distance = length(L);
        L = normalize(L);
        const float light_intensity = isDirectional ? 1.0f : (light.color[3] / (light_distance * light_distance));
        const vec3 H = normalize(V + L);
        const float NdotL = clampf(dot(N, L), 0.0f, 1.0f);
        const float NdotH = clampf(dot(N, H), 0.0f, 1.0f);
        float shadowFactor = rt_mode ? 0.0f : 1.0f;            /* closesthit.rchit:96 vs hybridHit.rchit:92 */

        if (rt_mode) {
            /* closesthit.rchit:99-120 */
            if (NdotL > 0.0f) {
                for (int a = 0; a < shadowSamples; a++) {
                    bool shadowed = true;
                    if (light_distance < light_max_distance) {
                        vec3 dir = sampleDisk(light, worldPos, L, prd.seed, advance);
                        Hit sh; C.rays[3]++;
                        shadowed = trace(S, worldPos + dir * 1e-2f, dir, 0.001f, light_distance + 1, true, false, sh);
                    }
                    if (!shadowed) shadowFactor++;
                }
                shadowFactor /= shadowSamples;
            }
            /* closesthit.rchit:123-131 */
            if (light_intensity == 0) attenuation = 0.0f;
            else {
                attenuation = light_max_distance - light_distance;
                attenuation /= light_max_distance;
                attenuation = maxf(attenuation, 0.0f);
                attenuation = attenuation * attenuation;
            }
        } else {
            /* hybridHit.rchit:96-104 */
            if (light_intensity == 0) attenuation = 0.0f;
            else {
                attenuation = light_max_distance - light_distance;
                attenuation /= light_max_distance;
                attenuation = maxf(attenuation, 0.0f);
                attenuation = isDirectional ? 0.3f : attenuation * attenuation;
            }
        }

        if (shadingMode == 0) {
            if (!rt_mode) {
                /* hybridHit.rchit:110-131: one secondary shadow ray per light, factor 1 -> 2 */
                if (NdotL > 0) {
                    bool shadowed = true;
                    if (light_distance < light_max_distance) {
                        vec3 dir = sampleDisk(light, worldPos, L, prd.seed, advance);
                        Hit sh; C.rays[3]++;
                        shadowed = trace(S, worldPos + dir * 1e-2f, dir, 0.001f, light_distance + 1, true, false, sh);
                    }
                    if (!shadowed) shadowFactor++;
                }
            }
            vec3 radiance = light_intensity * lcol * attenuation * shadowFactor;
            vec3 F = FresnelSchlick(NdotH, F0);
            float D = DistributionGGX(N, H, roughness);
            float G = GeometrySmith(N, V, L, roughness);
            vec3 numerator = D * G * F;
            float denominator = maxf(4.0f * clampf(dot(N, V), 0.0f, 1.0f) * NdotL, 0.000001f);
            vec3 specular = numerator / denominator;
            vec3 kS = F;
            vec3 kD = (vec3(1.0f) - kS) * (1.0f - metallic);
            Lo += (kD * albedo / PI + specular) * radiance * NdotL;
            direction = vec3(1, 1, 1); scatter = 0.0f;
        } else if (shadingMode == 3) {
            const vec3 reflected = reflect(normalize(rayD), N);
            const bool isScattered = dot(reflected, N) > 0;
            if (rt_mode)   /* closesthit.rchit:155-157 */
                Lo += (NdotL > 0.0f && light_intensity > 0.0f) ? light_intensity * lcol * attenuation * shadowFactor * metallic * albedo
                                                               : irradiance * albedo * metallic;
            else           /* hybridHit.rchit:153-155 */
                Lo += (NdotL > 0.0f && light_intensity > 0.0f) ? light_intensity * lcol * attenuation * albedo * metallic
                                                               : irradiance * albedo * metallic;
            direction = reflected; scatter = isScattered ? 1.0f : 0.0f;
        } else if (shadingMode == 4) {
            const float ior = mat.diffuse[3];
            const float NdotVg = dot(N, normalize(rayD));
            const vec3 refrNormal = NdotVg > 0.0f ? -N : N;
            const float refrEta = NdotVg > 0.0f ? 1 / ior : ior;
            Lo += (light_intensity > 0.0f) ? light_intensity * lcol * attenuation * albedo * metallic : irradiance * albedo * metallic;
            float radicand = 1 + e_pow(refrEta, 2.0f) * (NdotVg * NdotVg - 1);
            if (radicand < 0.0f) direction = reflect(rayD, N);
            else direction = rt_mode ? refract(rayD, refrNormal, refrEta)                 /* closesthit.rchit:174 (un-normalised) */
                                     : refract(normalize(rayD), refrNormal, refrEta);     /* hybridHit.rchit:172 */
            scatter = 1.0f;
        }
    }

    /* ambient from IBL: hybridHit.rchit:177-182 / closesthit.rchit:179-184 */
    vec3 F = FresnelSchlick(NdotV, F0);
    vec3 kD = (1.0f - F) * (1.0f - metallic);
    vec3 diffuse = kD * albedo * irradiance;
    vec3 color = Lo + diffuse + emissive;
    prd.color = color; prd.dist = h.t; prd.direction = direction; prd.scatter = scatter; prd.origin = origin;
}

} // namespace

/* ------------------------------------------------------------------------------------------------ */
extern "C" {

int orc_abi_version(void) { return 1; }
void orc_set_row_sampling(int stride, int offset) { g_row_stride = stride > 0 ? stride : 1; g_row_offset = offset; }
int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int orc_uses_libm(void)
{
#ifdef ORC_USE_LIBM
    return 1;
#else
    return 0;
#endif
}

orc_scene* orc_scene_create(const orc_scene_desc* d)
{
    orc_scene* S = new orc_scene();
    for (int i = 0; i < d->n_meshes; i++) {
        OMesh m; m.v.assign(d->meshes[i].vertices, d->meshes[i].vertices + d->meshes[i].n_vertices);
        m.idx.assign(d->meshes[i].indices, d->meshes[i].indices + d->meshes[i].n_indices);
        S->meshes.push_back(std::move(m));
    }
    for (int i = 0; i < d->n_textures; i++) {
        OTex t; t.w = d->textures[i].width; t.h = d->textures[i].height;
        t.rgba.assign(d->textures[i].rgba, d->textures[i].rgba + 4 * (size_t)t.w * t.h);
        S->textures.push_back(std::move(t));
    }
    S->sky_tex = d->sky_tex; S->irr_tex = d->irradiance_tex;
    S->materials.assign(d->materials, d->materials + d->n_materials);
    S->lights.assign(d->lights, d->lights + d->n_lights);
    set_instances(*S, d->instances, d->n_instances);
    return S;
}
void orc_scene_destroy(orc_scene* s) { delete s; }
void orc_scene_set_lights(orc_scene* s, const vkp_light* l, int n) { s->lights.assign(l, l + n); }
void orc_scene_set_instances(orc_scene* s, const vkp_instance* inst, int n) { set_instances(*s, inst, n); }
uint64_t orc_scene_triangles(const orc_scene* s)
{
    uint64_t t = 0; for (const OInst& I : s->instances) t += I.rec.index_count / 3u; return t;
}

void orc_trace_rays(const orc_scene* s, const vkp_ray* rays, uint32_t n, int any_hit, int brute_force, vkp_hit* hits)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        Hit h;
        vec3 o(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]), d(rays[i].dir[0], rays[i].dir[1], rays[i].dir[2]);
        bool hit = trace(*s, o, d, rays[i].tmin, rays[i].tmax, any_hit != 0, brute_force != 0, h);
        hits[i].t = hit ? h.t : -1.0f; hits[i].u = h.u; hits[i].v = h.v; hits[i].instance = h.inst; hits[i].primitive = h.prim;
    }
}

/* G-buffer content of basic.vert:36-48 + geometry_shader.frag:51-77 + skybox.frag:14-22, produced by
 * ray casting the camera ray of raygen.rgen:32-49 instead of rasterising (declared deviation: the
 * raster top-left rule / depth test are out of contract; SURVEY section 7 "hard parts"). */
void orc_gbuffer_pass(const orc_scene* sp, const vkp_camera* cam, const vkp_params* p, orc_gbuffer* g, orc_counters* C)
{
    const orc_scene& S = *sp;
    const int W = p->width, H = p->height;
    uint64_t nrays = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : nrays)
    for (int y = 0; y < H; y++) {
        if (g_row_stride > 1 && y % g_row_stride != g_row_offset) continue;
        for (int x = 0; x < W; x++) {
            const size_t px = (size_t)y * W + x;
            vec3 o, d; camera_ray(*cam, x, y, 0.5f, 0.5f, W, H, o, d);
            Hit h; nrays++;
            bool hit = trace(S, o + d * 1e-2f, d, 0.001f, 10000.0f, false, false, h);
            if (g->hit_ids) { g->hit_ids[2 * px] = h.inst; g->hit_ids[2 * px + 1] = h.prim; }
            if (!hit) {
                /* skybox.frag:16-21: everything zero, albedo = sky texel */
                for (int k = 0; k < 4; k++) { g->position[4 * px + k] = 0; g->normal[4 * px + k] = 0; g->material[4 * px + k] = 0; g->emissive[4 * px + k] = 0; }
                g->motion[2 * px] = 0; g->motion[2 * px + 1] = 0;
                vec2 uv = env_uv(normalize(d), PI_MISS);
                vec4 sky = tex_fetch(S.textures[S.sky_tex], uv.x, uv.y);
                g->albedo[4 * px + 0] = to_unorm8(sky.x); g->albedo[4 * px + 1] = to_unorm8(sky.y);
                g->albedo[4 * px + 2] = to_unorm8(sky.z); g->albedo[4 * px + 3] = to_unorm8(sky.w);
                continue;
            }
            const OInst& I = S.instances[h.inst];
            const OBlas& B = S.blas[I.blas];
            const OMesh& M = S.meshes[B.mesh];
            const uint32_t* ind = &M.idx[3u * (uint32_t)h.prim + I.rec.first_index];
            const vkp_vertex& v0 = M.v[ind[0]]; const vkp_vertex& v1 = M.v[ind[1]]; const vkp_vertex& v2 = M.v[ind[2]];
            const float b0 = 1.0f - h.u - h.v, b1 = h.u, b2 = h.v;
            /* basic.vert:42 outPosition = model * pos, interpolated */
            vec3 P0 = mul_point(I.rec.model, vec3(v0.position[0], v0.position[1], v0.position[2]));
            vec3 P1 = mul_point(I.rec.model, vec3(v1.position[0], v1.position[1], v1.position[2]));
            vec3 P2 = mul_point(I.rec.model, vec3(v2.position[0], v2.position[1], v2.position[2]));
            vec3 worldPos = P0 * b0 + P1 * b1 + P2 * b2;
            /* basic.vert:44 outNormal = mat3(transpose(inv_matrix)) * normal */
            vec3 n = vec3(v0.normal[0], v0.normal[1], v0.normal[2]) * b0 + vec3(v1.normal[0], v1.normal[1], v1.normal[2]) * b1 + vec3(v2.normal[0], v2.normal[1], v2.normal[2]) * b2;
            vec3 inNormal = mul_transpose3(I.rec.inv_model, n);
            vec3 inColor = vec3(v0.color[0], v0.color[1], v0.color[2]) * b0 + vec3(v1.color[0], v1.color[1], v1.color[2]) * b1 + vec3(v2.color[0], v2.color[1], v2.color[2]) * b2;
            vec2 uv; uv.x = v0.uv[0] * b0 + v1.uv[0] * b1 + v2.uv[0] * b2; uv.y = v0.uv[1] * b0 + v1.uv[1] * b1 + v2.uv[1] * b2;
            const vkp_material& mat = S.materials[std::min<int>(std::max<int>(I.rec.material_id, 0), (int)S.materials.size() - 1)];
            /* geometry_shader.frag:54-59 */
            vec3 color = mat.textures[0] > -1 ? xyz(tex_fetch(S.textures[(int)mat.textures[0]], uv.x, uv.y)) * inColor : vec3(mat.diffuse[0], mat.diffuse[1], mat.diffuse[2]);
            vec3 N = normalize(inNormal);
            if (mat.textures[1] > -1) {
                /* geometry_shader.frag:55,60-63 perturbNormal(inNormal, inWorldPos, inUV, texel) with cotangent_frame :28-49.
                 * dFdx / dFdy of the world position are ray differentials here (declared deviation, no rasteriser quads):
                 * the camera rays of the right / lower neighbour pixel intersected with the plane of the hit triangle.
                 * As upstream: the frame is built from the UN-normalised inNormal, and duv1 / duv2 are dFdx(p).xy /
                 * dFdy(p).xy (:32-33 take them from p, not from uv). */
                vec3 normal_pixel = xyz(tex_fetch(S.textures[(int)mat.textures[1]], uv.x, uv.y));
                normal_pixel = normal_pixel * 255.0f / 127.0f - vec3(128.0f / 127.0f);
                const vec3 ng = cross(P1 - P0, P2 - P0);
                vec3 ox_, dx_, oy_, dy_;
                camera_ray(*cam, x + 1, y, 0.5f, 0.5f, W, H, ox_, dx_);
                camera_ray(*cam, x, y + 1, 0.5f, 0.5f, W, H, oy_, dy_);
                const vec3 px_ = ox_ + dx_ * (dot(P0 - ox_, ng) / dot(dx_, ng));
                const vec3 py_ = oy_ + dy_ * (dot(P0 - oy_, ng) / dot(dy_, ng));
                const vec3 dp1 = px_ - worldPos, dp2 = py_ - worldPos;
                const float duv1x = dp1.x, duv1y = dp1.y, duv2x = dp2.x, duv2y = dp2.y;
                const vec3 dp2perp = cross(dp2, inNormal), dp1perp = cross(inNormal, dp1);
                const vec3 T = dp2perp * duv1x + dp1perp * duv2x;
                const vec3 Bt = dp2perp * duv1y + dp1perp * duv2y;
                const float invmax = 1.0f / sqrtf(fmaxf(dot(T, T), dot(Bt, Bt)));
                const vec3 Tn = T * invmax, Bn = Bt * invmax;
                /* mat3(T, B, N) * v = T * v.x + B * v.y + N * v.z */
                N = normalize(Tn * normal_pixel.x + Bn * normal_pixel.y + inNormal * normal_pixel.z);
            }
            vec3 emissive = mat.textures[2] > -1 ? xyz(tex_fetch(S.textures[(int)mat.textures[2]], uv.x, uv.y)) : vec3(0.0f);
            vec3 material = mat.textures[3] > -1 ? xyz(tex_fetch(S.textures[(int)mat.textures[3]], uv.x, uv.y)) : vec3(0.0f, mat.shadingMetallicRoughness[2], mat.shadingMetallicRoughness[1]);
            float materialIdx = mat.shadingMetallicRoughness[3];
            /* geometry_shader.frag:66-70 into RGBA16F / RGBA8 attachments (renderer.cpp:255-260) */
            g->position[4 * px + 0] = to_half(worldPos.x); g->position[4 * px + 1] = to_half(worldPos.y);
            g->position[4 * px + 2] = to_half(worldPos.z); g->position[4 * px + 3] = to_half(materialIdx);
            vec3 nn = N * 0.5f + vec3(0.5f);
            g->normal[4 * px + 0] = to_half(nn.x); g->normal[4 * px + 1] = to_half(nn.y); g->normal[4 * px + 2] = to_half(nn.z); g->normal[4 * px + 3] = to_half(1.0f);
            g->albedo[4 * px + 0] = to_unorm8(color.x); g->albedo[4 * px + 1] = to_unorm8(color.y); g->albedo[4 * px + 2] = to_unorm8(color.z); g->albedo[4 * px + 3] = 255;
            g->material[4 * px + 0] = to_unorm8(material.x); g->material[4 * px + 1] = to_unorm8(material.y); g->material[4 * px + 2] = to_unorm8(material.z); g->material[4 * px + 3] = 255;
            g->emissive[4 * px + 0] = to_unorm8(emissive.x); g->emissive[4 * px + 1] = to_unorm8(emissive.y); g->emissive[4 * px + 2] = to_unorm8(emissive.z); g->emissive[4 * px + 3] = 255;
            /* geometry_shader.frag:73-76 motion = ndc.xy - ndcPrev.xy, stored *0.5+0.5 in RG16F -- where the fragment shader's
             * `inNdc` (location 4) is what basic.vert:47 wrote as `ndcPrev` (location 4) and `inNdcPrev` (location 5) is the
             * vertex shader's `ndc` (location 5, basic.vert:46): the names are crossed over, the stored vector is
             * previous - current.  Found by executing basic.vert.spv + geometry_shader.frag.spv (make_spirv_vectors.py). */
            vec4 wp; wp.x = worldPos.x; wp.y = worldPos.y; wp.z = worldPos.z; wp.w = 1.0f;
            vec4 clip = mul_vec4(cam->proj, mul_vec4(cam->view, wp));
            vec4 clipPrev = mul_vec4(cam->prev_proj, mul_vec4(cam->prev_view, wp));
            float mx = clipPrev.x / clipPrev.w - clip.x / clip.w;
            float my = clipPrev.y / clipPrev.w - clip.y / clip.w;
            g->motion[2 * px + 0] = to_half(mx * 0.5f + 0.5f); g->motion[2 * px + 1] = to_half(my * 0.5f + 0.5f);
        }
    }
    if (C) C->rays[0] += nrays;
}

/* shadowRaygen.rgen:28-117 */
void orc_shadow_pass(const orc_scene* sp, const vkp_camera* cam, const vkp_params* p, const orc_gbuffer* g, uint8_t* shadow, orc_counters* C)
{
    (void)cam;
    const orc_scene& S = *sp;
    const int W = p->width, H = p->height;
    const size_t plane = (size_t)W * H;
    const int nl = (int)S.lights.size();
    uint64_t nrays = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : nrays)
    for (int y = 0; y < H; y++) {
        if (g_row_stride > 1 && y % g_row_stride != g_row_offset) continue;
        for (int x = 0; x < W; x++) {
            const size_t px = (size_t)y * W + x;
            const int frame = p->frame;
            uint32_t seed = tea((uint32_t)y * (uint32_t)W + (uint32_t)x, (uint32_t)frame);
            vec4 pos4 = load_half4(g->position, px);
            vec4 nrm4 = load_half4(g->normal, px);
            vec3 position = xyz(pos4);
            vec3 normal = xyz(nrm4) * 2.0f - vec3(1.0f);
            float matIdx = pos4.w;
            int mode = (int)material_at(S, matIdx).shadingMetallicRoughness[0];
            vec3 N = normalize(normal);
            if (mode == 4) {
                for (int i = 0; i < nl; i++) shadow[i * plane + px] = to_unorm8(1.0f);
                continue;
            }
            for (int i = 0; i < nl; i++) {
                const vkp_light& light = S.lights[i];
                const vec3 lpos(light.position[0], light.position[1], light.position[2]);
                const bool isDirectional = light.position[3] < 0;
                vec3 L = isDirectional ? lpos : (lpos - position);
                const float light_max_distance = light.position[3];
                const float light_distance = length(L);
                L = normalize(L);
                const float NdotL = clampf(dot(N, L), 0.0f, 1.0f);
                float shadowFactor = 0.0f;
                int shadowSamples = p->shadow_samples;
                if (NdotL > 0) {
                    for (int s = 0; s < shadowSamples; s++) {
                        bool shadowed;
                        if (light_distance < light_max_distance) {
                            const vec3 dir = sampleDisk(light, position, L, seed, !p->faithful_rng);
                            float tmin = 0.001f, tmax = light_distance + 1;
                            Hit h; nrays++;
                            shadowed = trace(S, position + dir * 1e-1f, dir, tmin, tmax, true, false, h);
                        } else {
                            shadowed = false;
                        }
                        if (!shadowed) shadowFactor++;
                    }
                    shadowFactor /= shadowSamples;
                }
                shadow[i * plane + px] = to_unorm8(shadowFactor);     /* R8_UNORM, renderer.cpp:1538 */
            }
        }
    }
    if (C) C->rays[1] += nrays;
}

/* vkp_params.denoiser == 1 (no reference counterpart: SURVEY 8f-3 "a better denoiser").
 * One iteration of an edge-aware a-trous filter over the R8 shadow planes, guided by the G-buffer: 5 x 5 taps `step` pixels apart,
 * B3-spline weights (3/8, 1/4, 1/16 per axis), times a normal weight max(0, N_p . N_q)^32 and a plane-distance weight
 * max(0, 1 - |N_p . (P_q - P_p)| / (0.1 * step)).  Background pixels (normal.w == 0) pass through and are never taps.
 * Taps outside the image are skipped.  out = unorm8(sum w v / sum w); the centre tap has weight 9/64, so the sum is never empty. */
static void atrous_iteration(const vkp_params* p, int n_lights, const orc_gbuffer* g, const uint8_t* in, uint8_t* out, int step)
{
    const int W = p->width, H = p->height;
    const size_t plane = (size_t)W * H;
    static const float k[3] = {0.375f, 0.25f, 0.0625f};
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        if (g_row_stride > 1 && y % g_row_stride != g_row_offset) continue;
        for (int x = 0; x < W; x++) {
            const size_t px = (size_t)y * W + x;
            const vec4 np4 = load_half4(g->normal, px);
            if (np4.w == 0.0f) { for (int l = 0; l < n_lights; l++) out[l * plane + px] = in[l * plane + px]; continue; }
            const vec3 Np = normalize(xyz(np4) * 2.0f - vec3(1.0f));
            const vec3 Pp = xyz(load_half4(g->position, px));
            const float inv_sigma = 1.0f / (0.1f * (float)step);
            float sum[VKP_MAX_LIGHTS], wsum = 0.0f;
            for (int l = 0; l < n_lights; l++) sum[l] = 0.0f;
            for (int dy = -2; dy <= 2; dy++)
                for (int dx = -2; dx <= 2; dx++) {
                    const int qx = x + dx * step, qy = y + dy * step;
                    if (qx < 0 || qy < 0 || qx >= W || qy >= H) continue;
                    const size_t q = (size_t)qy * W + qx;
                    const vec4 nq4 = load_half4(g->normal, q);
                    if (nq4.w == 0.0f) continue;
                    const vec3 Nq = normalize(xyz(nq4) * 2.0f - vec3(1.0f));
                    const vec3 Pq = xyz(load_half4(g->position, q));
                    float wn = maxf(dot(Np, Nq), 0.0f);
                    wn = wn * wn; wn = wn * wn; wn = wn * wn; wn = wn * wn; wn = wn * wn;          /* ^32 */
                    const float dist = fabsf(dot(Np, Pq - Pp));
                    const float wp = maxf(1.0f - dist * inv_sigma, 0.0f);
                    const float w = (k[dx < 0 ? -dx : dx] * k[dy < 0 ? -dy : dy]) * wn * wp;
                    wsum = wsum + w;
                    for (int l = 0; l < n_lights; l++) sum[l] = sum[l] + w * from_unorm8(in[l * plane + q]);
                }
            for (int l = 0; l < n_lights; l++) out[l * plane + px] = to_unorm8(sum[l] / wsum);
        }
    }
}

/* blur.comp:35-82 with runtime W/H/NLIGHTS and integer motion fetch; R8 storage */
void orc_denoise_pass(const vkp_params* p, int n_lights, const orc_gbuffer* g, const uint8_t* shadow, uint8_t* denoised)
{
    const int W = p->width, H = p->height;
    const size_t plane = (size_t)W * H;
    const int frame = p->denoise_frame;
    /* renderer.cpp:2672 dispatches W/16 x H/16 groups: remainder pixels are never written */
    const int Wc = p->denoise_cover_remainder ? W : (W / 16) * 16;
    const int Hc = p->denoise_cover_remainder ? H : (H / 16) * 16;
    std::vector<uint8_t> prev(denoised, denoised + plane * n_lights);   /* reads see last frame's image */
    std::vector<uint8_t> filtered, tmp;
    if (p->denoiser == 1) {
        /* three a-trous iterations (steps 1, 2, 4: a 29 x 29 footprint) replace the raw shadow planes as the input of blur.comp's
         * temporal stage below.  Row sampling (bench timing aid) does not apply: every iteration needs its neighbours' rows. */
        const int keep_stride = g_row_stride; g_row_stride = 1;
        filtered.assign(plane * n_lights, 0); tmp.assign(plane * n_lights, 0);
        atrous_iteration(p, n_lights, g, shadow, filtered.data(), 1);
        atrous_iteration(p, n_lights, g, filtered.data(), tmp.data(), 2);
        atrous_iteration(p, n_lights, g, tmp.data(), filtered.data(), 4);
        shadow = filtered.data();
        g_row_stride = keep_stride;
    }
    auto in_at = [&](int l, int x, int y) -> float {
        if (x < 0 || y < 0 || x >= W || y >= H) return 0.0f;            /* out-of-bounds imageLoad -> 0 */
        return from_unorm8(shadow[l * plane + (size_t)y * W + x]);
    };
    auto old_at = [&](int l, int x, int y) -> float {
        if (x < 0 || y < 0 || x >= W || y >= H) return 0.0f;
        return from_unorm8(prev[l * plane + (size_t)y * W + x]);
    };
#pragma omp parallel for schedule(static)
    for (int y = 0; y < Hc; y++) {
        if (g_row_stride > 1 && y % g_row_stride != g_row_offset) continue;
        for (int x = 0; x < Wc; x++) {
            const size_t px = (size_t)y * W + x;
            for (int l = 0; l < n_lights; l++) {
                float pixelColor;
                if (frame > 0) {
                    float center = in_at(l, x, y);
                    float old = old_at(l, x, y);
                    float a = 1.0f / (float)(frame + 1.0f);
                    pixelColor = mixf(old, center, a);
                } else {
                    float rx = from_half(g->motion[2 * px + 0]) * 2.0f - 1.0f;
                    /* blur.comp:53: lastUV = (x + r.x*W, y + r*H) -- the vec2 ctor keeps r.x for the second
                     * component too (App. A #12) */
                    float lastx = (float)x + rx * (float)W, lasty = (float)y + rx * (float)H;
                    if (p->reproject_fix) {
                        /* vkp_params.reproject_fix: motion = previous - current NDC position (see orc_gbuffer_pass), NDC spans
                         * 2 over the screen: the pixel this surface point covered last frame */
                        float ry = from_half(g->motion[2 * px + 1]) * 2.0f - 1.0f;
                        lastx = floorf(((float)x + 0.5f) + (rx * 0.5f) * (float)W);
                        lasty = floorf(((float)y + 0.5f) + (ry * 0.5f) * (float)H);
                    }
                    float center = in_at(l, x, y);
                    float minColor = center, maxColor = center;
                    for (int yy = -1; yy <= 1; yy++)
                        for (int xx = -1; xx <= 1; xx++) {
                            if (xx == 0 && yy == 0) continue;
                            float c = in_at(l, x + xx, y + yy);
                            minColor = minf(minColor, c); maxColor = maxf(maxColor, c);
                        }
                    /* ivec2(lastUV): float -> int truncation; clamp keeps the cast defined */
                    float cx = fminf(fmaxf(lastx, -2147483000.0f), 2147483000.0f), cy = fminf(fmaxf(lasty, -2147483000.0f), 2147483000.0f);
                    if (cx != cx) cx = -1.0f; if (cy != cy) cy = -1.0f;
                    float old = old_at(l, (int)cx, (int)cy);
                    old = maxf(minColor, old);
                    old = minf(maxColor, old);
                    pixelColor = mixf(old, center, 0.4f);
                }
                denoised[l * plane + px] = to_unorm8(pixelColor);
            }
        }
    }
}

/* hybridRaygen.rgen:27-186 (+ hybridHit.rchit, miss.rmiss through shade_hit / shade_miss) */
void orc_resolve_pass(const orc_scene* sp, const vkp_camera* cam, const vkp_params* p, const orc_gbuffer* g,
                      const uint8_t* denoised, float* out, orc_counters* C)
{
    const orc_scene& S = *sp;
    const int W = p->width, H = p->height;
    const size_t plane = (size_t)W * H;
    uint64_t r_refl = 0, r_sec = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : r_refl, r_sec)
    for (int y = 0; y < H; y++) {
        if (g_row_stride > 1 && y % g_row_stride != g_row_offset) continue;
        for (int x = 0; x < W; x++) {
            const size_t px = (size_t)y * W + x;
            orc_counters lc; memset(&lc, 0, sizeof(lc));
            Payload prd;
            prd.seed = tea((uint32_t)y * (uint32_t)W + (uint32_t)x, (uint32_t)p->frame);
            vec4 pos4 = load_half4(g->position, px);
            vec4 nrm4 = load_half4(g->normal, px);
            float matIdx = pos4.w;
            const vkp_material& mat = material_at(S, matIdx);
            int shadingMode = (int)mat.shadingMetallicRoughness[0];
            vec3 position = xyz(pos4);
            vec3 normal = xyz(nrm4) * 2.0f - vec3(1.0f);
            vec3 albedo = pow3(load_u8x3(g->albedo, px), 2.2f);
            vec3 material = load_u8x3(g->material, px);
            vec3 emissive = load_u8x3(g->emissive, px);
            bool background = pos4.w == 0 && nrm4.w == 0;
            const float roughness = material.y;
            const float metallic = material.z;
            const vec3 F0 = mix(vec3(0.04f), albedo, metallic);
            const vec3 camPosition(cam->inv_view[12], cam->inv_view[13], cam->inv_view[14]);
            const vec3 V = normalize(camPosition - position);
            const vec3 N = normalize(normal);
            const float NdotV = clampf(dot(N, V), 0.0f, 1.0f);
            vec2 environmentUV = env_uv(N, PI);
            vec3 irradiance = xyz(tex_fetch(S.textures[S.irr_tex], environmentUV.x, environmentUV.y));

            float tmin = 0.001f, tmax = 1000.0f;
            vec3 finalColor(0.0f), origin(0.0f), direction(0.0f);
            float attenuation = 1.0f;                 /* declared outside the loop: stale across lights (App. A #8) */
            float shadowFactor = 0.0f;
            vec3 rayColor(0.0f);
            for (size_t i = 0; i < S.lights.size(); i++) {
                const vkp_light& light = S.lights[i];
                const vec3 lpos(light.position[0], light.position[1], light.position[2]);
                const vec3 lcol(light.color[0], light.color[1], light.color[2]);
                const bool isDirectional = light.position[3] < 0;
                vec3 L = isDirectional ? lpos : (lpos - position);
                const float light_max_distance = light.position[3];
                const float light_distance = length(L);
                const float light_intensity = isDirectional ? 1.0f : (light.color[3] / (light_distance * light_distance));
                L = normalize(L);
                const float NdotL = clampf(dot(N, L), 0.0f, 1.0f);
                shadowFactor = from_unorm8(denoised[i * plane + px]);
                if (NdotL > 0.0f) {
                    if (light_intensity == 0) attenuation = 0.0f;
                    else {
                        attenuation = light_max_distance - light_distance;
                        attenuation /= light_max_distance;
                        attenuation = maxf(attenuation, 0.0f);
                        attenuation = isDirectional ? 0.3f : (attenuation * attenuation);
                    }
                }
                tmax = 1000.0f;
                if (shadingMode == 0) {
                    vec3 radiance = lcol * light_intensity * attenuation * shadowFactor;
                    const vec3 Hh = normalize(V + L);
                    float D = DistributionGGX(N, Hh, roughness);
                    float G = GeometrySmith(N, V, L, roughness);
                    vec3 F = FresnelSchlick(maxf(dot(Hh, V), 0.0f), F0);
                    vec3 kD = vec3(1.0f) - F;
                    kD *= 1.0f - metallic;
                    vec3 numerator = D * G * F;
                    float denominator = 4.0f * NdotV * NdotL;
                    vec3 specular = numerator / maxf(denominator, 0.001f);
                    rayColor += (kD * albedo / PI + specular) * radiance * NdotL;
                }
                if (shadingMode == 3) {
                    direction = reflect(normalize(-V), N);
                    origin = position;
                    rayColor += (NdotL > 0.0f && light_intensity > 0.0f) ? lcol * light_intensity * attenuation * shadowFactor * albedo * metallic
                                                                         : irradiance * albedo * metallic;
                }
                if (shadingMode == 4) {
                    float ior = mat.diffuse[3];
                    origin = position;
                    const float cosAlpha = dot(N, V);
                    const vec3 Iv = -V;
                    float NdotI = dot(N, Iv);
                    vec3 refrNormal = NdotI > 0.0f ? -N : N;
                    float refrEta = NdotI > 0.0f ? 1 / ior : ior;
                    float radicand = 1 + e_pow(refrEta, 2.0f) * (cosAlpha * cosAlpha - 1);
                    direction = radicand < 0.0f ? reflect(Iv, N) : refract(Iv, refrNormal, refrEta);
                    rayColor += (light_intensity > 0.0f) ? light_intensity * lcol * attenuation * albedo * metallic
                                                         : irradiance * albedo * metallic;
                }
            }
            /* hybridRaygen.rgen:153-170 */
            for (int depth = 0; depth < p->max_bounces; depth++) {
                if (shadingMode == 0) break;
                vec3 ro = origin + direction * 1e-2f;
                Hit h; lc.rays[2]++;
                bool hit = trace(S, ro, direction, tmin, tmax, false, false, h);
                if (hit) shade_hit(S, *p, false, ro, direction, h, prd, lc); else shade_miss(S, direction, prd);
                rayColor *= prd.color;
                const float hitDistance = prd.dist;
                const bool isScattered = hit ? prd.scatter > 0 : false;
                if (hitDistance < 0 || !isScattered) break;
                origin = prd.origin; direction = prd.direction;
            }
            finalColor = rayColor;
            vec3 F = FresnelSchlick(NdotV, F0);
            vec3 kD = (1.0f - F) * (1.0f - metallic);
            vec3 diffuse = kD * albedo * irradiance;
            finalColor += diffuse + emissive;
            if (background) finalColor = albedo;
            out[4 * px + 0] = finalColor.x; out[4 * px + 1] = finalColor.y; out[4 * px + 2] = finalColor.z; out[4 * px + 3] = 1.0f;
            r_refl += lc.rays[2]; r_sec += lc.rays[3];
        }
    }
    if (C) { C->rays[2] += r_refl; C->rays[3] += r_sec; }
}

/* image store to B8G8R8A8_UNORM (renderer.cpp:1535) then postFragment.frag:9-21 into the 8-bit swapchain */
void orc_post_pass(const vkp_params* p, const float* in, uint8_t* out)
{
    const size_t n = (size_t)p->width * p->height;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        for (int k = 0; k < 3; k++) {
            float c = from_unorm8(to_unorm8(in[4 * i + k]));
            c = maxf(c, 0.001f);
            c = e_pow(c, 1.0f / 2.2f);
            out[4 * i + k] = to_unorm8(c);
        }
        out[4 * i + 3] = to_unorm8(in[4 * i + 3]);
    }
}

/* raygen.rgen:20-81 with NSAMPLES = p->spp; accumulation kept in fp32 (declared deviation from the
 * 8-bit image of renderer.cpp:1535) */
void orc_raytrace_pass(const orc_scene* sp, const vkp_camera* cam, const vkp_params* p, float* accum, orc_counters* C)
{
    const orc_scene& S = *sp;
    const int W = p->width, H = p->height;
    uint64_t r_refl = 0, r_sec = 0;
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : r_refl, r_sec)
    for (int y = 0; y < H; y++) {
        if (g_row_stride > 1 && y % g_row_stride != g_row_offset) continue;
        for (int x = 0; x < W; x++) {
            const size_t px = (size_t)y * W + x;
            orc_counters lc; memset(&lc, 0, sizeof(lc));
            const int frame = p->frame;
            Payload prd;
            prd.seed = tea((uint32_t)y * (uint32_t)W + (uint32_t)x, (uint32_t)frame);
            vec3 finalColor(0.0f);
            for (int nS = 0; nS < p->spp; nS++) {
                float r1 = rnd(prd.seed);
                float r2 = rnd(prd.seed);
                float ox = nS == 0 ? 0.5f : r1, oy = nS == 0 ? 0.5f : r2;
                vec3 origin, direction; camera_ray(*cam, x, y, ox, oy, W, H, origin, direction);
                float tmin = 0.001f, tmax = 10000.0f;
                vec3 rayColor(1.0f);
                for (int i = 0; i < p->max_bounces; i++) {
                    vec3 ro = origin + direction * 1e-2f;
                    Hit h; lc.rays[2]++;
                    bool hit = trace(S, ro, direction, tmin, tmax, false, false, h);
                    if (hit) shade_hit(S, *p, true, ro, direction, h, prd, lc); else shade_miss(S, direction, prd);
                    rayColor *= prd.color;
                    const float hitDistance = prd.dist;
                    const bool isScattered = hit ? prd.scatter > 0 : false;
                    if (hitDistance < 0 || !isScattered) break;
                    origin = prd.origin; direction = prd.direction;
                }
                finalColor += rayColor;
            }
            finalColor = finalColor / (float)p->spp;
            if (frame >= 0) {
                float a = 1.0f / (float)(1.0f + frame);
                vec3 last(accum[4 * px], accum[4 * px + 1], accum[4 * px + 2]);
                finalColor = mix(last, finalColor, a);
            }
            accum[4 * px + 0] = finalColor.x; accum[4 * px + 1] = finalColor.y; accum[4 * px + 2] = finalColor.z; accum[4 * px + 3] = 1.0f;
            r_refl += lc.rays[2]; r_sec += lc.rays[3];
        }
    }
    if (C) { C->rays[2] += r_refl; C->rays[3] += r_sec; }
}

/* deferred.frag:30-133 */
void orc_deferred_pass(const orc_scene* sp, const vkp_camera* cam, const vkp_params* p, const orc_gbuffer* g, float* out)
{
    const orc_scene& S = *sp;
    const int W = p->width, H = p->height;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        if (g_row_stride > 1 && y % g_row_stride != g_row_offset) continue;
        for (int x = 0; x < W; x++) {
            const size_t px = (size_t)y * W + x;
            vec4 pos4 = load_half4(g->position, px);
            vec4 nrm4 = load_half4(g->normal, px);
            vec3 position = xyz(pos4);
            vec3 normal = xyz(nrm4) * 2.0f - vec3(1.0f);
            vec3 albedo = load_u8x3(g->albedo, px);
            vec3 motion(from_half(g->motion[2 * px]) * 2.0f - 1.0f, from_half(g->motion[2 * px + 1]) * 2.0f - 1.0f, 0.0f * 2.0f - 1.0f);
            vec3 material = load_u8x3(g->material, px);
            vec3 emissive = load_u8x3(g->emissive, px);
            bool background = pos4.w == 0 && nrm4.w == 0;
            float metallic = material.z, roughness = material.y;
            vec3 N = normalize(normal);
            vec3 V = normalize(vec3(cam->position[0], cam->position[1], cam->position[2]) - position);
            float NdotV = maxf(dot(N, V), 0.0f);
            vec3 F0 = mix(vec3(0.04f), pow3(albedo, 2.2f), metallic);
            vec2 envUV = env_uv(N, PI);
            vec3 irradiance = xyz(tex_fetch(S.textures[S.irr_tex], envUV.x, envUV.y));
            vec3 color(1.0f);
            float alpha = 1.0f;
            if (p->debug_target > 0) {
                switch (p->debug_target) {
                case 1: color = position; alpha = 1; break;
                case 2: color = normal; alpha = 0; break;
                case 3: color = albedo; break;
                case 4: color = motion; break;
                case 5: color = material; break;
                case 6: color = emissive; break;
                default: color = vec3(0.0f); alpha = 0.0f; break;
                }
                out[4 * px] = color.x; out[4 * px + 1] = color.y; out[4 * px + 2] = color.z; out[4 * px + 3] = alpha;
                continue;
            }
            vec3 Lo(0.0f);
            for (size_t i = 0; i < S.lights.size(); i++) {
                const vkp_light& light = S.lights[i];
                const vec3 lpos(light.position[0], light.position[1], light.position[2]);
                const vec3 lcol(light.color[0], light.color[1], light.color[2]);
                bool isDirectional = light.position[3] < 0;
                vec3 L = isDirectional ? lpos : (lpos - position);
                vec3 Hh = normalize(V + normalize(L));
                float NdotL = maxf(dot(N, normalize(L)), 0.0f);
                if (isDirectional) {
                    Lo += (NdotL * lcol);
                } else {
                    float light_max_distance = light.position[3];
                    float light_distance = length(L);
                    float light_intensity = light.color[3] / (light_distance * light_distance);
                    float attenuation = light_max_distance - light_distance;
                    attenuation /= light_max_distance;
                    attenuation = maxf(attenuation, 0.0f);
                    attenuation = attenuation * attenuation;
                    vec3 radiance = lcol * light_intensity * attenuation;
                    float NDF = DistributionGGX(N, Hh, roughness);
                    float G = GeometrySmith(N, V, L, roughness);          /* un-normalised L: deferred.frag:104 */
                    vec3 F = FresnelSchlick(maxf(dot(Hh, V), 0.0f), F0);
                    vec3 kD = vec3(1.0f) - F;
                    kD *= 1.0f - metallic;
                    vec3 numerator = NDF * G * F;
                    float denominator = 4.0f * NdotV * maxf(dot(N, L), 0.0f);
                    vec3 specular = numerator / maxf(denominator, 0.001f);
                    Lo += (kD * pow3(albedo, 2.2f) / PI + specular) * radiance * NdotL;
                }
            }
            if (!background) {
                vec3 F = FresnelSchlick(NdotV, F0);
                vec3 kD = (1.0f - F) * (1.0f - metallic);
                vec3 diffuse = kD * albedo * irradiance;
                color = Lo + diffuse;
                color += emissive;
            } else {
                color = albedo;
            }
            out[4 * px] = color.x; out[4 * px + 1] = color.y; out[4 * px + 2] = color.z; out[4 * px + 3] = 1.0f;
        }
    }
}

/* ---- scalar helpers for known-answer tests ----------------------------------------------------- */
uint32_t orc_tea(uint32_t v0, uint32_t v1) { return tea(v0, v1); }
uint32_t orc_lcg(uint32_t* s) { return lcg(*s); }
float orc_rnd(uint32_t* s) { return rnd(*s); }
void orc_sample_disk(const vkp_light* light, const float pos[3], const float L[3], uint32_t seed, float out[3])
{
    vec3 d = sampleDisk(*light, vec3(pos[0], pos[1], pos[2]), vec3(L[0], L[1], L[2]), seed, false);
    out[0] = d.x; out[1] = d.y; out[2] = d.z;
}
float orc_distribution_ggx(const float N[3], const float H[3], float r) { return DistributionGGX(vec3(N[0], N[1], N[2]), vec3(H[0], H[1], H[2]), r); }
float orc_geometry_smith(const float N[3], const float V[3], const float L[3], float r)
{
    return GeometrySmith(vec3(N[0], N[1], N[2]), vec3(V[0], V[1], V[2]), vec3(L[0], L[1], L[2]), r);
}
void orc_fresnel_schlick(float c, const float F0[3], float out[3])
{
    vec3 f = FresnelSchlick(c, vec3(F0[0], F0[1], F0[2])); out[0] = f.x; out[1] = f.y; out[2] = f.z;
}
void orc_env_uv(const float d[3], int use_miss_pi, float out[2])
{
    vec2 uv = env_uv(vec3(d[0], d[1], d[2]), use_miss_pi ? PI_MISS : PI); out[0] = uv.x; out[1] = uv.y;
}
uint16_t orc_to_half(float f) { return to_half(f); }
float orc_from_half(uint16_t h) { return from_half(h); }
uint8_t orc_to_unorm8(float f) { return to_unorm8(f); }

} // extern "C"
