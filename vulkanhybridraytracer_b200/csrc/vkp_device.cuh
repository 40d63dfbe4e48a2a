// vkp_device.cuh -- device-side math and shading helpers shared by every kernel of libvkp.
//
// Numerics contract (DESIGN.md "Numerics"): all translation units are compiled with -fmad=false, IEEE
// div/sqrt (nvcc defaults), and every expression below keeps the left-to-right order of the GLSL source
// it implements, so results are bit-identical to a scalar fp32 evaluation of the same shader.  FMA is
// only used where it is written: the contract-v2 primitives of include/vkp_detmath.h (dot, cross, matrix
// rows, mix, reflect / refract -- the CPU checker calls the same functions) and the conservative box tests.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/vkp_types.h"
#include "../../include/vkp_detmath.h"

#define VKP_D __device__ __forceinline__

// ---- float3 in contract order -------------------------------------------------------------------------
VKP_D float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
VKP_D float3 f3s(float s) { return make_float3(s, s, s); }
VKP_D float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
VKP_D float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
VKP_D float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
VKP_D float3 operator/(float3 a, float3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
VKP_D float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
VKP_D float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
VKP_D float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
VKP_D float3 operator+(float3 a, float s) { return f3(a.x + s, a.y + s, a.z + s); }
VKP_D float3 operator-(float s, float3 a) { return f3(s - a.x, s - a.y, s - a.z); }
VKP_D float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
// contract v2: dot / cross / matrix rows / mix are the FMA forms of include/vkp_detmath.h, shared with the CPU checker
VKP_D float dot3(float3 a, float3 b) { return vkp_dot3(a.x, a.y, a.z, b.x, b.y, b.z); }
VKP_D float3 cross3(float3 a, float3 b)
{
    return f3(vkp_diff_of_products(a.y, b.z, a.z, b.y), vkp_diff_of_products(a.z, b.x, a.x, b.z), vkp_diff_of_products(a.x, b.y, a.y, b.x));
}
VKP_D float length3(float3 a) { return sqrtf(dot3(a, a)); }
VKP_D float3 normalize3(float3 a) { float inv = 1.0f / sqrtf(dot3(a, a)); return a * inv; }
VKP_D float clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
VKP_D float mix1(float x, float y, float a) { return vkp_mixf(x, y, a); }
VKP_D float3 mix3(float3 x, float3 y, float a) { return f3(vkp_mixf(x.x, y.x, a), vkp_mixf(x.y, y.y, a), vkp_mixf(x.z, y.z, a)); }
VKP_D float3 pow3(float3 a, float e) { return f3(vkp_powf(a.x, e), vkp_powf(a.y, e), vkp_powf(a.z, e)); }
VKP_D float3 reflect3(float3 I, float3 N) { const float k = 2.0f * dot3(N, I); return f3(vkp_nmsubf(N.x, k, I.x), vkp_nmsubf(N.y, k, I.y), vkp_nmsubf(N.z, k, I.z)); }
VKP_D float3 refract3(float3 I, float3 N, float eta)
{
    float d = dot3(N, I);
    float k = vkp_nmsubf(eta * eta, vkp_nmsubf(d, d, 1.0f), 1.0f);
    if (k < 0.0f) return f3s(0.0f);
    const float s = vkp_fmaf(eta, d, sqrtf(k));
    return f3(vkp_nmsubf(N.x, s, I.x * eta), vkp_nmsubf(N.y, s, I.y * eta), vkp_nmsubf(N.z, s, I.z * eta));
}

// column-major 4x4 (m[c*4+r]) applied to points / directions; same association as the oracle
VKP_D float3 xform_point(const float* __restrict__ m, float3 p)
{
    return f3(vkp_row_point(m[0], m[4], m[8], m[12], p.x, p.y, p.z),
              vkp_row_point(m[1], m[5], m[9], m[13], p.x, p.y, p.z),
              vkp_row_point(m[2], m[6], m[10], m[14], p.x, p.y, p.z));
}
VKP_D float3 xform_dir(const float* __restrict__ m, float3 d)
{
    return f3(vkp_dot3(m[0], m[4], m[8], d.x, d.y, d.z),
              vkp_dot3(m[1], m[5], m[9], d.x, d.y, d.z),
              vkp_dot3(m[2], m[6], m[10], d.x, d.y, d.z));
}
VKP_D float4 xform_vec4(const float* __restrict__ m, float4 v)
{
    float4 r;
    r.x = vkp_row_vec4(m[0], m[4], m[8], m[12], v.x, v.y, v.z, v.w);
    r.y = vkp_row_vec4(m[1], m[5], m[9], m[13], v.x, v.y, v.z, v.w);
    r.z = vkp_row_vec4(m[2], m[6], m[10], m[14], v.x, v.y, v.z, v.w);
    r.w = vkp_row_vec4(m[3], m[7], m[11], m[15], v.x, v.y, v.z, v.w);
    return r;
}
// mat3(transpose(M)) * n
VKP_D float3 xform_normal_by_inverse(const float* __restrict__ inv, float3 n)
{
    return f3(vkp_dot3(inv[0], inv[1], inv[2], n.x, n.y, n.z),
              vkp_dot3(inv[4], inv[5], inv[6], n.x, n.y, n.z),
              vkp_dot3(inv[8], inv[9], inv[10], n.x, n.y, n.z));
}

// ---- RNG: random.glsl:3-33 ---------------------------------------------------------------------------------
VKP_D uint32_t tea16(uint32_t val0, uint32_t val1)
{
    uint32_t v0 = val0, v1 = val1, s0 = 0u;
#pragma unroll
    for (int n = 0; n < 16; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
VKP_D float rnd01(uint32_t& state)
{
    state = 1664525u * state + 1013904223u;
    return (float)(state & 0x00ffffffu) / 16777216.0f;
}

// ---- device scene ---------------------------------------------------------------------------------------------
struct DTexture { const uchar4* texels; int w, h; };

struct DInstance {
    float model[16];
    float inv_model[16];
    const vkp_bvh_node* nodes;   // BLAS nodes
    const float4* tris;          // 3 x float4 per leaf-ordered triangle: (v0, prim id bits), (v1, 0), (v2, 0)
    const vkp_vertex* vertices;  // mesh vertex buffer
    const uint32_t* indices;     // mesh index buffer + first_index
    int32_t material_id;
    int32_t _pad[3];
};

// what traversal needs to enter an instance, in one 80-byte record (5 independent 16-byte loads = one memory round trip):
// the rows of inverse(model) 3x4 (applied with vkp_row_point / vkp_dot3), the BLAS arrays and the BLAS root reference
struct DInstTrav {
    float4 r0, r1, r2;
    const float4* nodes;        // BLAS nodes (32 B each)
    const float4* tris;
    uint32_t root_ref;          // nodes[0].ref of the BLAS
    uint32_t _pad[3];
};

struct DScene {
    const float* lut_pow22;
    const DInstTrav* inst_trav;          // same table as DFrame::lut_pow22 (hit / miss shading has only the scene)
    const float4* tlas_nodes2;       // TLAS nodes (2 x float4 each)
    uint32_t tlas_root_ref2;         // nodes[0].ref
    const DInstance* instances;
    const vkp_material* materials;
    const vkp_light* lights;
    const DTexture* textures;
    int n_instances, n_materials, n_lights, n_textures;
    int sky_tex, irr_tex;
};

// one in-flight reflection / refraction path of the wavefront (64 bytes)
struct PathItem {
    float4 a;   // origin.xyz, pixel index | background flag (bit 31)
    float4 b;   // direction.xyz, RNG seed bits
    float4 c;   // throughput (rayColor)
    float4 d;   // ambient + emissive tail (or the sky albedo for background pixels)
};
#define VKP_MAX_BOUNCES 64
#define VKP_CHAIN_MAX 8          /* stored per-bounce colours per path in chain mode 2 (MAX_RECURSION, helpers.glsl:8) */

struct DFrame {
    vkp_camera cam;
    vkp_params p;
    // multi-GPU tiling: this context renders tiles t with t % world == rank
    int tile_w, tile_h, tiles_x, tiles_y, rank, world;
    int count_trav;       // accumulate node fetches / triangle tests / instance entries into counters[4..6]
    int fused;            // k_resolve also runs the shadow filter (vkp_render); 0 when the stages are run one by one
    int fuse_accumulate;  // accumulation frames inside vkp_render: k_shadow's threads write the filter's new history byte themselves (no filter pass)
    int static_camera;    // prev_view/prev_proj are bit-identical to view/proj: motion vectors are exactly 0
    // planes
    uint2* position;      // 4 x fp16
    uint2* normal;        // 4 x fp16
    uchar4* albedo;
    uint32_t* motion;     // 2 x fp16
    uchar4* material;
    uchar4* emissive;
    int2* hit_ids;
    uint32_t* seed_w;               // k_gbuffer writes tea(pixel, frame) here (NULL: not wanted) ...
    const uint32_t* seed_r;         // ... and the later stages of the same frame read it (NULL: evaluate TEA; see pixel_seed)
    short* sky_shadow_w;            // k_gbuffer (block 0) writes, per light, the shadow byte of a sky pixel or -1 if sky pixels trace towards it ...
    const short* sky_shadow_r;      // ... and k_shadow of the same frame reads it (NULL: every thread does the arithmetic)
    float4* final_f32;
    uchar4* final_srgb8;
    uint8_t* shadow;      // n_lights planes
    uint8_t* denoised;    // n_lights planes (history, updated in place through denoised_next)
    uint8_t* denoised_next;
    unsigned long long* counters;   // [4] rays by type
    const float* lut_pow22;         // [256] vkp_powf(k/255, 2.2): decode of 8-bit albedo / sky texels (bit-identical to the call)
    const uint8_t* lut_post;        // [256] postFragment.frag on an 8-bit input: unorm8(pow(max(k/255, 0.001), 1/2.2))
    PathItem* path_q[2];            // ping-pong path queues of the reflection wavefront
    unsigned int* path_count;       // [VKP_MAX_BOUNCES + 2] queue length per bounce
    float4* chain_buf;              // chain mode 2: VKP_CHAIN_MAX colours per compacted path
    int* chain_len;                 // chain mode 2: number of bounces traced per path
    int* chain_slot;                // chain mode 2: per-pixel path slot (written for mirror / glass pixels only)
    int* chain_pix;                 // chain mode 2: pixel index of every path slot (the queue k_chain's warps pull from)
    PathItem* chain_base;           // schedules 5-7: per path slot, the direct term (c) and the ambient tail (d) k_resolve<3> leaves for k_resolve_paths
    uint8_t* chain_long;            // per pixel: the path of this pixel went beyond its first bounce in the previous frame (schedule 7's prediction)
    int chain_split;                // schedule 7: two path lists (see k_chain_scan)
    int* chain_vl_list;             // schedules 7 / 8: slots of the paths predicted very long (k_chain_quad's list; NULL = no such split)
    uint8_t* chain_kind;            // per slot of the long list: 1 = very long (k_chain skips it)
    int chain_quad_all;             // VKP_CHAIN_LONG_QUAD=1: k_chain_quad walks the whole long list
    unsigned int chain_cap;         // slots in the path arrays (= pixels)
    int chain_in_gbuffer;           // schedule 8: k_gbuffer compacts the path pixels itself (no k_chain_scan)
    unsigned int first_blocks;      // schedule 8: leading blocks of k_shadow's grid that trace the queued paths' first rays (0 otherwise)
    unsigned int mid_at;            // schedule 8: first_blocks * n_lights blocks from this shadow-block index on trace the first hits' any-hit rays
    uint32_t chain_tag;             // schedule 8: "record complete" mark of this frame's first-bounce records (never 0)
    // longest-first CTA scheduling: cost[v] = SM cycles the slowest warp of virtual block v took in the
    // previous frame, order[b] = the virtual block hardware block b works on (NULL = identity).  [0] k_gbuffer, [1] k_shadow.
    const uint32_t* cta_order[2];
    uint32_t* cta_cost[2];
    int n_cta;                      // CTAs per pixel pass on this rank (k_shadow runs n_cta * n_lights of them)
    // fast_div (kernels_frame.cu): floor(2^40 / d) + 1 for the four divisors of the block-index arithmetic
    unsigned long long magic_bpt, magic_tiles_x, magic_bpt_x, magic_n_cta;
    int div_exact;                  // all four are exact for the index ranges of this launch (n * d < 2^40)
};

// k / 255.0f for the 256 possible bytes, filled on the device by k_build_luts with the same IEEE division the contract
// names (a multiplication by 1/255 differs in 126 of the 256 cases): one L1-resident load instead of a division sequence
static __device__ float g_unit255[256];
VKP_D float unit255(unsigned char x) { return __ldg(&g_unit255[x]); }

// ---- texture / env -----------------------------------------------------------------------------------------
// REPEAT addressing: i mod n in [0, n).  Power-of-two sizes (every texture of the default scene) take a mask -- the two's
// complement AND is the positive modulo -- instead of the ~20-instruction integer division (4.5 % of k_resolve's instructions).
VKP_D int wrap_repeat(int i, int n)
{
    if ((n & (n - 1)) == 0) return i & (n - 1);
    i %= n;
    return i < 0 ? i + n : i;
}

// texture(): VK_FILTER_NEAREST + REPEAT on R8G8B8A8_UNORM (vk_textures.cpp:15,26; vk_initializers.h:86)
VKP_D float4 tex_nearest(const DTexture& T, float u, float v)
{
    float fx = u * (float)T.w, fy = v * (float)T.h;
    int ix = (fx == fx) ? (int)floorf(fx) : 0, iy = (fy == fy) ? (int)floorf(fy) : 0;
    ix = wrap_repeat(ix, T.w);
    iy = wrap_repeat(iy, T.h);
    uchar4 t = __ldg(&T.texels[(size_t)iy * T.w + ix]);
    return make_float4(unit255(t.x), unit255(t.y), unit255(t.z), unit255(t.w));
}
VKP_D uchar4 tex_nearest_raw(const DTexture& T, float u, float v)
{
    float fx = u * (float)T.w, fy = v * (float)T.h;
    int ix = (fx == fx) ? (int)floorf(fx) : 0, iy = (fy == fy) ? (int)floorf(fy) : 0;
    ix = wrap_repeat(ix, T.w);
    iy = wrap_repeat(iy, T.h);
    return __ldg(&T.texels[(size_t)iy * T.w + ix]);
}
VKP_D float3 xyz(float4 v) { return f3(v.x, v.y, v.z); }

#define VKP_PI 3.14159265359f        /* helpers.glsl:5 */
#define VKP_PI_MISS 3.141592f        /* miss.rmiss:11  */

// miss.rmiss:16, hybridRaygen.rgen:59
VKP_D float2 equirect_uv(float3 d, float pi)
{
    float2 uv;
    uv.x = 0.5f + vkp_atan2f(d.x, d.z) / (2.0f * pi);
    uv.y = 0.5f - vkp_asinf(d.y) / pi;
    return uv;
}

// ---- BRDF: helpers.glsl:33-93 ------------------------------------------------------------------------------
VKP_D float distribution_ggx(float3 N, float3 H, float roughness)
{
    float a = roughness * roughness;
    float a2 = a * a;
    float NdotH = fmaxf(dot3(N, H), 0.0f);
    float NdotH2 = NdotH * NdotH;
    float denom = (NdotH2 * (a2 - 1.0f) + 1.0f);
    denom = VKP_PI * denom * denom;
    return a2 / denom;
}
VKP_D float geometry_schlick_ggx(float NdotV, float roughness)
{
    float r = roughness + 1.0f;
    float k = (r * r) / 8.0f;
    float denom = NdotV * (1.0f - k) + k;
    return NdotV / denom;
}
VKP_D float geometry_smith(float3 N, float3 V, float3 L, float roughness)
{
    float NdotV = fmaxf(dot3(N, V), 0.0f);
    float NdotL = fmaxf(dot3(N, L), 0.0f);
    return geometry_schlick_ggx(NdotV, roughness) * geometry_schlick_ggx(NdotL, roughness);
}
VKP_D float3 fresnel_schlick(float cosTheta, float3 F0)
{
    return F0 + (1.0f - F0) * vkp_powf(1.0f - cosTheta, 5.0f);
}
// helpers.glsl:82-93; the reference passes `seed` by value (faithful); advance=true writes it back
VKP_D float3 sample_disk(const vkp_light& light, float3 position, float3 L, uint32_t& seed_io, bool advance)
{
    uint32_t seed = seed_io;
    float radius = light.radius * sqrtf(rnd01(seed));
    float angle = rnd01(seed) * 2.0f * VKP_PI;
    float sa, ca; vkp_sincosf(angle, &sa, &ca);
    float px = radius * ca, py = radius * sa;
    float3 tangent = normalize3(cross3(L, f3(0.0f, 1.0f, 0.0f)));
    if (L.x == 0.0f && L.y == 1.0f && L.z == 0.0f) tangent = f3(0.0f, 0.0f, 1.0f);
    float3 bitangent = normalize3(cross3(tangent, L));
    float3 target = position + L + px * tangent + py * bitangent;
    if (advance) seed_io = seed;
    return normalize3(target - position);
}

// ---- storage formats ---------------------------------------------------------------------------------------
VKP_D uint2 pack_half4(float x, float y, float z, float w)
{
    return make_uint2((uint32_t)vkp_f32_to_f16(x) | ((uint32_t)vkp_f32_to_f16(y) << 16),
                      (uint32_t)vkp_f32_to_f16(z) | ((uint32_t)vkp_f32_to_f16(w) << 16));
}
VKP_D float4 unpack_half4(uint2 v)
{
    return make_float4(vkp_f16_to_f32((uint16_t)(v.x & 0xffffu)), vkp_f16_to_f32((uint16_t)(v.x >> 16)),
                       vkp_f16_to_f32((uint16_t)(v.y & 0xffffu)), vkp_f16_to_f32((uint16_t)(v.y >> 16)));
}
VKP_D uchar4 pack_unorm4(float x, float y, float z, float w)
{
    return make_uchar4(vkp_unorm8(x), vkp_unorm8(y), vkp_unorm8(z), vkp_unorm8(w));
}
VKP_D float3 lut3(const float* __restrict__ lut, uchar4 c) { return f3(__ldg(&lut[c.x]), __ldg(&lut[c.y]), __ldg(&lut[c.z])); }
VKP_D float3 unorm3(uchar4 c) { return f3(unit255(c.x), unit255(c.y), unit255(c.z)); }

// camera ray of raygen.rgen:32-40
VKP_D void camera_ray(const vkp_camera& cam, int x, int y, float ox, float oy, int W, int H, float3& origin, float3& dir)
{
    float pcx = (float)x + ox, pcy = (float)y + oy;
    float u = pcx / (float)W, v = pcy / (float)H;
    float dx = u * 2.0f - 1.0f, dy = v * 2.0f - 1.0f;
    origin = f3(cam.inv_view[12], cam.inv_view[13], cam.inv_view[14]);
    float4 target = xform_vec4(cam.inv_proj, make_float4(dx, dy, 1.0f, 1.0f));
    float3 tn = normalize3(xyz(target));
    dir = xform_dir(cam.inv_view, tn);
}

VKP_D const vkp_material& material_by_float_index(const DScene& S, float matIdx)
{
    int i = (int)matIdx;
    i = max(0, min(i, S.n_materials - 1));
    return S.materials[i];
}
