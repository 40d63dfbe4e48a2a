// kernels_frame.cu -- the frame kernels of the Hybrid / Raytraced / Deferred pipelines, one per reference GPU
// program (SURVEY.md section 7 kernel plan):
//   k_gbuffer   basic.vert + geometry_shader.frag + skybox.frag (G-buffer content, ray-cast primary visibility)
//   k_shadow    shadowRaygen.rgen + shadow.rmiss (any-hit shadow rays seeded from the G-buffer)
//   k_denoise   blur.comp
//   k_resolve   hybridRaygen.rgen + hybridHit.rchit + miss.rmiss (+ postFragment.frag fused into the store)
//   k_raytrace  raygen.rgen + closesthit.rchit + miss.rmiss
//   k_deferred  deferred.frag
//   k_post      postFragment.frag (stand-alone, for stage tests and the non-fused modes)
// Thread mapping: one thread per pixel; a warp covers an 8x4 pixel patch (coherent rays, 64-byte row segments
// for the fp16x4 planes), a 128-thread CTA covers a 16x8 block; CTAs enumerate only the screen tiles this
// rank owns (tile t belongs to rank t % world).
#include "vkp_internal.h"
#include "vkp_bvh.cuh"
#include <cstdlib>

// occupancy knobs (min resident CTAs per SM -> register cap), tuned by measurement
#ifndef VKP_LB_GBUFFER
#define VKP_LB_GBUFFER 5      /* 96 registers.  Round 1 (local-memory stack): 4 CTAs / 128 registers beat 8 / 64 (0.148 vs 0.183 ms at 1080p).  Round 2
                                 (shared-memory stack): 3 / 4 / 5 / 6 CTAs -> 0.152 / 0.151 / 0.143 / 0.166 ms (profiles/r2p_*) */
#endif
#ifndef VKP_LB_SHADOW
#define VKP_LB_SHADOW 8       /* latency-bound: more resident warps beat the 176 bytes it spills at 64 registers */
#endif
#ifndef VKP_LB_SHADOW_MANY
#define VKP_LB_SHADOW_MANY 5
#endif
#ifndef VKP_LB_RT
#define VKP_LB_RT 4
#endif
#ifndef VKP_LB_RESOLVE
#define VKP_LB_RESOLVE 6      /* 4 / 5 / 6 CTAs -> 0.137 / 0.121 / 0.110 ms for the resolve stage (profiles/r2p_*) */
#endif
#ifndef VKP_CHAIN_CALL
#define VKP_CHAIN_CALL 0      /* 1: k_chain calls the traversal (trace_ray_call) instead of inlining it */
#endif
#ifndef VKP_LB_SHADOW_STREAM
#define VKP_LB_SHADOW_STREAM 4
#endif
#ifndef VKP_KEEPW_MANY
#define VKP_KEEPW_MANY 1      /* many-sample shadow loop and Raytraced mode: trace_ray<.., KEEPW> (vkp_bvh.cuh) */
#endif
#ifndef VKP_CHAIN_VL_BOUNCES
#define VKP_CHAIN_VL_BOUNCES 4  /* paths with at least this many bounces last frame go to k_chain_quad (schedules 7 / 8) */
#endif
#ifndef VKP_LB_CHAIN_TRACE
#define VKP_LB_CHAIN_TRACE 8  /* the lean traversal kernels of the staged chain (k_chain_first, k_chain_shadow0) */
#endif
#ifndef VKP_LB_CHAIN
#define VKP_LB_CHAIN 3        /* 168 registers: the sparse chain kernel gains nothing from occupancy, a lot from not spilling */
#endif

namespace {

struct Pixel { int x, y; bool valid; };

// n / d for the block-index arithmetic every thread of every frame kernel starts with.  d is a launch constant the compiler cannot
// see, so `/` and `%` were ~20-instruction sequences: 10 % of all instructions k_shadow executed (ncu source page,
// profiles/r1o_*).  magic = floor(2^40 / d) + 1 (host, make_launch) is exact while n * d < 2^40; otherwise F.div_exact is 0.
__device__ __forceinline__ unsigned int fast_div(const DFrame& F, unsigned int n, unsigned long long magic, unsigned int d)
{
    return F.div_exact ? (unsigned int)(((unsigned long long)n * magic) >> 40) : n / d;
}

__device__ __forceinline__ Pixel pixel_of_block(const DFrame& F, const unsigned int block, const int warp_override = -1)
{
    const int bpt_x = F.tile_w >> 4, bpt_y = F.tile_h >> 3;
    const int bpt = bpt_x * bpt_y;
    const int own_tile = (int)fast_div(F, block, F.magic_bpt, (unsigned int)bpt), sub = (int)block - own_tile * bpt;
    const int tile = F.rank + own_tile * F.world;
    const int ty = (int)fast_div(F, (unsigned int)tile, F.magic_tiles_x, (unsigned int)F.tiles_x), tx = tile - ty * F.tiles_x;
    const int sby = (int)fast_div(F, (unsigned int)sub, F.magic_bpt_x, (unsigned int)bpt_x), sbx = sub - sby * bpt_x;
    const int warp = warp_override >= 0 ? warp_override : (int)(threadIdx.x >> 5), lane = threadIdx.x & 31;
    Pixel p;
    p.x = tx * F.tile_w + sbx * 16 + (warp & 1) * 8 + (lane & 7);
    p.y = ty * F.tile_h + sby * 8 + (warp >> 1) * 4 + (lane >> 3);
    p.valid = p.x < F.p.width && p.y < F.p.height;
    return p;
}

__device__ __forceinline__ Pixel pixel_of_thread(const DFrame& F) { return pixel_of_block(F, blockIdx.x); }

// The per-pixel RNG seed of shadowRaygen.rgen:31 / hybridRaygen.rgen:31: tea(pixel index, frame), 16 rounds = ~190 integer instructions.
// k_shadow runs one thread per (pixel, light), the chain one per path: inside a frame (vkp_render) k_gbuffer evaluates it once per pixel
// into a plane and the later stages load it (26 % of k_shadow's instructions were TEA).  A stage run on its own has no such plane
// (seed_r == NULL) and evaluates it.
__device__ __forceinline__ uint32_t pixel_seed(const DFrame& F, size_t i, int x, int y)
{
    // (k_gbuffer fills the plane for the pixels that hit something; a sky pixel that does need a seed -- it rarely traces anything --
    // evaluates it: that keeps 190 instructions out of 57 % of k_gbuffer's threads)
    if (F.seed_r && F.hit_ids[i].x >= 0) return __ldg(F.seed_r + i);
    return tea16((uint32_t)y * (uint32_t)F.p.width + (uint32_t)x, (uint32_t)F.p.frame);
}

// Longest-first scheduling.  A frame kernel ends with its slowest thread, and which pixels are slow (rays grazing the dense meshes)
// hardly changes from one frame to the next: every warp leaves the cycles it took in cost[virtual block]; k_build_order turns
// last frame's costs into this frame's block order, expensive blocks first, so that the long traversals start at t = 0
// instead of in the last wave.  Only the schedule changes -- every pixel is still computed by exactly one thread.
__device__ __forceinline__ unsigned int virtual_block(const uint32_t* __restrict__ order) { return order ? __ldg(order + blockIdx.x) : blockIdx.x; }
__device__ __forceinline__ void record_cost(uint32_t* __restrict__ cost, unsigned int vblock, long long t0)
{
    if (!cost) return;
    const long long dt = clock64() - t0;          // the warp is converged here (after the warp-wide counter reduction)
    if ((threadIdx.x & 31) == 0) atomicMax(cost + vblock, (uint32_t)(dt > 0xffffffffLL ? 0xffffffffLL : dt));
}

// One CTA of 256 threads per cost array (small enough to slip in beside the frame kernels): stable partition of the virtual blocks into 4 cost classes (> 8x, > 4x, > 2x the mean, rest);
// resets the costs for the next frame.  Runs off the critical path on a side stream.
__global__ void __launch_bounds__(256) k_build_order(uint32_t* __restrict__ cost, uint32_t* __restrict__ order, uint32_t n)
{
    __shared__ unsigned long long s_sum;
    __shared__ uint32_t s_cnt[4][256 + 1];
    const uint32_t t = threadIdx.x;
    const uint32_t chunk = (n + 255u) / 256u, lo = min(t * chunk, n), hi = min(lo + chunk, n);
    if (t == 0) s_sum = 0ull;
    __syncthreads();
    unsigned long long sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += cost[i];
    for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    if ((t & 31u) == 0u && sum) atomicAdd(&s_sum, sum);
    __syncthreads();
    const unsigned long long mean = s_sum / (n ? n : 1u);
    uint32_t c[4] = {0u, 0u, 0u, 0u};
    for (uint32_t i = lo; i < hi; i++) { const unsigned long long v = cost[i]; c[v > 8ull * mean ? 0 : v > 4ull * mean ? 1 : v > 2ull * mean ? 2 : 3]++; }
#pragma unroll
    for (int k = 0; k < 4; k++) s_cnt[k][t] = c[k];
    __syncthreads();
    if (t < 4) {      // exclusive scan over the 256 chunks of class t (serial, off the critical path)
        uint32_t run = 0;
        for (uint32_t j = 0; j < 256u; j++) { const uint32_t v = s_cnt[t][j]; s_cnt[t][j] = run; run += v; }
        s_cnt[t][256] = run;
    }
    __syncthreads();
    uint32_t base[4];
    base[0] = 0u; base[1] = s_cnt[0][256]; base[2] = base[1] + s_cnt[1][256]; base[3] = base[2] + s_cnt[2][256];
    uint32_t pos[4];
#pragma unroll
    for (int k = 0; k < 4; k++) pos[k] = base[k] + s_cnt[k][t];
    for (uint32_t i = lo; i < hi; i++) {
        const unsigned long long v = cost[i];
        const int k = v > 8ull * mean ? 0 : v > 4ull * mean ? 1 : v > 2ull * mean ? 2 : 3;
        order[pos[k]++] = i;
        cost[i] = 0u;
    }
}

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_volatile_u32(const unsigned int* p) { return *reinterpret_cast<const volatile unsigned int*>(p); }

__device__ __forceinline__ void add_counter(unsigned long long* counters, int slot, unsigned int n)
{
    n = __reduce_add_sync(0xffffffffu, n);      // redux.sync: one instruction instead of five shuffle + add pairs
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(&counters[slot], (unsigned long long)n);
}

__device__ __forceinline__ void add_trav(unsigned long long* counters, const TravCount& tc, unsigned int tag = 0u)
{
    add_counter(counters, 4, tc.nodes); add_counter(counters, 5, tc.tris); add_counter(counters, 6, tc.entries);
    // the thread with the longest traversal sets the kernel's tail: remember it (stage | light | pixel)
    if (tc.nodes > 128u) atomicMax(&counters[7], ((unsigned long long)tc.nodes << 32) | tag);
}

// image store into the B8G8R8A8_UNORM target (renderer.cpp:1535) followed by postFragment.frag:9-21
// (the 8-bit store leaves 256 possible inputs per channel: the gamma curve is a table built on the device by k_build_luts)
__device__ __forceinline__ uchar4 post_pixel(const uint8_t* __restrict__ lut_post, float4 c)
{
    return make_uchar4(__ldg(&lut_post[vkp_unorm8(c.x)]), __ldg(&lut_post[vkp_unorm8(c.y)]), __ldg(&lut_post[vkp_unorm8(c.z)]), vkp_unorm8(c.w));
}

__global__ void k_build_luts(float* lut_pow22, uint8_t* lut_post)
{
    const int k = threadIdx.x;
    const float x = (float)k / 255.0f;
    g_unit255[k] = x;
    lut_pow22[k] = vkp_powf(x, 2.2f);
    lut_post[k] = vkp_unorm8(vkp_powf(fmaxf(x, 0.001f), 1.0f / 2.2f));
}

// warp-aggregated queue append: one atomic per warp, slots assigned with ballot + popc prefix (ray compaction)
__device__ __forceinline__ void path_push(PathItem* q, unsigned int* count, bool want, const PathItem& item)
{
    const unsigned mask = __ballot_sync(0xffffffffu, want);
    if (!mask) return;
    const int lane = threadIdx.x & 31;
    unsigned base = 0;
    if (lane == (__ffs(mask) - 1)) base = atomicAdd(count, (unsigned)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, __ffs(mask) - 1);
    if (want) q[base + __popc(mask & ((1u << lane) - 1u))] = item;
}

struct Payload { float3 color; float dist; float3 direction; float scatter; float3 origin; uint32_t seed; };

// trace_ray as a real call (CALL == true below): in the chain kernels the hit shader's state and the traversal's state then do not
// have to be in registers at the same time -- what is live across the call goes to the stack once per ray instead of the register
// allocator spilling inside the node loop
template <bool ANY>
__device__ __noinline__ bool trace_ray_call(const DScene& S, uint32_t* __restrict__ sstack, float3 o, float3 d, float tmin, float tmax, HitRec& best, TravCount* tcount)
{
    return trace_ray<ANY>(S, sstack, o, d, tmin, tmax, best, nullptr, tcount);
}

// miss.rmiss:13-19
__device__ __forceinline__ void shade_miss(const DScene& S, float3 rayDir, Payload& prd)
{
    float3 dir = normalize3(rayDir);
    float2 uv = equirect_uv(dir, VKP_PI_MISS);
    prd.color = lut3(S.lut_pow22, tex_nearest_raw(S.textures[S.sky_tex], uv.x, uv.y));
    prd.dist = -1.0f;
}

// hybridHit.rchit:24-184 (RT == false) / closesthit.rchit:26-186 (RT == true)
// WARP != 0: the 32 lanes of the warp hold the SAME path; everything is computed redundantly (one SIMT instruction stream,
// broadcast loads) except the shadow rays: WARP == 1 (k_raytrace_warp) spreads the per-light shadow samples over the lanes,
// WARP == 2 (k_chain) traces the one shadow ray per light with the whole warp (warp_trace).
// PRE (Hybrid only): the per-light secondary shadow rays of this hit were traced by k_chain_shadow0 (staged chain, mode 4);
// bit l of pre_mask = the ray towards light l was blocked.  Nothing is traced here; the RNG still advances as if it were.
template <bool RT, int WARP = 0, bool PRE = false, bool CALL = false>
__device__ void shade_hit(const DScene& S, uint32_t* __restrict__ sstack, const vkp_params& P, float3 rayO, float3 rayD, const HitRec& h, Payload& prd,
                          unsigned int& n_secondary, TravCount* tcp, const WarpQueue* wq = nullptr, uint32_t pre_mask = 0u)
{
    const DInstance& I = S.instances[h.inst];
    const float b0 = 1.0f - h.u - h.v, b1 = h.u, b2 = h.v;
    const uint32_t* ind = I.indices + 3u * (uint32_t)h.prim;
    const vkp_vertex& v0 = I.vertices[__ldg(ind)]; const vkp_vertex& v1 = I.vertices[__ldg(ind + 1)]; const vkp_vertex& v2 = I.vertices[__ldg(ind + 2)];
    const float3 normal = f3(v0.normal[0], v0.normal[1], v0.normal[2]) * b0 + f3(v1.normal[0], v1.normal[1], v1.normal[2]) * b1 +
                          f3(v2.normal[0], v2.normal[1], v2.normal[2]) * b2;
    const float uvx = v0.uv[0] * b0 + v1.uv[0] * b1 + v2.uv[0] * b2;
    const float uvy = v0.uv[1] * b0 + v1.uv[1] * b1 + v2.uv[1] * b2;
    const float3 N = normalize3(xform_normal_by_inverse(I.inv_model, normal));
    const float3 V = normalize3(-rayD);
    const float NdotV = clamp01(dot3(N, V));
    const float3 worldPos = rayO + rayD * h.t;

    float3 Lo = f3s(0.0f);
    float attenuation = 1.0f;

    const vkp_material& mat = S.materials[max(0, min(I.material_id, S.n_materials - 1))];
    const int shadingMode = (int)mat.shadingMetallicRoughness[0];
    float3 albedo = mat.textures[0] > -1 ? lut3(S.lut_pow22, tex_nearest_raw(S.textures[(int)mat.textures[0]], uvx, uvy))
                                         : pow3(f3(mat.diffuse[0], mat.diffuse[1], mat.diffuse[2]), 2.2f);
    const float3 emissive = mat.textures[2] > -1 ? xyz(tex_nearest(S.textures[(int)mat.textures[2]], uvx, uvy)) : f3s(0.0f);
    const float3 roughnessMetallic = mat.textures[3] > -1 ? xyz(tex_nearest(S.textures[(int)mat.textures[3]], uvx, uvy))
                                                          : f3(0.0f, mat.shadingMetallicRoughness[2], mat.shadingMetallicRoughness[1]);
    const float roughness = roughnessMetallic.y, metallic = roughnessMetallic.z;
    const float3 F0 = mix3(f3s(0.04f), albedo, metallic);

    const float2 envUV = equirect_uv(N, VKP_PI);
    const float3 irradiance = xyz(tex_nearest(S.textures[S.irr_tex], envUV.x, envUV.y));

    float3 direction = f3(1.0f, 1.0f, 1.0f);
    float scatter = 0.0f;
    const bool advance = P.faithful_rng == 0;

    for (int i = 0; i < S.n_lights; i++) {
        const vkp_light light = S.lights[i];
        const float3 lpos = f3(light.position[0], light.position[1], light.position[2]);
        const float3 lcol = f3(light.color[0], light.color[1], light.color[2]);
        const bool isDirectional = light.position[3] < 0.0f;
        float3 L = isDirectional ? lpos : (lpos - worldPos);
        const float light_max_distance = light.position[3];
        const float light_distance = length3(L);
        L = normalize3(L);
        const float light_intensity = isDirectional ? 1.0f : (light.color[3] / (light_distance * light_distance));
        const float3 H = normalize3(V + L);
        const float NdotL = clamp01(dot3(N, L));
        const float NdotH = clamp01(dot3(N, H));
        float shadowFactor = RT ? 0.0f : 1.0f;

        if (RT && WARP == 1) {
            // closesthit.rchit:99-120 with the samples of this light traced by different lanes.  Every traced sample
            // advances the RNG by exactly two LCG steps whatever it hits, so sample a starts from LCG^(2a)(seed); the
            // unshadowed count is an integer, so the sum does not depend on the order.
            if (NdotL > 0.0f) {
                unsigned lit = 0;
                if (light_distance < light_max_distance) {
                    const int lane = threadIdx.x & 31;
                    for (int a0 = 0; a0 < P.shadow_samples; a0 += 32) {
                        const int a = a0 + lane;
                        bool unshadowed = false;
                        if (a < P.shadow_samples) {
                            uint32_t sd = prd.seed;
                            if (advance) for (int r = 0; r < 2 * a; r++) sd = 1664525u * sd + 1013904223u;
                            float3 dir = sample_disk(light, worldPos, L, sd, false);
                            HitRec sh; n_secondary++;
                            unshadowed = !trace_ray<true, false, VKP_KEEPW_MANY != 0>(S, sstack, worldPos + dir * 1e-2f, dir, 0.001f, light_distance + 1.0f, sh, nullptr, tcp);
                        }
                        lit += (unsigned)__popc(__ballot_sync(0xffffffffu, unshadowed));
                    }
                    if (advance) for (int r = 0; r < 2 * P.shadow_samples; r++) prd.seed = 1664525u * prd.seed + 1013904223u;
                }
                shadowFactor = (float)lit;            // == lit times "+= 1.0f"
                shadowFactor /= (float)P.shadow_samples;
            }
            if (light_intensity == 0.0f) attenuation = 0.0f;
            else {
                attenuation = light_max_distance - light_distance;
                attenuation /= light_max_distance;
                attenuation = fmaxf(attenuation, 0.0f);
                attenuation = attenuation * attenuation;
            }
        } else if (RT) {
            if (NdotL > 0.0f) {
                for (int a = 0; a < P.shadow_samples; a++) {
                    bool shadowed = true;
                    if (light_distance < light_max_distance) {
                        float3 dir = sample_disk(light, worldPos, L, prd.seed, advance);
                        HitRec sh; n_secondary++;
                        shadowed = trace_ray<true>(S, sstack, worldPos + dir * 1e-2f, dir, 0.001f, light_distance + 1.0f, sh, nullptr, tcp);
                    }
                    if (!shadowed) shadowFactor += 1.0f;
                }
                shadowFactor /= (float)P.shadow_samples;
            }
            if (light_intensity == 0.0f) attenuation = 0.0f;
            else {
                attenuation = light_max_distance - light_distance;
                attenuation /= light_max_distance;
                attenuation = fmaxf(attenuation, 0.0f);
                attenuation = attenuation * attenuation;
            }
        } else {
            if (light_intensity == 0.0f) attenuation = 0.0f;
            else {
                attenuation = light_max_distance - light_distance;
                attenuation /= light_max_distance;
                attenuation = fmaxf(attenuation, 0.0f);
                attenuation = isDirectional ? 0.3f : attenuation * attenuation;
            }
        }

        if (shadingMode == 0) {
            if (!RT) {
                if (NdotL > 0.0f) {
                    bool shadowed = true;
                    if (light_distance < light_max_distance) {
                        if (PRE) {
                            if (advance) { prd.seed = 1664525u * prd.seed + 1013904223u; prd.seed = 1664525u * prd.seed + 1013904223u; }   // the two rnd() of sampleDisk
                            shadowed = ((pre_mask >> i) & 1u) != 0u;
                        } else {
                            float3 dir = sample_disk(light, worldPos, L, prd.seed, advance);
                            HitRec sh; n_secondary++;
                            if (WARP == 2) shadowed = warp_trace<true>(S, *wq, worldPos + dir * 1e-2f, dir, 0.001f, light_distance + 1.0f, sh, tcp);
                            else if (CALL) shadowed = trace_ray_call<true>(S, sstack, worldPos + dir * 1e-2f, dir, 0.001f, light_distance + 1.0f, sh, tcp);
                            else shadowed = trace_ray<true>(S, sstack, worldPos + dir * 1e-2f, dir, 0.001f, light_distance + 1.0f, sh, nullptr, tcp);
                        }
                    }
                    if (!shadowed) shadowFactor += 1.0f;
                }
            }
            float3 radiance = light_intensity * lcol * attenuation * shadowFactor;
            float3 F = fresnel_schlick(NdotH, F0);
            float D = distribution_ggx(N, H, roughness);
            float G = geometry_smith(N, V, L, roughness);
            float3 numerator = (D * G) * F;
            float denominator = fmaxf(4.0f * clamp01(dot3(N, V)) * NdotL, 0.000001f);
            float3 specular = numerator / denominator;
            float3 kD = (f3s(1.0f) - F) * (1.0f - metallic);
            Lo = Lo + (kD * albedo / VKP_PI + specular) * radiance * NdotL;
            direction = f3(1.0f, 1.0f, 1.0f); scatter = 0.0f;
        } else if (shadingMode == 3) {
            const float3 reflected = reflect3(normalize3(rayD), N);
            const bool isScattered = dot3(reflected, N) > 0.0f;
            float3 add;
            if (NdotL > 0.0f && light_intensity > 0.0f)
                add = RT ? light_intensity * lcol * attenuation * shadowFactor * metallic * albedo
                         : light_intensity * lcol * attenuation * albedo * metallic;
            else
                add = irradiance * albedo * metallic;
            Lo = Lo + add;
            direction = reflected; scatter = isScattered ? 1.0f : 0.0f;
        } else if (shadingMode == 4) {
            const float ior = mat.diffuse[3];
            const float NdotVg = dot3(N, normalize3(rayD));
            const float3 refrNormal = NdotVg > 0.0f ? -N : N;
            const float refrEta = NdotVg > 0.0f ? 1.0f / ior : ior;
            Lo = Lo + ((light_intensity > 0.0f) ? light_intensity * lcol * attenuation * albedo * metallic : irradiance * albedo * metallic);
            float radicand = 1.0f + vkp_powf(refrEta, 2.0f) * (NdotVg * NdotVg - 1.0f);
            if (radicand < 0.0f) direction = reflect3(rayD, N);
            else direction = RT ? refract3(rayD, refrNormal, refrEta) : refract3(normalize3(rayD), refrNormal, refrEta);
            scatter = 1.0f;
        }
    }

    float3 F = fresnel_schlick(NdotV, F0);
    float3 kD = (1.0f - F) * (1.0f - metallic);
    float3 diffuse = kD * albedo * irradiance;
    prd.color = Lo + diffuse + emissive;
    prd.dist = h.t; prd.direction = direction; prd.scatter = scatter; prd.origin = worldPos;
}

// per-light pieces of shadowRaygen.rgen:55-61 that decide whether rays are traced (and hence how far the RNG advances)
struct LightGeom { float3 L; float dist, maxd, NdotL; };
__device__ __forceinline__ LightGeom light_geom(const vkp_light& light, float3 position, float3 N)
{
    LightGeom g;
    const float3 lpos = f3(light.position[0], light.position[1], light.position[2]);
    const bool isDirectional = light.position[3] < 0.0f;
    float3 L = isDirectional ? lpos : (lpos - position);
    g.maxd = light.position[3];
    g.dist = length3(L);
    g.L = normalize3(L);
    g.NdotL = clamp01(dot3(N, g.L));
    return g;
}

// (defined with the chain kernels below; schedule 8 calls them from k_gbuffer and k_shadow)
__device__ __forceinline__ void chain_enqueue(const DFrame& F, size_t i, bool path);
__device__ __forceinline__ void chain_first_paths(const DScene& S, const DFrame& F, uint32_t* __restrict__ sstack, unsigned int k0, unsigned int stride,
                                                  unsigned int& n_refl, TravCount* tcp);
__device__ __forceinline__ void chain_shadow0_paths(const DScene& S, const DFrame& F, uint32_t* __restrict__ sstack, int l, unsigned int k0, unsigned int stride,
                                                    unsigned int& n_sec, TravCount* tcp, bool wait_ready);

// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VKP_TRAV_BLOCK, VKP_LB_GBUFFER) k_gbuffer(const __grid_constant__ DScene S, const __grid_constant__ DFrame F)
{
    __shared__ uint32_t s_trav[VKP_TRAV_SMEM_WORDS];
    uint32_t* const sstack = s_trav + threadIdx.x;
    const long long t_start = clock64();
    const unsigned int vb = virtual_block(F.cta_order[0]);
    const Pixel px = pixel_of_block(F, vb);
    unsigned int nrays = 0;
    TravCount tc = {0u, 0u, 0u};
    TravCount* const tcp = F.count_trav ? &tc : nullptr;      // traversal accounting only on request (vkp_set_profiling(ctx, 2))
    if (px.valid) {
        const int W = F.p.width, H = F.p.height;
        const size_t i = (size_t)px.y * W + px.x;
        float3 o, d; camera_ray(F.cam, px.x, px.y, 0.5f, 0.5f, W, H, o, d);
        HitRec h; nrays = 1;
        bool hit = trace_ray<false>(S, sstack, o + d * 1e-2f, d, 0.001f, 10000.0f, h, nullptr, tcp);
        F.hit_ids[i] = make_int2(h.inst, h.prim);
        if (F.seed_w && hit) F.seed_w[i] = tea16((uint32_t)px.y * (uint32_t)W + (uint32_t)px.x, (uint32_t)F.p.frame);      // pixel_seed() of the later stages
        if (!hit) {
            // skybox.frag:16-21
            F.position[i] = make_uint2(0u, 0u); F.normal[i] = make_uint2(0u, 0u);
            F.material[i] = make_uchar4(0, 0, 0, 0); F.emissive[i] = make_uchar4(0, 0, 0, 0);
            F.motion[i] = 0u;
            float2 uv = equirect_uv(normalize3(d), VKP_PI_MISS);
            float4 sky = tex_nearest(S.textures[S.sky_tex], uv.x, uv.y);
            F.albedo[i] = pack_unorm4(sky.x, sky.y, sky.z, sky.w);
        } else {
            const DInstance& I = S.instances[h.inst];
            const uint32_t* ind = I.indices + 3u * (uint32_t)h.prim;
            const vkp_vertex& v0 = I.vertices[__ldg(ind)]; const vkp_vertex& v1 = I.vertices[__ldg(ind + 1)]; const vkp_vertex& v2 = I.vertices[__ldg(ind + 2)];
            const float b0 = 1.0f - h.u - h.v, b1 = h.u, b2 = h.v;
            float3 P0 = xform_point(I.model, f3(v0.position[0], v0.position[1], v0.position[2]));
            float3 P1 = xform_point(I.model, f3(v1.position[0], v1.position[1], v1.position[2]));
            float3 P2 = xform_point(I.model, f3(v2.position[0], v2.position[1], v2.position[2]));
            float3 worldPos = P0 * b0 + P1 * b1 + P2 * b2;
            float3 n = f3(v0.normal[0], v0.normal[1], v0.normal[2]) * b0 + f3(v1.normal[0], v1.normal[1], v1.normal[2]) * b1 +
                       f3(v2.normal[0], v2.normal[1], v2.normal[2]) * b2;
            float3 inNormal = xform_normal_by_inverse(I.inv_model, n);
            float3 inColor = f3(v0.color[0], v0.color[1], v0.color[2]) * b0 + f3(v1.color[0], v1.color[1], v1.color[2]) * b1 +
                             f3(v2.color[0], v2.color[1], v2.color[2]) * b2;
            float uvx = v0.uv[0] * b0 + v1.uv[0] * b1 + v2.uv[0] * b2, uvy = v0.uv[1] * b0 + v1.uv[1] * b1 + v2.uv[1] * b2;
            const vkp_material& mat = S.materials[max(0, min(I.material_id, S.n_materials - 1))];
            float3 color = mat.textures[0] > -1 ? xyz(tex_nearest(S.textures[(int)mat.textures[0]], uvx, uvy)) * inColor : f3(mat.diffuse[0], mat.diffuse[1], mat.diffuse[2]);
            float3 N = normalize3(inNormal);
            if (mat.textures[1] > -1) {
                // geometry_shader.frag:55,60-63 perturbNormal + cotangent_frame :28-49.  dFdx / dFdy of the world position are ray
                // differentials: the camera rays of the right / lower neighbour pixel intersected with the hit triangle's plane.
                // As upstream, the frame uses the un-normalised inNormal and takes duv1 / duv2 from dFdx(p).xy / dFdy(p).xy.
                float3 normal_pixel = xyz(tex_nearest(S.textures[(int)mat.textures[1]], uvx, uvy));
                normal_pixel = normal_pixel * 255.0f / 127.0f - f3s(128.0f / 127.0f);
                const float3 ng = cross3(P1 - P0, P2 - P0);
                float3 ox_, dx_, oy_, dy_;
                camera_ray(F.cam, px.x + 1, px.y, 0.5f, 0.5f, W, H, ox_, dx_);
                camera_ray(F.cam, px.x, px.y + 1, 0.5f, 0.5f, W, H, oy_, dy_);
                const float3 px_ = ox_ + dx_ * (dot3(P0 - ox_, ng) / dot3(dx_, ng));
                const float3 py_ = oy_ + dy_ * (dot3(P0 - oy_, ng) / dot3(dy_, ng));
                const float3 dp1 = px_ - worldPos, dp2 = py_ - worldPos;
                const float3 dp2perp = cross3(dp2, inNormal), dp1perp = cross3(inNormal, dp1);
                const float3 T = dp2perp * dp1.x + dp1perp * dp2.x;
                const float3 Bt = dp2perp * dp1.y + dp1perp * dp2.y;
                const float invmax = 1.0f / sqrtf(fmaxf(dot3(T, T), dot3(Bt, Bt)));
                const float3 Tn = T * invmax, Bn = Bt * invmax;
                N = normalize3((Tn * normal_pixel.x + Bn * normal_pixel.y) + inNormal * normal_pixel.z);
            }
            float3 emissive = mat.textures[2] > -1 ? xyz(tex_nearest(S.textures[(int)mat.textures[2]], uvx, uvy)) : f3s(0.0f);
            float3 material = mat.textures[3] > -1 ? xyz(tex_nearest(S.textures[(int)mat.textures[3]], uvx, uvy))
                                                   : f3(0.0f, mat.shadingMetallicRoughness[2], mat.shadingMetallicRoughness[1]);
            float materialIdx = mat.shadingMetallicRoughness[3];
            F.position[i] = pack_half4(worldPos.x, worldPos.y, worldPos.z, materialIdx);
            float3 nn = N * 0.5f + f3s(0.5f);
            F.normal[i] = pack_half4(nn.x, nn.y, nn.z, 1.0f);
            uchar4 a = pack_unorm4(color.x, color.y, color.z, 1.0f); F.albedo[i] = a;
            F.material[i] = pack_unorm4(material.x, material.y, material.z, 1.0f);
            F.emissive[i] = pack_unorm4(emissive.x, emissive.y, emissive.z, 1.0f);
            if (F.static_camera) {
                // identical matrices give ndc - ndcPrev == 0 exactly (a hit in front of the camera has clip.w != 0): 0.5 in both halves
                F.motion[i] = 0x38003800u;
            } else {
                float4 wp = make_float4(worldPos.x, worldPos.y, worldPos.z, 1.0f);
                float4 clip = xform_vec4(F.cam.proj, xform_vec4(F.cam.view, wp));
                float4 clipPrev = xform_vec4(F.cam.prev_proj, xform_vec4(F.cam.prev_view, wp));
                // basic.vert writes `ndc` to location 5 and `ndcPrev` to location 4, geometry_shader.frag reads location 4 as inNdc and
                // location 5 as inNdcPrev: what the fragment shader calls ndc is the PREVIOUS frame's clip position, so the stored
                // motion is previous - current (seen by executing the two binaries, tests/golden/make_spirv_vectors.py "gbuffer")
                float mx = clipPrev.x / clipPrev.w - clip.x / clip.w;
                float my = clipPrev.y / clipPrev.w - clip.y / clip.w;
                F.motion[i] = (uint32_t)vkp_f32_to_f16(mx * 0.5f + 0.5f) | ((uint32_t)vkp_f32_to_f16(my * 0.5f + 0.5f) << 16);
            }
        }
    }
    if (F.sky_shadow_w && blockIdx.x == 0 && (int)threadIdx.x < S.n_lights) {
        // What shadowRaygen.rgen:28-117 gives for a SKY pixel and light threadIdx.x, if it traces nothing: every sky pixel has the same
        // G-buffer content (all-zero position / normal planes), so whether it faces the light and is within its range is one decision
        // per light and frame -- 57 % of k_shadow's threads then store that byte instead of redoing the arithmetic.  -1 = it does trace.
        const int l = (int)threadIdx.x;
        const float4 pos4 = unpack_half4(make_uint2(0u, 0u)), nrm4 = unpack_half4(make_uint2(0u, 0u));
        const int mode = (int)material_by_float_index(S, pos4.w).shadingMetallicRoughness[0];
        int code;
        if (mode == 4) code = 255;
        else {
            const float3 N = normalize3(xyz(nrm4) * 2.0f - f3s(1.0f));
            const LightGeom g = light_geom(S.lights[l], xyz(pos4), N);
            float shadowFactor = 0.0f;
            code = 0;
            if (g.NdotL > 0.0f) {
                if (g.dist < g.maxd) code = -1;
                else {
                    for (int k = 0; k < F.p.shadow_samples; k++) shadowFactor += 1.0f;
                    shadowFactor /= (float)F.p.shadow_samples;
                }
            }
            if (code == 0) code = (int)vkp_unorm8(shadowFactor);
        }
        F.sky_shadow_w[l] = (short)code;
    }
    if (F.chain_in_gbuffer) {
        // schedule 8: the mirror / glass pixels are compacted into the path lists here (no k_chain_scan pass over the frame).  The test is
        // k_chain_scan's: the material index as the fp16 position plane stores it (0 for a sky pixel), looked up again.
        bool path = false;
        size_t pi = 0;
        if (px.valid) {
            pi = (size_t)px.y * F.p.width + px.x;
            const float matIdx = vkp_f16_to_f32((uint16_t)(F.position[pi].y >> 16));
            path = (int)material_by_float_index(S, matIdx).shadingMetallicRoughness[0] != 0 && F.p.max_bounces > 0;
        }
        chain_enqueue(F, pi, path);
    }
    add_counter(F.counters, 0, nrays);
    if (F.count_trav) add_trav(F.counters, tc, (0u << 28) | (unsigned)(px.valid ? px.y * F.p.width + px.x : 0));
    record_cost(F.cta_cost[0], vb, t_start);
}

// shadowRaygen.rgen:28-117.  One thread per (pixel, light) -- blockIdx.y is the light -- so the longest any-hit
// traversal of a pixel is not serialised behind the other lights' (the kernel's duration is set by its slowest
// thread: DESIGN.md section 4).  With the non-faithful RNG the seed a light starts from is replayed from the
// cheap per-light conditions of the lights before it.
// MINB = resident CTAs per SM the register budget is set for: 8 (64 registers) when a thread traces one ray and waits on it
// (latency-bound), 5 (96 registers) when it loops over many soft-shadow samples (issue-bound: cfg5).
// ONE: the frame traces one sample per light (BASELINE configs[1]): no sample loop, so nothing of the light / G-buffer state
// stays in registers across the traversal.
// The shadow byte of (pixel i, light l), and -- on accumulation frames (blur.comp:41-48: out = mix(history, in, 1 / (frame + 1)), no
// neighbourhood) -- the filter's new history byte with it: the thread that has just produced the filter's only fresh input applies
// the filter, and the frame has one dependent pass less (F.fuse_accumulate; the arithmetic is denoise_pixel's).
__device__ __forceinline__ void store_shadow(const DFrame& F, int l, size_t plane, size_t i, int x, int y, uint8_t value)
{
    F.shadow[l * plane + i] = value;
    if (F.fuse_accumulate) {
        const int Wc = F.p.denoise_cover_remainder ? F.p.width : (F.p.width / 16) * 16;
        const int Hc = F.p.denoise_cover_remainder ? F.p.height : (F.p.height / 16) * 16;
        const uint8_t hist = F.denoised[l * plane + i];
        uint8_t out = hist;
        if (x < Wc && y < Hc) {
            const float center = unit255(value), old = unit255(hist);
            const float a = 1.0f / (float)((float)F.p.denoise_frame + 1.0f);
            out = vkp_unorm8(mix1(old, center, a));
        }
        F.denoised_next[l * plane + i] = out;
    }
}

template <bool ONE>
__device__ __forceinline__ void shadow_patch(const DScene& S, const DFrame& F, uint32_t* __restrict__ sstack, const unsigned int vb, const int warp_override,
                                             unsigned int& nrays, TravCount& tc, TravCount* const tcp, int& l_out, Pixel& px_out)
{
    const int l = (int)fast_div(F, vb, F.magic_n_cta, (unsigned int)F.n_cta);
    const Pixel px = pixel_of_block(F, vb - (unsigned int)l * (unsigned int)F.n_cta, warp_override);
    l_out = l; px_out = px;
    (void)tc;
    if (px.valid) {
        const int W = F.p.width, H = F.p.height;
        const size_t i = (size_t)px.y * W + px.x;
        const size_t plane = (size_t)W * H;
        const uint2 praw = F.position[i], nraw = F.normal[i];
        if (F.sky_shadow_r && (praw.x | praw.y | nraw.x | nraw.y) == 0u) {
            // a sky pixel (all-zero planes): the per-light answer k_gbuffer's block 0 left, unless sky pixels trace towards this light
            const int code = (int)F.sky_shadow_r[l];
            if (code >= 0) { store_shadow(F, l, plane, i, px.x, px.y, (uint8_t)code); return; }
        }
        float4 pos4 = unpack_half4(praw);
        float4 nrm4 = unpack_half4(nraw);
        float3 position = xyz(pos4);
        float3 normal = xyz(nrm4) * 2.0f - f3s(1.0f);
        int mode = (int)material_by_float_index(S, pos4.w).shadingMetallicRoughness[0];
        if (mode == 4) {
            store_shadow(F, l, plane, i, px.x, px.y, 255);
        } else {
            float3 N = normalize3(normal);
            const int shadowSamples = F.p.shadow_samples;
            const bool advance = F.p.faithful_rng == 0;
            const vkp_light light = S.lights[l];
            const LightGeom g = light_geom(light, position, N);
            float shadowFactor = 0.0f;
            if (g.NdotL > 0.0f) {
                if (g.dist < g.maxd) {
                    // the RNG is only needed when rays are traced: TEA (16 rounds) is the most expensive part of an unlit pixel
                    uint32_t seed = pixel_seed(F, i, px.x, px.y);
                    if (advance) {
                        for (int k = 0; k < l; k++) {
                            const LightGeom gk = light_geom(S.lights[k], position, N);
                            if (gk.NdotL > 0.0f && gk.dist < gk.maxd)
                                for (int r = 0; r < 2 * shadowSamples; r++) seed = 1664525u * seed + 1013904223u;   // two rnd() per sampleDisk
                        }
                    }
                    if (ONE) {
                        const float3 dir = sample_disk(light, position, g.L, seed, advance);
                        HitRec h; nrays++;
                        if (!trace_ray<true>(S, sstack, position + dir * 1e-1f, dir, 0.001f, g.dist + 1.0f, h, nullptr, tcp)) shadowFactor += 1.0f;
                    } else
                    for (int s = 0; s < shadowSamples; s++) {
                        const float3 dir = sample_disk(light, position, g.L, seed, advance);
                        HitRec h; nrays++;
                        if (!trace_ray<true, false, VKP_KEEPW_MANY != 0>(S, sstack, position + dir * 1e-1f, dir, 0.001f, g.dist + 1.0f, h, nullptr, tcp)) shadowFactor += 1.0f;
                    }
                } else {
                    for (int s = 0; s < shadowSamples; s++) shadowFactor += 1.0f;      // beyond maxDistance: unshadowed (shadowRaygen.rgen:81-83)
                }
                shadowFactor /= (float)shadowSamples;
            }
            store_shadow(F, l, plane, i, px.x, px.y, vkp_unorm8(shadowFactor));
        }
    }
}

// FOLD: the grid also carries the first-bounce blocks of schedule 8 (an instantiation of its own: the extra code paths cost the plain
// kernel 550 bytes of spills and 0.02 ms when they shared it)
template <int MINB, bool ONE, bool FOLD = false>
__global__ void __launch_bounds__(VKP_TRAV_BLOCK, MINB) k_shadow(const __grid_constant__ DScene S, const __grid_constant__ DFrame F)
{
    __shared__ uint32_t s_trav[VKP_TRAV_SMEM_WORDS];
    uint32_t* const sstack = s_trav + threadIdx.x;
    const long long t_start = clock64();
    unsigned int nrays = 0;
    TravCount tc = {0u, 0u, 0u};
    TravCount* const tcp = F.count_trav ? &tc : nullptr;      // traversal accounting only on request (vkp_set_profiling(ctx, 2))
    if (FOLD && blockIdx.x < F.first_blocks) {
        // schedule 8: the leading blocks of the grid trace the FIRST reflection / refraction ray of the queued paths (k_chain_first's work)
        // inside this kernel: the long rays among them start at t = 0 and their tail hides under 48 000 blocks of shadow rays instead
        // of ending a kernel of its own
        unsigned int n_refl = 0;
        chain_first_paths(S, F, sstack, blockIdx.x * blockDim.x + threadIdx.x, F.first_blocks * blockDim.x, n_refl, tcp);
        add_counter(F.counters, 2, n_refl);
        if (F.count_trav) add_trav(F.counters, tc);
        return;
    }
    unsigned int hb = FOLD ? blockIdx.x - F.first_blocks : blockIdx.x;
    if (FOLD && F.first_blocks) {
        // ... and further down the grid, first_blocks x n_lights blocks trace the per-light any-hit rays of those first hits
        // (k_chain_shadow0's work).  They are dispatched long after the leading blocks, whose records they read -- each carries a tag
        // written with st.release when it is complete -- so the wait below practically never spins, and it cannot deadlock: a block
        // only ever waits for blocks dispatched before it.
        const unsigned int mid = F.first_blocks * (unsigned int)S.n_lights;
        if (hb >= F.mid_at && hb < F.mid_at + mid) {
            const unsigned int unit = hb - F.mid_at;
            const int l = (int)(unit / F.first_blocks);
            const unsigned int kb = unit - (unsigned int)l * F.first_blocks;
            unsigned int n_sec = 0;
            chain_shadow0_paths(S, F, sstack, l, kb * blockDim.x + threadIdx.x, F.first_blocks * blockDim.x, n_sec, tcp, true);
            add_counter(F.counters, 3, n_sec);
            if (F.count_trav) add_trav(F.counters, tc);
            return;
        }
        if (hb >= F.mid_at) hb -= mid;
    }
    const unsigned int vb = F.cta_order[1] ? __ldg(F.cta_order[1] + hb) : hb;      // = light * n_cta + pixel block
    int l; Pixel px;
    shadow_patch<ONE>(S, F, sstack, vb, -1, nrays, tc, tcp, l, px);
    add_counter(F.counters, 1, nrays);
    if (F.count_trav) add_trav(F.counters, tc, (1u << 28) | ((unsigned)l << 23) | (unsigned)(px.valid ? px.y * F.p.width + px.x : 0));
    record_cost(F.cta_cost[1], vb, t_start);
}

// The same pass as a PERSISTENT grid (north_star: "persistent-thread ... traversal"; A/B knob VKP_SHADOW_PERSIST = CTAs per SM): a few
// resident CTAs per SM whose WARPS take the next 8 x 4 pixel patch of the next (light, block) from a counter until none is left, in the
// longest-first order of the block schedule.  The hardware block scheduler is replaced by one atomic per warp and 32 pixels, and a warp
// that drew a cheap patch is back for the next one without waiting for the other three warps of its CTA.
template <int MINB, bool ONE>
__global__ void __launch_bounds__(VKP_TRAV_BLOCK, MINB) k_shadow_persistent(const __grid_constant__ DScene S, const __grid_constant__ DFrame F, unsigned int* __restrict__ next,
                                                                            const unsigned int n_units)
{
    __shared__ uint32_t s_trav[VKP_TRAV_SMEM_WORDS];
    uint32_t* const sstack = s_trav + threadIdx.x;
    const int lane = (int)(threadIdx.x & 31u);
    unsigned int nrays = 0;
    TravCount tc = {0u, 0u, 0u};
    TravCount* const tcp = F.count_trav ? &tc : nullptr;
    while (true) {
        unsigned int u = 0u;
        if (lane == 0) u = atomicAdd(next, 1u);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= n_units) break;
        const long long t0 = clock64();
        const unsigned int b = u >> 2;
        const unsigned int vb = F.cta_order[1] ? __ldg(F.cta_order[1] + b) : b;
        int l; Pixel px;
        shadow_patch<ONE>(S, F, sstack, vb, (int)(u & 3u), nrays, tc, tcp, l, px);
        __syncwarp();
        record_cost(F.cta_cost[1], vb, t0);
    }
    add_counter(F.counters, 1, nrays);
    if (F.count_trav) add_trav(F.counters, tc, 1u << 28);
}

#include "kernels_shadow_variants.cuh"      // k_shadow_classify / k_shadow_items / k_shadow_stream (off by default)

// vkp_params.denoiser == 1 (SURVEY 8f-3 "a better denoiser"; no reference counterpart, restated by the oracle's
// atrous_iteration): one iteration of an edge-aware a-trous filter over the R8 shadow planes, guided by the G-buffer.  5 x 5 taps `step`
// pixels apart, B3-spline weights, a normal weight max(0, N_p . N_q)^32 and a plane-distance weight max(0, 1 - |N_p . (P_q - P_p)| /
// (0.1 step)); background pixels pass through and are never taps.  Lights are filtered four at a time (the weights do not depend on
// the light); the accumulation order -- rows, then columns -- is the oracle's.
__global__ void __launch_bounds__(128) k_atrous(const __grid_constant__ DScene S, const __grid_constant__ DFrame F, const uint8_t* __restrict__ in,
                                                uint8_t* __restrict__ out, int step)
{
    const Pixel px = pixel_of_thread(F);
    if (!px.valid) return;
    const int W = F.p.width, H = F.p.height;
    const size_t plane = (size_t)W * H, i = (size_t)px.y * W + px.x;
    const float4 np4 = unpack_half4(F.normal[i]);
    if (np4.w == 0.0f) { for (int l = 0; l < S.n_lights; l++) out[l * plane + i] = in[l * plane + i]; return; }
    const float3 Np = normalize3(xyz(np4) * 2.0f - f3s(1.0f));
    const float3 Pp = xyz(unpack_half4(F.position[i]));
    const float inv_sigma = 1.0f / (0.1f * (float)step);
    for (int l0 = 0; l0 < S.n_lights; l0 += 4) {
        float sum[4] = {0.0f, 0.0f, 0.0f, 0.0f}, wsum = 0.0f;
        for (int dy = -2; dy <= 2; dy++)
            for (int dx = -2; dx <= 2; dx++) {
                const int qx = px.x + dx * step, qy = px.y + dy * step;
                if (qx < 0 || qy < 0 || qx >= W || qy >= H) continue;
                const size_t q = (size_t)qy * W + qx;
                const float4 nq4 = unpack_half4(F.normal[q]);
                if (nq4.w == 0.0f) continue;
                const float3 Nq = normalize3(xyz(nq4) * 2.0f - f3s(1.0f));
                const float3 Pq = xyz(unpack_half4(F.position[q]));
                float wn = fmaxf(dot3(Np, Nq), 0.0f);
                wn = wn * wn; wn = wn * wn; wn = wn * wn; wn = wn * wn; wn = wn * wn;
                const float dist = fabsf(dot3(Np, Pq - Pp));
                const float wp = fmaxf(1.0f - dist * inv_sigma, 0.0f);
                const int ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy;
                const float kx = ax == 0 ? 0.375f : ax == 1 ? 0.25f : 0.0625f, ky = ay == 0 ? 0.375f : ay == 1 ? 0.25f : 0.0625f;
                const float w = (kx * ky) * wn * wp;
                wsum = wsum + w;
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (l0 + k < S.n_lights) sum[k] = sum[k] + w * unit255(in[(l0 + k) * plane + q]);
            }
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (l0 + k < S.n_lights) out[(l0 + k) * plane + i] = vkp_unorm8(sum[k] / wsum);
    }
}

__global__ void __launch_bounds__(128) k_denoise(const __grid_constant__ DScene S, const __grid_constant__ DFrame F)
{
    const Pixel px = pixel_of_thread(F);
    if (!px.valid) return;
    const size_t i = (size_t)px.y * F.p.width + px.x;
    const size_t plane = (size_t)F.p.width * F.p.height;
    for (int l = 0; l < S.n_lights; l++) F.denoised_next[l * plane + i] = denoise_pixel(F, l, px.x, px.y, i, plane);
}

// blur.comp:41-48, the accumulation path of every frame after a reset: out = mix(out_prev, in, 1 / (frame + 1)) -- pure byte
// streaming (2L bytes in, L out per pixel).  One thread = 4 adjacent pixels of a row (32-bit loads / stores), one warp = one
// 16 x 8 pixel block, same tile enumeration as the other frame kernels.  Needs W % 4 == 0 (rows stay 4-byte aligned).
__global__ void __launch_bounds__(32) k_denoise_accumulate4(const __grid_constant__ DScene S, const __grid_constant__ DFrame F)
{
    const int W = F.p.width, H = F.p.height;
    const int bpt_x = F.tile_w >> 4, bpt = bpt_x * (F.tile_h >> 3);
    const int own_tile = (int)blockIdx.x / bpt, sub = (int)blockIdx.x - own_tile * bpt;
    const int tile = F.rank + own_tile * F.world;
    const int tx = tile % F.tiles_x, ty = tile / F.tiles_x;
    const int x = tx * F.tile_w + (sub % bpt_x) * 16 + (threadIdx.x & 3) * 4;
    const int y = ty * F.tile_h + (sub / bpt_x) * 8 + (threadIdx.x >> 2);
    if (x >= W || y >= H) return;
    const int Wc = F.p.denoise_cover_remainder ? W : (W / 16) * 16;
    const int Hc = F.p.denoise_cover_remainder ? H : (H / 16) * 16;
    const size_t plane = (size_t)W * H, i = (size_t)y * W + x;
    const bool covered = x < Wc && y < Hc;       // a group of 4 is entirely inside or outside (Wc, x are multiples of 4)
    const float a = 1.0f / (float)((float)F.p.denoise_frame + 1.0f);
    for (int l = 0; l < S.n_lights; l++) {
        const uint32_t h4 = __ldg(reinterpret_cast<const uint32_t*>(F.denoised + l * plane + i));
        uint32_t o4 = h4;
        if (covered) {
            const uint32_t c4 = __ldg(reinterpret_cast<const uint32_t*>(F.shadow + l * plane + i));
            o4 = 0u;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float center = unit255((unsigned char)(c4 >> (8 * k))), old = unit255((unsigned char)(h4 >> (8 * k)));
                o4 |= (uint32_t)vkp_unorm8(mix1(old, center, a)) << (8 * k);
            }
        }
        *reinterpret_cast<uint32_t*>(F.denoised_next + l * plane + i) = o4;
    }
}

#include "kernels_chain.cuh"                // the reflection chain in all its schedules

// CHAIN == 0: the bounce chain runs inline in this kernel (its latency hides behind the other pixels' shading);
// CHAIN == 1: mirror / glass pixels are compacted into a path queue consumed by k_bounce, one launch per bounce;
// CHAIN == 2: the chain was traced by k_chain concurrently with the shadow pass; multiply its stored colours in.
// CHAIN == 3: as 2, but this kernel does not wait for the chain: a mirror / glass pixel leaves its direct term and its ambient tail in
//             the path's record and k_resolve_paths finishes it once the chain is there -- the chain's longest path then overlaps
//             the shadow pass, the filter AND the resolve of the other 92 % of the pixels.
template <int CHAIN>
__global__ void __launch_bounds__(VKP_TRAV_BLOCK, VKP_LB_RESOLVE) k_resolve(const __grid_constant__ DScene S, const __grid_constant__ DFrame F)
{
    // (only the inline chain, CHAIN == 0, traces rays here; the other instantiations drop the array)
    __shared__ uint32_t s_trav[CHAIN == 0 ? VKP_TRAV_SMEM_WORDS : 1];
    uint32_t* const sstack = s_trav + (CHAIN == 0 ? threadIdx.x : 0u);
    const Pixel px = pixel_of_thread(F);
    bool wants_path = false;
    unsigned int n_refl = 0, n_sec = 0;
    TravCount tc = {0u, 0u, 0u};
    TravCount* const tcp = F.count_trav ? &tc : nullptr;      // traversal accounting only on request (vkp_set_profiling(ctx, 2))
    PathItem item;
    if (px.valid) {
        const int W = F.p.width, H = F.p.height;
        const size_t i = (size_t)px.y * W + px.x;
        const size_t plane = (size_t)W * H;
        Payload prd;
        prd.seed = pixel_seed(F, i, px.x, px.y);
        float4 pos4 = unpack_half4(F.position[i]);
        float4 nrm4 = unpack_half4(F.normal[i]);
        const vkp_material& mat = material_by_float_index(S, pos4.w);
        const int shadingMode = (int)mat.shadingMetallicRoughness[0];
        float3 position = xyz(pos4);
        float3 normal = xyz(nrm4) * 2.0f - f3s(1.0f);
        float3 albedo = lut3(F.lut_pow22, F.albedo[i]);
        float3 material = unorm3(F.material[i]);
        float3 emissive = unorm3(F.emissive[i]);
        const bool background = pos4.w == 0.0f && nrm4.w == 0.0f;
        if (background && shadingMode == 0 && !F.fused) {
            // a sky pixel (material 0 diffuse): no light loop, no path -- the pixel ends as its albedo, the sky texel (hybridRaygen.rgen:182-183);
            // none of the view / irradiance / Fresnel terms below can reach the result (57 % of the 1080p default frame is sky)
            const float4 outc = make_float4(albedo.x, albedo.y, albedo.z, 1.0f);
            F.final_f32[i] = outc;
            F.final_srgb8[i] = post_pixel(F.lut_post, outc);
        } else {
        const float roughness = material.y, metallic = material.z;
        const float3 F0 = mix3(f3s(0.04f), albedo, metallic);
        const float3 camPosition = f3(F.cam.inv_view[12], F.cam.inv_view[13], F.cam.inv_view[14]);
        const float3 V = normalize3(camPosition - position);
        const float3 N = normalize3(normal);
        const float NdotV = clamp01(dot3(N, V));
        const float2 envUV = equirect_uv(N, VKP_PI);
        const float3 irradiance = xyz(tex_nearest(S.textures[S.irr_tex], envUV.x, envUV.y));

        float3 origin = f3s(0.0f), direction = f3s(0.0f);
        float attenuation = 1.0f;   // declared outside the light loop upstream: stale across lights (App. A #8)
        float3 rayColor = f3s(0.0f);
        // a background pixel with a diffuse material 0 traces nothing and ends as `albedo` (hybridRaygen.rgen:182-183): its light loop is dead code
        const int n_lights_eff = (background && shadingMode == 0) ? 0 : S.n_lights;
        // F.fused: the shadow filter (blur.comp) runs here, per light, instead of in its own pass over the frame; the new
        // history goes to the other ping-pong buffer, so reprojected reads of the old history by other pixels stay valid
        for (int l = 0; l < (F.fused ? S.n_lights : n_lights_eff); l++) {
            uint8_t dq;
            if (F.fused) { dq = denoise_pixel(F, l, px.x, px.y, i, plane); F.denoised_next[l * plane + i] = dq; }
            else dq = F.denoised[l * plane + i];
            if (l >= n_lights_eff) continue;
            const vkp_light light = S.lights[l];
            const float3 lpos = f3(light.position[0], light.position[1], light.position[2]);
            const float3 lcol = f3(light.color[0], light.color[1], light.color[2]);
            const bool isDirectional = light.position[3] < 0.0f;
            float3 L = isDirectional ? lpos : (lpos - position);
            const float light_max_distance = light.position[3];
            const float light_distance = length3(L);
            const float light_intensity = isDirectional ? 1.0f : (light.color[3] / (light_distance * light_distance));
            L = normalize3(L);
            const float NdotL = clamp01(dot3(N, L));
            const float shadowFactor = unit255(dq);
            if (NdotL > 0.0f) {
                if (light_intensity == 0.0f) attenuation = 0.0f;
                else {
                    attenuation = light_max_distance - light_distance;
                    attenuation /= light_max_distance;
                    attenuation = fmaxf(attenuation, 0.0f);
                    attenuation = isDirectional ? 0.3f : (attenuation * attenuation);
                }
            }
            if (shadingMode == 0) {
                float3 radiance = lcol * light_intensity * attenuation * shadowFactor;
                const float3 Hh = normalize3(V + L);
                float D = distribution_ggx(N, Hh, roughness);
                float G = geometry_smith(N, V, L, roughness);
                float3 Fr = fresnel_schlick(fmaxf(dot3(Hh, V), 0.0f), F0);
                float3 kD = f3s(1.0f) - Fr;
                kD = kD * (1.0f - metallic);
                float3 numerator = (D * G) * Fr;
                float denominator = 4.0f * NdotV * NdotL;
                float3 specular = numerator / fmaxf(denominator, 0.001f);
                rayColor = rayColor + (kD * albedo / VKP_PI + specular) * radiance * NdotL;
            }
            if (shadingMode == 3) {
                direction = reflect3(normalize3(-V), N);
                origin = position;
                rayColor = rayColor + ((NdotL > 0.0f && light_intensity > 0.0f) ? lcol * light_intensity * attenuation * shadowFactor * albedo * metallic
                                                                                : irradiance * albedo * metallic);
            }
            if (shadingMode == 4) {
                const float ior = mat.diffuse[3];
                origin = position;
                const float cosAlpha = dot3(N, V);
                const float3 Iv = -V;
                const float NdotI = dot3(N, Iv);
                const float3 refrNormal = NdotI > 0.0f ? -N : N;
                const float refrEta = NdotI > 0.0f ? 1.0f / ior : ior;
                const float radicand = 1.0f + vkp_powf(refrEta, 2.0f) * (cosAlpha * cosAlpha - 1.0f);
                direction = radicand < 0.0f ? reflect3(Iv, N) : refract3(Iv, refrNormal, refrEta);
                rayColor = rayColor + ((light_intensity > 0.0f) ? light_intensity * lcol * attenuation * albedo * metallic
                                                                : irradiance * albedo * metallic);
            }
        }
        // hybridRaygen.rgen:172-183: ambient + emissive tail; for background pixels the result is the sky albedo
        float3 Fa = fresnel_schlick(NdotV, F0);
        float3 kDa = (1.0f - Fa) * (1.0f - metallic);
        float3 diffuse = kDa * albedo * irradiance;
        float3 tail = diffuse + emissive;
        wants_path = shadingMode != 0 && F.p.max_bounces > 0;
        if (CHAIN == 2 && wants_path) {
            const int slot = F.chain_slot[i];
            const float4* rec = F.chain_buf + (size_t)slot * VKP_CHAIN_MAX;
            const int n = F.chain_len[slot];
            for (int k = 0; k < n; k++) { float4 c = rec[k]; rayColor = rayColor * f3(c.x, c.y, c.z); }
        }
        if (CHAIN == 0 && wants_path) {
            // hybridRaygen.rgen:153-170
            for (int depth = 0; depth < F.p.max_bounces; depth++) {
                float3 ro = origin + direction * 1e-2f;
                HitRec h; n_refl++;
                bool hit = trace_ray<false>(S, sstack, ro, direction, 0.001f, 1000.0f, h, nullptr, tcp);
                if (hit) shade_hit<false>(S, sstack, F.p, ro, direction, h, prd, n_sec, tcp); else shade_miss(S, direction, prd);
                rayColor = rayColor * prd.color;
                if (!hit || !(prd.scatter > 0.0f)) break;
                origin = prd.origin; direction = prd.direction;
            }
        }
        if (CHAIN == 3 && wants_path) {
            PathItem* rec = F.chain_base + F.chain_slot[i];
            rec->c = make_float4(rayColor.x, rayColor.y, rayColor.z, background ? 1.0f : 0.0f);
            rec->d = background ? make_float4(albedo.x, albedo.y, albedo.z, 0.0f) : make_float4(tail.x, tail.y, tail.z, 0.0f);
        } else if (CHAIN != 1 || !wants_path) {
            float3 finalColor = rayColor + tail;
            if (background) finalColor = albedo;
            float4 outc = make_float4(finalColor.x, finalColor.y, finalColor.z, 1.0f);
            F.final_f32[i] = outc;
            F.final_srgb8[i] = post_pixel(F.lut_post, outc);
        } else {
            // hybridRaygen.rgen:153-170 becomes the wavefront below: queue the first reflection / refraction ray
            item.a = make_float4(origin.x, origin.y, origin.z, __uint_as_float((uint32_t)i | (background ? 0x80000000u : 0u)));
            item.b = make_float4(direction.x, direction.y, direction.z, __uint_as_float(prd.seed));
            item.c = make_float4(rayColor.x, rayColor.y, rayColor.z, 0.0f);
            item.d = background ? make_float4(albedo.x, albedo.y, albedo.z, 0.0f) : make_float4(tail.x, tail.y, tail.z, 0.0f);
        }
        }
    }
    if (CHAIN == 1) path_push(F.path_q[0], &F.path_count[0], wants_path, item);
    if (CHAIN == 0) { add_counter(F.counters, 2, n_refl); add_counter(F.counters, 3, n_sec); }
}

// One bounce of the reflection / refraction chain of hybridRaygen.rgen:153-170 for every queued path: closest-hit
// traversal, hit shading (hybridHit.rchit, with its per-light secondary shadow rays) or sky (miss.rmiss), throughput
// update, then either re-queue (compacted with warp ballot + prefix) or finish the pixel.
__global__ void __launch_bounds__(VKP_TRAV_BLOCK) k_bounce(const __grid_constant__ DScene S, const __grid_constant__ DFrame F, int depth)
{
    __shared__ uint32_t s_trav[VKP_TRAV_SMEM_WORDS];
    uint32_t* const sstack = s_trav + threadIdx.x;
    const uint32_t n_in = F.path_count[depth];
    const PathItem* __restrict__ qin = F.path_q[depth & 1];
    PathItem* qout = F.path_q[(depth + 1) & 1];
    unsigned int n_refl = 0, n_sec = 0;
    TravCount tc = {0u, 0u, 0u};
    TravCount* const tcp = F.count_trav ? &tc : nullptr;      // traversal accounting only on request (vkp_set_profiling(ctx, 2))
    const uint32_t n_round = (n_in + 31u) & ~31u;       // whole warps stay in the loop so that the ballots are convergent
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_round; k += gridDim.x * blockDim.x) {
        bool again = false;
        PathItem item;
        if (k < n_in) {
            item = qin[k];
            const uint32_t tag = __float_as_uint(item.a.w);
            const uint32_t i = tag & 0x7fffffffu;
            const bool background = (tag & 0x80000000u) != 0;
            float3 origin = f3(item.a.x, item.a.y, item.a.z), direction = f3(item.b.x, item.b.y, item.b.z);
            float3 rayColor = f3(item.c.x, item.c.y, item.c.z);
            Payload prd; prd.seed = __float_as_uint(item.b.w);
            float3 ro = origin + direction * 1e-2f;
            HitRec h; n_refl++;
            bool hit = trace_ray<false>(S, sstack, ro, direction, 0.001f, 1000.0f, h, nullptr, tcp);
            if (hit) shade_hit<false>(S, sstack, F.p, ro, direction, h, prd, n_sec, tcp); else shade_miss(S, direction, prd);
            rayColor = rayColor * prd.color;
            again = hit && (prd.scatter > 0.0f) && (depth + 1 < F.p.max_bounces);
            if (again) {
                item.a = make_float4(prd.origin.x, prd.origin.y, prd.origin.z, item.a.w);
                item.b = make_float4(prd.direction.x, prd.direction.y, prd.direction.z, __uint_as_float(prd.seed));
                item.c = make_float4(rayColor.x, rayColor.y, rayColor.z, 0.0f);
            } else {
                float3 tail = f3(item.d.x, item.d.y, item.d.z);
                float3 finalColor = background ? tail : rayColor + tail;
                float4 outc = make_float4(finalColor.x, finalColor.y, finalColor.z, 1.0f);
                F.final_f32[i] = outc;
                F.final_srgb8[i] = post_pixel(F.lut_post, outc);
            }
        }
        path_push(qout, &F.path_count[depth + 1], again, item);
    }
    add_counter(F.counters, 2, n_refl);
    add_counter(F.counters, 3, n_sec);
    if (F.count_trav) add_trav(F.counters, tc);
}

// the end of hybridRaygen.rgen:153-183 for the mirror / glass pixels of k_resolve<3>: throughput = direct term x the chain's colours in
// bounce order, + ambient tail (or the sky albedo); one thread per path slot
__global__ void __launch_bounds__(128) k_resolve_paths(const __grid_constant__ DFrame F)
{
    const unsigned int n_low = F.path_count[0], n_paths = n_low + (F.chain_split ? F.path_count[4] : 0u);
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_paths; k += gridDim.x * blockDim.x) {
        const unsigned int slot = k < n_low ? k : F.chain_cap - 1u - (k - n_low);
        const size_t i = (size_t)F.chain_pix[slot];
        const float4 c0 = F.chain_base[slot].c, t0 = F.chain_base[slot].d;
        float3 rayColor = f3(c0.x, c0.y, c0.z);
        const float4* rec = F.chain_buf + (size_t)slot * VKP_CHAIN_MAX;
        const int n = F.chain_len[slot];
        for (int k = 0; k < n; k++) { float4 c = rec[k]; rayColor = rayColor * f3(c.x, c.y, c.z); }
        const float3 tail = f3(t0.x, t0.y, t0.z);
        float3 finalColor = rayColor + tail;
        if (c0.w != 0.0f) finalColor = tail;                    // background: the sky albedo
        float4 outc = make_float4(finalColor.x, finalColor.y, finalColor.z, 1.0f);
        F.final_f32[i] = outc;
        F.final_srgb8[i] = post_pixel(F.lut_post, outc);
    }
}

__global__ void __launch_bounds__(128) k_post(const __grid_constant__ DFrame F)
{
    const Pixel px = pixel_of_thread(F);
    if (!px.valid) return;
    const size_t i = (size_t)px.y * F.p.width + px.x;
    F.final_srgb8[i] = post_pixel(F.lut_post, F.final_f32[i]);
}

// raygen.rgen:20-81 with NSAMPLES = params.spp; the accumulation target is kept in fp32
__global__ void __launch_bounds__(VKP_TRAV_BLOCK, VKP_LB_RT) k_raytrace(const __grid_constant__ DScene S, const __grid_constant__ DFrame F)
{
    __shared__ uint32_t s_trav[VKP_TRAV_SMEM_WORDS];
    uint32_t* const sstack = s_trav + threadIdx.x;
    const Pixel px = pixel_of_thread(F);
    unsigned int n_refl = 0, n_sec = 0;
    TravCount tc = {0u, 0u, 0u};
    TravCount* const tcp = F.count_trav ? &tc : nullptr;      // traversal accounting only on request (vkp_set_profiling(ctx, 2))
    if (px.valid) {
        const int W = F.p.width, H = F.p.height;
        const size_t i = (size_t)px.y * W + px.x;
        const int frame = F.p.frame;
        Payload prd;
        prd.seed = tea16((uint32_t)px.y * (uint32_t)W + (uint32_t)px.x, (uint32_t)frame);
        float3 finalColor = f3s(0.0f);
        for (int nS = 0; nS < F.p.spp; nS++) {
            float r1 = rnd01(prd.seed);
            float r2 = rnd01(prd.seed);
            float ox = nS == 0 ? 0.5f : r1, oy = nS == 0 ? 0.5f : r2;
            float3 origin, direction; camera_ray(F.cam, px.x, px.y, ox, oy, W, H, origin, direction);
            float3 rayColor = f3s(1.0f);
            for (int b = 0; b < F.p.max_bounces; b++) {
                float3 ro = origin + direction * 1e-2f;
                HitRec h; n_refl++;
                bool hit = trace_ray<false>(S, sstack, ro, direction, 0.001f, 10000.0f, h, nullptr, tcp);
                if (hit) shade_hit<true>(S, sstack, F.p, ro, direction, h, prd, n_sec, tcp); else shade_miss(S, direction, prd);
                rayColor = rayColor * prd.color;
                if (!hit || !(prd.scatter > 0.0f)) break;
                origin = prd.origin; direction = prd.direction;
            }
            finalColor = finalColor + rayColor;
        }
        finalColor = finalColor / (float)F.p.spp;
        if (frame >= 0) {
            float a = 1.0f / (float)(1.0f + (float)frame);
            float4 last = F.final_f32[i];
            finalColor = mix3(f3(last.x, last.y, last.z), finalColor, a);
        }
        float4 outc = make_float4(finalColor.x, finalColor.y, finalColor.z, 1.0f);
        F.final_f32[i] = outc;
        F.final_srgb8[i] = post_pixel(F.lut_post, outc);
    }
    add_counter(F.counters, 2, n_refl);
    add_counter(F.counters, 3, n_sec);
    if (F.count_trav) add_trav(F.counters, tc);
}

// raygen.rgen with ONE WARP PER PIXEL, for many shadow samples per hit (BASELINE configs[2]: 64 spp x 64 samples): a pixel
// is ~12 000 rays, so one thread per pixel leaves a tile-sharded GPU with less than one wave of threads and a ~100 ms
// critical path per pixel.  Here the lanes share the path and split each light's shadow samples (shade_hit<true, true>).
__global__ void __launch_bounds__(VKP_TRAV_BLOCK, VKP_LB_RT) k_raytrace_warp(const __grid_constant__ DScene S, const __grid_constant__ DFrame F)
{
    __shared__ uint32_t s_trav[VKP_TRAV_SMEM_WORDS];
    uint32_t* const sstack = s_trav + threadIdx.x;
    const int tpp = F.tile_w * F.tile_h;
    const long long gw = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int own_tile = (int)(gw / tpp), within = (int)(gw - (long long)own_tile * tpp);
    const int tile = F.rank + own_tile * F.world;
    const int tx = tile % F.tiles_x, ty = tile / F.tiles_x;
    const int x = tx * F.tile_w + within % F.tile_w, y = ty * F.tile_h + within / F.tile_w;
    unsigned int n_refl = 0, n_sec = 0;
    TravCount tc = {0u, 0u, 0u};
    TravCount* const tcp = F.count_trav ? &tc : nullptr;      // traversal accounting only on request (vkp_set_profiling(ctx, 2))
    if (ty < F.tiles_y && x < F.p.width && y < F.p.height) {           // warp-uniform
        const int W = F.p.width, H = F.p.height;
        const size_t i = (size_t)y * W + x;
        const int frame = F.p.frame;
        Payload prd;
        prd.seed = tea16((uint32_t)y * (uint32_t)W + (uint32_t)x, (uint32_t)frame);
        float3 finalColor = f3s(0.0f);
        for (int nS = 0; nS < F.p.spp; nS++) {
            float r1 = rnd01(prd.seed);
            float r2 = rnd01(prd.seed);
            float ox = nS == 0 ? 0.5f : r1, oy = nS == 0 ? 0.5f : r2;
            float3 origin, direction; camera_ray(F.cam, x, y, ox, oy, W, H, origin, direction);
            float3 rayColor = f3s(1.0f);
            for (int b = 0; b < F.p.max_bounces; b++) {
                float3 ro = origin + direction * 1e-2f;
                HitRec h; n_refl++;
                bool hit = trace_ray<false>(S, sstack, ro, direction, 0.001f, 10000.0f, h, nullptr, tcp);
                if (hit) shade_hit<true, 1>(S, sstack, F.p, ro, direction, h, prd, n_sec, tcp); else shade_miss(S, direction, prd);
                rayColor = rayColor * prd.color;
                if (!hit || !(prd.scatter > 0.0f)) break;
                origin = prd.origin; direction = prd.direction;
            }
            finalColor = finalColor + rayColor;
        }
        finalColor = finalColor / (float)F.p.spp;
        if (frame >= 0) {
            float a = 1.0f / (float)(1.0f + (float)frame);
            float4 last = F.final_f32[i];
            finalColor = mix3(f3(last.x, last.y, last.z), finalColor, a);
        }
        if ((threadIdx.x & 31) == 0) {
            float4 outc = make_float4(finalColor.x, finalColor.y, finalColor.z, 1.0f);
            F.final_f32[i] = outc;
            F.final_srgb8[i] = post_pixel(F.lut_post, outc);
        } else {
            n_refl = 0;                       // the closest-hit rays are the warp's, count them once
        }
    }
    add_counter(F.counters, 2, n_refl);
    add_counter(F.counters, 3, n_sec);
    if (F.count_trav) add_trav(F.counters, tc);
}

// deferred.frag:30-133 (no gamma pass in DEFERRED mode upstream; the sRGB8 plane gets the plain 8-bit store)
__global__ void __launch_bounds__(128) k_deferred(const __grid_constant__ DScene S, const __grid_constant__ DFrame F)
{
    const Pixel px = pixel_of_thread(F);
    if (!px.valid) return;
    const int W = F.p.width;
    const size_t i = (size_t)px.y * W + px.x;
    float4 pos4 = unpack_half4(F.position[i]);
    float4 nrm4 = unpack_half4(F.normal[i]);
    float3 position = xyz(pos4);
    float3 normal = xyz(nrm4) * 2.0f - f3s(1.0f);
    float3 albedo = unorm3(F.albedo[i]);
    uint32_t m = F.motion[i];
    float3 motion = f3(vkp_f16_to_f32((uint16_t)(m & 0xffffu)) * 2.0f - 1.0f, vkp_f16_to_f32((uint16_t)(m >> 16)) * 2.0f - 1.0f, 0.0f * 2.0f - 1.0f);
    float3 material = unorm3(F.material[i]);
    float3 emissive = unorm3(F.emissive[i]);
    const bool background = pos4.w == 0.0f && nrm4.w == 0.0f;
    const float metallic = material.z, roughness = material.y;
    float3 N = normalize3(normal);
    float3 V = normalize3(f3(F.cam.position[0], F.cam.position[1], F.cam.position[2]) - position);
    float NdotV = fmaxf(dot3(N, V), 0.0f);
    const float3 albedo_lin = lut3(F.lut_pow22, F.albedo[i]);
    float3 F0 = mix3(f3s(0.04f), albedo_lin, metallic);
    float2 envUV = equirect_uv(N, VKP_PI);
    float3 irradiance = xyz(tex_nearest(S.textures[S.irr_tex], envUV.x, envUV.y));
    float3 color = f3s(1.0f);
    float alpha = 1.0f;
    if (F.p.debug_target > 0) {
        switch (F.p.debug_target) {
        case 1: color = position; alpha = 1.0f; break;
        case 2: color = normal; alpha = 0.0f; break;
        case 3: color = albedo; break;
        case 4: color = motion; break;
        case 5: color = material; break;
        case 6: color = emissive; break;
        default: color = f3s(0.0f); alpha = 0.0f; break;
        }
    } else {
        float3 Lo = f3s(0.0f);
        for (int l = 0; l < S.n_lights; l++) {
            const vkp_light light = S.lights[l];
            const float3 lpos = f3(light.position[0], light.position[1], light.position[2]);
            const float3 lcol = f3(light.color[0], light.color[1], light.color[2]);
            const bool isDirectional = light.position[3] < 0.0f;
            float3 L = isDirectional ? lpos : (lpos - position);
            float3 Hh = normalize3(V + normalize3(L));
            float NdotL = fmaxf(dot3(N, normalize3(L)), 0.0f);
            if (isDirectional) {
                Lo = Lo + (NdotL * lcol);
            } else {
                float light_max_distance = light.position[3];
                float light_distance = length3(L);
                float light_intensity = light.color[3] / (light_distance * light_distance);
                float attenuation = light_max_distance - light_distance;
                attenuation /= light_max_distance;
                attenuation = fmaxf(attenuation, 0.0f);
                attenuation = attenuation * attenuation;
                float3 radiance = lcol * light_intensity * attenuation;
                float NDF = distribution_ggx(N, Hh, roughness);
                float G = geometry_smith(N, V, L, roughness);     // un-normalised L (deferred.frag:104)
                float3 Fr = fresnel_schlick(fmaxf(dot3(Hh, V), 0.0f), F0);
                float3 kD = f3s(1.0f) - Fr;
                kD = kD * (1.0f - metallic);
                float3 numerator = (NDF * G) * Fr;
                float denominator = 4.0f * NdotV * fmaxf(dot3(N, L), 0.0f);
                float3 specular = numerator / fmaxf(denominator, 0.001f);
                Lo = Lo + (kD * albedo_lin / VKP_PI + specular) * radiance * NdotL;
            }
        }
        if (!background) {
            float3 Fa = fresnel_schlick(NdotV, F0);
            float3 kD = (1.0f - Fa) * (1.0f - metallic);
            float3 diffuse = kD * albedo * irradiance;
            color = Lo + diffuse;
            color = color + emissive;
        } else {
            color = albedo;
        }
    }
    float4 outc = make_float4(color.x, color.y, color.z, alpha);
    F.final_f32[i] = outc;
    F.final_srgb8[i] = pack_unorm4(outc.x, outc.y, outc.z, outc.w);
}

// traceRayEXT over an explicit ray list (vkp_trace_rays)
template <bool ANY>
__global__ void __launch_bounds__(VKP_TRAV_BLOCK) k_trace_list(const __grid_constant__ DScene S, const vkp_ray* __restrict__ rays, uint32_t n, vkp_hit* __restrict__ hits)
{
    __shared__ uint32_t s_trav[VKP_TRAV_SMEM_WORDS];
    uint32_t* const sstack = s_trav + threadIdx.x;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    vkp_ray r = rays[i];
    HitRec h;
    bool hit = trace_ray<ANY>(S, sstack, f3(r.origin[0], r.origin[1], r.origin[2]), f3(r.dir[0], r.dir[1], r.dir[2]), r.tmin, r.tmax, h);
    vkp_hit o; o.t = hit ? h.t : -1.0f; o.u = h.u; o.v = h.v; o.instance = h.inst; o.primitive = h.prim;
    hits[i] = o;
}

// instrumented traversal (vkp_trace_stats): node-pair fetches and triangle tests per ray, for build-quality studies
template <bool ANY>
__global__ void __launch_bounds__(VKP_TRAV_BLOCK) k_trace_stats(const __grid_constant__ DScene S, const vkp_ray* __restrict__ rays, uint32_t n, uint32_t* __restrict__ stats)
{
    __shared__ uint32_t s_trav[VKP_TRAV_SMEM_WORDS];
    uint32_t* const sstack = s_trav + threadIdx.x;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    vkp_ray r = rays[i];
    HitRec h;
    uint32_t st[2] = {0u, 0u};
    const long long t0 = clock64();
    trace_ray<ANY, true>(S, sstack, f3(r.origin[0], r.origin[1], r.origin[2]), f3(r.dir[0], r.dir[1], r.dir[2]), r.tmin, r.tmax, h, st);
    const long long t1 = clock64();
    stats[3 * (size_t)i] = st[0]; stats[3 * (size_t)i + 1] = st[1]; stats[3 * (size_t)i + 2] = (uint32_t)(t1 - t0);   // SM cycles for this ray's traversal
}

// ---- multi-GPU tile pack / unpack (SURVEY 8e composite) -------------------------------------------------------
template <typename E>
__global__ void k_tiles_copy(E* __restrict__ plane, E* __restrict__ packed, int W, int H, int tile_w, int tile_h, int tiles_x, int rank, int world, int to_plane)
{
    // one CTA per owned tile, one element (pixel) per thread iteration; rows of a tile are contiguous in both layouts
    const int own_tile = blockIdx.x;
    const int tile = rank + own_tile * world;
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    for (int e = threadIdx.x; e < tile_w * tile_h; e += blockDim.x) {
        const int r = e / tile_w, c = e - r * tile_w;
        const int x = tx * tile_w + c, y = ty * tile_h + r;
        if (x >= W || y >= H) continue;
        E* pp = plane + (size_t)y * W + x;
        E* qq = packed + (size_t)own_tile * tile_w * tile_h + e;
        if (to_plane) *pp = *qq; else *qq = *pp;
    }
}

// Brings a range into L2 without touching registers or L1 (prefetch.global.L2, one 128-byte line per thread iteration).  The frame
// kernels are bound by chains of DEPENDENT node fetches -- the longest ray of a kernel sets its floor -- and every one of those that
// misses L2 costs a round trip to HBM; streaming the 24 MB BVH arena in takes a few microseconds of bandwidth instead.
__global__ void k_prefetch_l2(const char* __restrict__ p, size_t bytes)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x * 128u;
    for (size_t off = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 128u; off < bytes; off += stride)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
}

__global__ void k_l2_flush(uint4* buf, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) buf[i] = make_uint4((unsigned)i, 1u, 2u, 3u);
}

// per-frame uniforms: copy `words` 32-bit words from mapped pinned host memory (or anywhere) into device memory, or clear them when
// src == nullptr.  A kernel rather than a DMA: copy-engine work queued on the render stream would sit behind the present copy of
// the previous frame in the engine's queue and serialise the frames in flight.
__global__ void k_stage_words(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint32_t words)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) dst[i] = src ? src[i] : 0u;
}

// the same for blocks of a few hundred bytes, with the data travelling as the kernel's own parameter: nothing is read over PCIe while
// the kernel runs (a read of pinned memory queues behind the posted writes of a present copy in flight -- measured +0.1 ms/frame)
template <int WORDS>
struct StagePayload { uint32_t w[WORDS]; };
template <int WORDS>
__global__ void k_stage_inline(uint32_t* __restrict__ dst, const __grid_constant__ StagePayload<WORDS> payload, uint32_t words)
{
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) dst[i] = payload.w[i];
}

// ---- frame barrier over NVLink peer memory (multi-GPU composite) ------------------------------------------
// The other ranks' resolve kernels store their pixels straight into rank 0's image; what remains of a "gather" is knowing when
// they are all there, and when rank 0 is done with a buffer.  Both are 4-byte flags in rank 0's memory, written with
// st.release.sys after a system fence and polled with ld.acquire.sys by a one-warp kernel in the waiting rank's stream: no
// collective library call, no host round trip per frame.  Flags carry frame numbers (wrap-safe compare); a wait gives up after
// ~10 s and raises the context's timeout flag instead of hanging the GPU when a peer died.
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__global__ void k_sync_signal(uint32_t* flag, uint32_t value)
{
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}
__global__ void k_sync_wait(const uint32_t* flags, int count, uint32_t value, uint32_t* timeout_flag)
{
    if ((int)threadIdx.x >= count) return;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(flags + threadIdx.x) - value) < 0) {
        if (clock64() - t0 > 20000000000LL) { atomicExch(timeout_flag, 1u); break; }
        __nanosleep(200);
    }
}

__global__ void k_eval_math(int func, const float* __restrict__ x, const float* __restrict__ y, uint32_t n, float* __restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float r;
    switch (func) {
    case 0: r = vkp_sinf(x[i]); break;
    case 1: r = vkp_cosf(x[i]); break;
    case 2: r = vkp_atan2f(x[i], y[i]); break;
    case 3: r = vkp_asinf(x[i]); break;
    case 4: r = vkp_powf(x[i], y[i]); break;
    case 5: r = vkp_log2f(x[i]); break;
    case 6: r = vkp_exp2f(x[i]); break;
    case 7: r = vkp_f16_to_f32(vkp_f32_to_f16(x[i])); break;
    default: r = (float)vkp_unorm8(x[i]); break;
    }
    out[i] = r;
}

inline dim3 frame_grid(const FrameLaunch& L)
{
    const int bpt = (L.frame.tile_w >> 4) * (L.frame.tile_h >> 3);
    return dim3((unsigned)(L.n_own_tiles * bpt));
}

} // namespace

// Every kernel that can share an SM with another one of the frame (the chain stream runs beside the main stream) asks for the SAME
// shared-memory / L1 split: an SM has to drain before its split can change, so two concurrent kernels that prefer different
// splits take turns on it instead of sharing it.  pct = share of the 228 KB that is shared memory (-1: leave the driver's choice).
cudaError_t launch_set_carveout(int pct)
{
    if (pct < 0) return cudaSuccess;
    const void* fns[] = {(const void*)k_gbuffer, (const void*)k_shadow<VKP_LB_SHADOW, true>, (const void*)k_shadow<VKP_LB_SHADOW, false>,
                         (const void*)k_shadow<VKP_LB_SHADOW_MANY, false>, (const void*)k_shadow<VKP_LB_SHADOW, true, true>, (const void*)k_chain, (const void*)k_chain_quad, (const void*)k_chain_q, (const void*)k_chain_scan, (const void*)k_chain_warp,
                         (const void*)k_chain_first, (const void*)k_chain_shadow0, (const void*)k_chain_shade0, (const void*)k_chain_rest,
                         (const void*)k_denoise, (const void*)k_denoise_accumulate4, (const void*)k_resolve<0>, (const void*)k_resolve<1>,
                         (const void*)k_resolve<2>, (const void*)k_resolve<3>, (const void*)k_resolve_paths, (const void*)k_bounce, (const void*)k_build_order,
                         (const void*)k_raytrace, (const void*)k_raytrace_warp, (const void*)k_atrous};
    for (const void* f : fns) {
        cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_gbuffer(const FrameLaunch& L) { k_gbuffer<<<frame_grid(L), 128, 0, L.stream>>>(L.scene, L.frame); return cudaGetLastError(); }
cudaError_t launch_shadow(const FrameLaunch& L, unsigned int* persist_counter, uint32_t* shadow_queue, unsigned int* shadow_qstate)
{
    if (L.scene.n_lights == 0) return cudaSuccess;
    dim3 g = frame_grid(L); g.x *= (unsigned)L.scene.n_lights;      // light-major: virtual block = light * n_cta + pixel block
    // Persistent grid or one CTA per block?  Measured (profiles/r3d_*, r3e_*): with one sample per light (cfg2) the hardware-scheduled grid
    // wins -- the persistent one fills every SM before the chain's kernels on the other streams get a slot, 0.615 -> 0.66-0.71 ms/frame;
    // with many samples per light (cfg5: 16 lights x 16 samples, the pass is 99 % of the frame and a thread's work varies 16-fold) the
    // warp-granular queue wins, 238.8 -> 223.5 ms.  VKP_SHADOW_PERSIST=<CTAs per SM> forces one or the other (0 = never).
    static int forced = -2, n_sm = 0;
    if (forced == -2) {
        const char* e = getenv("VKP_SHADOW_PERSIST"); forced = e ? atoi(e) : -1; if (forced > 16) forced = 16;
        int dev = 0; cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
    }
    const int persist = forced >= 0 ? forced : (L.frame.p.shadow_samples >= 4 ? VKP_LB_SHADOW_MANY : 0);
    if (L.frame.p.shadow_samples >= 4 && shadow_queue && shadow_qstate && L.frame.first_blocks == 0u) {
        // classify + compact + trace (above): VKP_SHADOW_COMPACT=1, with VKP_SHADOW_STREAM=0|1 choosing k_shadow_items or k_shadow_stream.
        // Both are bit-identical and both measured SLOWER than the persistent kernel below on cfg5 (16 lights x 16 samples, 4K): 221.6 ms
        // against 231.4 (compacted) and 310-338 (lanes refilled; profiles/r3o_*, r3p_*, r3q_*) -- off by default.  The 14 active lanes per
        // instruction ncu reports for that pass are not a throughput loss: the pass is bound by the latency of each warp's slowest ray
        // (IPC 2.3, 31 % warps active), which compaction does not shorten, and generating a ray (sampleDisk) costs as much as walking it,
        // so refilling single lanes replays that code at 1-3 active lanes.
        static int compact = -1;
        if (compact < 0) { const char* e = getenv("VKP_SHADOW_COMPACT"); compact = e ? (atoi(e) != 0) : 0; }
        if (compact) {
            const size_t qstride = (size_t)L.frame.p.width * L.frame.p.height;
            cudaError_t e = launch_stage_words(L.stream, shadow_qstate, nullptr, 64 * sizeof(unsigned int));
            if (e != cudaSuccess) return e;
            dim3 gc = frame_grid(L); gc.y = (unsigned)L.scene.n_lights;
            k_shadow_classify<<<gc, VKP_TRAV_BLOCK, 0, L.stream>>>(L.scene, L.frame, shadow_queue, shadow_qstate, qstride);
            static int stream_mode = -1;
            if (stream_mode < 0) { const char* es = getenv("VKP_SHADOW_STREAM"); stream_mode = es ? (atoi(es) != 0) : 1; }
            if (stream_mode) k_shadow_stream<VKP_LB_SHADOW_STREAM><<<(unsigned)(VKP_LB_SHADOW_STREAM * n_sm), VKP_TRAV_BLOCK, 0, L.stream>>>(L.scene, L.frame, shadow_queue, shadow_qstate, qstride);
            else k_shadow_items<VKP_LB_SHADOW_MANY><<<(unsigned)(VKP_LB_SHADOW_MANY * n_sm), VKP_TRAV_BLOCK, 0, L.stream>>>(L.scene, L.frame, shadow_queue, shadow_qstate, qstride);
            return cudaGetLastError();
        }
    }
    if (persist > 0 && persist_counter && g.x > 0) {
        cudaError_t e = launch_stage_words(L.stream, persist_counter, nullptr, sizeof(unsigned int));
        if (e != cudaSuccess) return e;
        const unsigned int units = g.x * 4u, grid = (unsigned)(persist * n_sm);
        if (L.frame.p.shadow_samples >= 4) k_shadow_persistent<VKP_LB_SHADOW_MANY, false><<<grid, VKP_TRAV_BLOCK, 0, L.stream>>>(L.scene, L.frame, persist_counter, units);
        else if (L.frame.p.shadow_samples == 1) k_shadow_persistent<VKP_LB_SHADOW, true><<<grid, VKP_TRAV_BLOCK, 0, L.stream>>>(L.scene, L.frame, persist_counter, units);
        else k_shadow_persistent<VKP_LB_SHADOW, false><<<grid, VKP_TRAV_BLOCK, 0, L.stream>>>(L.scene, L.frame, persist_counter, units);
        return cudaGetLastError();
    }
    if (L.frame.first_blocks) {
        g.x += L.frame.first_blocks * (1u + (unsigned)L.scene.n_lights);   // schedule 8: first-bounce blocks lead the grid, their any-hit blocks sit inside it
        if (L.frame.p.shadow_samples >= 4) k_shadow<VKP_LB_SHADOW_MANY, false, true><<<g, VKP_TRAV_BLOCK, 0, L.stream>>>(L.scene, L.frame);
        else if (L.frame.p.shadow_samples == 1) k_shadow<VKP_LB_SHADOW, true, true><<<g, VKP_TRAV_BLOCK, 0, L.stream>>>(L.scene, L.frame);
        else k_shadow<VKP_LB_SHADOW, false, true><<<g, VKP_TRAV_BLOCK, 0, L.stream>>>(L.scene, L.frame);
        return cudaGetLastError();
    }
    if (L.frame.p.shadow_samples >= 4) k_shadow<VKP_LB_SHADOW_MANY, false><<<g, VKP_TRAV_BLOCK, 0, L.stream>>>(L.scene, L.frame);
    else if (L.frame.p.shadow_samples == 1) k_shadow<VKP_LB_SHADOW, true><<<g, VKP_TRAV_BLOCK, 0, L.stream>>>(L.scene, L.frame);
    else k_shadow<VKP_LB_SHADOW, false><<<g, VKP_TRAV_BLOCK, 0, L.stream>>>(L.scene, L.frame);
    return cudaGetLastError();
}

// does launch_shadow run k_shadow as the persistent grid for this frame?  (then schedule 8 cannot fold the first bounces into it)
bool shadow_is_persistent(const FrameLaunch& L)
{
    const char* e = getenv("VKP_SHADOW_PERSIST");
    const int forced = e ? atoi(e) : -1;
    return (forced >= 0 ? forced : (L.frame.p.shadow_samples >= 4 ? VKP_LB_SHADOW_MANY : 0)) > 0;
}
unsigned int chain_first_blocks(const FrameLaunch& L) { return frame_grid(L).x / 4u + 1u; }

// schedule 8: what follows the shadow pass on the third stream -- the per-light any-hit rays of the first hits, their shading, and
// the (mispredicted) paths that go on
cudaError_t launch_chain_after_first(const FrameLaunch& L, cudaStream_t stream)
{
    const unsigned int px_blocks = frame_grid(L).x;
    const unsigned int g1 = px_blocks / 4u + 1u, g2 = px_blocks / 16u + 1u;
    if (L.scene.n_lights > 0 && !L.frame.first_blocks) k_chain_shadow0<<<dim3(g1, (unsigned)L.scene.n_lights), VKP_TRAV_BLOCK, 0, stream>>>(L.scene, L.frame);
    k_chain_shade0<<<g1, VKP_TRAV_BLOCK, 0, stream>>>(L.scene, L.frame);
    k_chain_rest<<<g2, VKP_TRAV_BLOCK, 0, stream>>>(L.scene, L.frame);
    return cudaGetLastError();
}
// schedules 7 / 8: the very long paths, a quad of lanes each, on their own stream beside k_chain
cudaError_t launch_chain_very_long(const FrameLaunch& L, cudaStream_t stream)
{
    if (!L.frame.chain_vl_list) return cudaSuccess;
    k_chain_quad<<<296, 32, 0, stream>>>(L.scene, L.frame);          // 8 paths per one-warp CTA: room for 2 368 very long paths in flight
    return cudaGetLastError();
}
cudaError_t launch_chain_first_only(const FrameLaunch& L, cudaStream_t stream)
{
    k_chain_first<<<frame_grid(L).x / 4u + 1u, VKP_TRAV_BLOCK, 0, stream>>>(L.scene, L.frame);
    return cudaGetLastError();
}
cudaError_t launch_build_order(cudaStream_t s, uint32_t* cost, uint32_t* order, uint32_t n)
{
    if (!n) return cudaSuccess;
    k_build_order<<<1, 256, 0, s>>>(cost, order, n);
    return cudaGetLastError();
}
cudaError_t launch_denoise(const FrameLaunch& L, uint8_t* atrous_a, uint8_t* atrous_b)
{
    if (L.frame.p.denoiser == 1 && atrous_a && atrous_b) {
        // three a-trous iterations (steps 1, 2, 4), then blur.comp's temporal stage on the filtered planes
        k_atrous<<<frame_grid(L), 128, 0, L.stream>>>(L.scene, L.frame, L.frame.shadow, atrous_a, 1);
        k_atrous<<<frame_grid(L), 128, 0, L.stream>>>(L.scene, L.frame, atrous_a, atrous_b, 2);
        k_atrous<<<frame_grid(L), 128, 0, L.stream>>>(L.scene, L.frame, atrous_b, atrous_a, 4);
        FrameLaunch T = L;
        T.frame.shadow = atrous_a;
        return launch_denoise(T, nullptr, nullptr);
    }
    if (L.frame.p.denoise_frame > 0 && L.frame.p.width % 4 == 0) k_denoise_accumulate4<<<frame_grid(L), 32, 0, L.stream>>>(L.scene, L.frame);
    else k_denoise<<<frame_grid(L), 128, 0, L.stream>>>(L.scene, L.frame);
    return cudaGetLastError();
}
cudaError_t launch_resolve(const FrameLaunch& L, int* launches)
{
    if (L.frame.p.wavefront == 0 || (L.frame.p.wavefront >= 2 && L.frame.p.max_bounces > VKP_CHAIN_MAX)) {
        k_resolve<0><<<frame_grid(L), 128, 0, L.stream>>>(L.scene, L.frame);
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    if (L.frame.p.wavefront >= 5) {
        k_resolve<3><<<frame_grid(L), 128, 0, L.stream>>>(L.scene, L.frame);     // needs k_chain_scan only; launch_resolve_paths follows the chain
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    if (L.frame.p.wavefront >= 2) {
        k_resolve<2><<<frame_grid(L), 128, 0, L.stream>>>(L.scene, L.frame);     // launch_chain has run (and been waited for) before
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    cudaError_t e = cudaMemsetAsync(L.frame.path_count, 0, sizeof(unsigned int) * (VKP_MAX_BOUNCES + 2), L.stream);
    if (e != cudaSuccess) return e;
    k_resolve<1><<<frame_grid(L), 128, 0, L.stream>>>(L.scene, L.frame);
    if (launches) *launches += 1;
    // the queue lengths live on the device: every bounce is a fixed-size grid-stride launch, no host round trip
    const int bounces = L.frame.p.max_bounces < VKP_MAX_BOUNCES ? L.frame.p.max_bounces : VKP_MAX_BOUNCES;
    for (int d = 0; d < bounces; d++) {
        k_bounce<<<148 * 8, 128, 0, L.stream>>>(L.scene, L.frame, d);
        if (launches) *launches += 1;
    }
    return cudaGetLastError();
}
cudaError_t launch_resolve_paths(const FrameLaunch& L)
{
    k_resolve_paths<<<frame_grid(L).x / 4u + 1u, 128, 0, L.stream>>>(L.frame);
    return cudaGetLastError();
}
// schedule 7, second half: the staged kernels over the paths predicted to end after one bounce, on their own stream beside k_chain
cudaError_t launch_chain_staged(const FrameLaunch& L, cudaStream_t stream)
{
    const unsigned int px_blocks = frame_grid(L).x;
    const unsigned int g1 = px_blocks / 4u + 1u, g2 = px_blocks / 16u + 1u;
    k_chain_first<<<g1, VKP_TRAV_BLOCK, 0, stream>>>(L.scene, L.frame);
    if (L.scene.n_lights > 0) k_chain_shadow0<<<dim3(g1, (unsigned)L.scene.n_lights), VKP_TRAV_BLOCK, 0, stream>>>(L.scene, L.frame);
    k_chain_shade0<<<g1, VKP_TRAV_BLOCK, 0, stream>>>(L.scene, L.frame);
    k_chain_rest<<<g2, VKP_TRAV_BLOCK, 0, stream>>>(L.scene, L.frame);
    return cudaGetLastError();
}

static bool chain_long_by_quad()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("VKP_CHAIN_LONG_QUAD"); v = e ? (atoi(e) != 0) : 0; }      // measured (profiles/r3g_*): 0.611 -> 0.657 ms/frame on one GPU: four times the 168-register threads take k_shadow's registers, and the chain stream is bound by the staged kernels, not by the long paths
    return v != 0;
}

static bool chain_long_by_warp(int world)
{
    static int forced = -2;
    if (forced == -2) { const char* e = getenv("VKP_CHAIN_LONG_WARP"); forced = e ? atoi(e) : -1; }
    (void)world;
    return forced > 0;      // measured on 8 GPUs: 0.194 vs 0.205 ms for the chain stream, no gain for the frame -- off unless asked for
}

cudaError_t launch_chain(const FrameLaunch& L, cudaStream_t stream, cudaEvent_t after_scan, uint32_t tag, unsigned int cont_cap)
{
    // path_count[0] = number of paths, [1] = head of the queue the persistent warps of mode 3 pull from
    // [2] = number of paths that go on after their first bounce (staged chain, mode 4)
    cudaError_t e = cudaSuccess;
    if (!L.frame.chain_in_gbuffer) {          // schedule 8: the counters were cleared before k_gbuffer, which queued the paths itself
        if ((e = launch_stage_words(stream, L.frame.path_count, nullptr, 7 * sizeof(unsigned int))) != cudaSuccess) return e;
        k_chain_scan<<<frame_grid(L), 128, 0, stream>>>(L.scene, L.frame);
    }
    if (after_scan && (e = cudaEventRecord(after_scan, stream)) != cudaSuccess) return e;
    if (L.frame.p.wavefront == 4) {
        // the queue lengths are only known on the device: grids sized for the usual share of path pixels, grid-stride loops beyond it
        const unsigned int px_blocks = frame_grid(L).x;
        const unsigned int g1 = px_blocks / 4u + 1u, g2 = px_blocks / 16u + 1u;
        k_chain_first<<<g1, VKP_TRAV_BLOCK, 0, stream>>>(L.scene, L.frame);
        if (L.scene.n_lights > 0) k_chain_shadow0<<<dim3(g1, (unsigned)L.scene.n_lights), VKP_TRAV_BLOCK, 0, stream>>>(L.scene, L.frame);
        k_chain_shade0<<<g1, VKP_TRAV_BLOCK, 0, stream>>>(L.scene, L.frame);
        k_chain_rest<<<g2, VKP_TRAV_BLOCK, 0, stream>>>(L.scene, L.frame);
    } else if (L.frame.p.wavefront == 6) {
        static int per_sm6 = 0, n_sm6 = 0;
        if (!per_sm6) {
            const char* e6 = getenv("VKP_CHAIN_PERSIST"); per_sm6 = e6 ? atoi(e6) : 3; if (per_sm6 < 1 || per_sm6 > 16) per_sm6 = 3;
            int dev = 0; cudaGetDevice(&dev);
            if (cudaDeviceGetAttribute(&n_sm6, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm6 <= 0) n_sm6 = 148;
        }
        k_chain_q<<<(unsigned)(per_sm6 * n_sm6), VKP_TRAV_BLOCK, 0, stream>>>(L.scene, L.frame, tag, cont_cap);
    } else if (L.frame.p.wavefront == 3) k_chain_warp<<<148 * VKP_LB_CHAIN, VKP_TRAV_BLOCK, 0, stream>>>(L.scene, L.frame);
    else if (L.frame.p.wavefront >= 7 && chain_long_by_quad()) {
        static int n_smq = 0;
        if (!n_smq) { int dev = 0; cudaGetDevice(&dev); if (cudaDeviceGetAttribute(&n_smq, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_smq <= 0) n_smq = 148; }
        k_chain_quad<<<(unsigned)(VKP_LB_CHAIN * n_smq), VKP_TRAV_BLOCK, 0, stream>>>(L.scene, L.frame);
    } else if (L.frame.p.wavefront >= 7 && chain_long_by_warp(L.frame.world)) {
        // VKP_CHAIN_LONG_WARP=1 (A/B): one WARP per long path walks every ray cooperatively (warp_trace) -- latency instead of
        // throughput.  1 GPU: more long paths than resident warps, 2x slower.  8 GPUs (2 k long paths per rank): chain stream 0.194 vs
        // 0.205 ms, frame unchanged -- there the staged kernels' four dependent tails are the floor.
        k_chain_warp<<<148 * VKP_LB_CHAIN, VKP_TRAV_BLOCK, 0, stream>>>(L.scene, L.frame);
    } else {
        // every queued path gets its own thread in ONE pass (a second pass would double the critical path): the queue length is
        // only known on the device, so the grid covers a quarter of the rank's pixels and the threads beyond the queue leave at once
        const unsigned int px_blocks = frame_grid(L).x;
        // CTAs of ONE warp (VKP_CHAIN_BLOCK, default 32): an SM gets a chain CTA's registers back when its last thread ends, and most
        // paths end after one bounce while a few go on for eight -- with 128-thread CTAs one long path kept 128 x 168 registers away
        // from k_shadow on the other stream
        // VKP_CHAIN_PERSIST = CTAs per SM of the persistent grid (default 3: what fits at 168 registers per thread; measured 1 .. 6)
        static int per_sm = 0, n_sm = 0;
        if (!per_sm) {
            const char* e = getenv("VKP_CHAIN_PERSIST"); per_sm = e ? atoi(e) : 3; if (per_sm < 1 || per_sm > 16) per_sm = 3;
            int dev = 0; cudaGetDevice(&dev);
            if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
        }
        (void)px_blocks;
        // Schedules 7 / 8 hand k_chain only the ~15 k long paths: 470 warps' worth, each alive for hundreds of microseconds.  Registers
        // are given back per CTA, so a 128-thread CTA with one warp still walking kept 21.5 K registers from k_shadow -- with 444 such
        // CTAs that was most of the GPU's register file (measured: the long paths cost the frame 0.08 ms).  One-warp CTAs return theirs
        // as soon as their own 32 paths are done (VKP_CHAIN_BLOCK=128 for the old shape).
        static int cblock = 0;
        if (!cblock) { const char* e = getenv("VKP_CHAIN_BLOCK"); cblock = e ? atoi(e) : 32; if (cblock != 32 && cblock != 64 && cblock != 128) cblock = 32; }
        const int blk = L.frame.p.wavefront >= 7 ? cblock : VKP_TRAV_BLOCK;
        k_chain<<<(unsigned)(per_sm * n_sm * (VKP_TRAV_BLOCK / blk)), blk, 0, stream>>>(L.scene, L.frame);
    }
    return cudaGetLastError();
}
cudaError_t launch_post(const FrameLaunch& L) { k_post<<<frame_grid(L), 128, 0, L.stream>>>(L.frame); return cudaGetLastError(); }
cudaError_t launch_raytrace(const FrameLaunch& L)
{
    if (L.frame.p.shadow_samples >= 8) {
        // many shadow samples per hit: one warp per pixel, lanes split the samples
        const long long warps = (long long)L.n_own_tiles * L.frame.tile_w * L.frame.tile_h;
        k_raytrace_warp<<<(unsigned)((warps + 3) / 4), 128, 0, L.stream>>>(L.scene, L.frame);
    } else {
        k_raytrace<<<frame_grid(L), 128, 0, L.stream>>>(L.scene, L.frame);
    }
    return cudaGetLastError();
}
cudaError_t launch_deferred(const FrameLaunch& L) { k_deferred<<<frame_grid(L), 128, 0, L.stream>>>(L.scene, L.frame); return cudaGetLastError(); }

cudaError_t launch_trace_rays(cudaStream_t s, const DScene& scene, const vkp_ray* rays, uint32_t n, int any_hit, vkp_hit* hits)
{
    if (n == 0) return cudaSuccess;
    unsigned blocks = (n + 127u) / 128u;
    if (any_hit) k_trace_list<true><<<blocks, 128, 0, s>>>(scene, rays, n, hits);
    else k_trace_list<false><<<blocks, 128, 0, s>>>(scene, rays, n, hits);
    return cudaGetLastError();
}

cudaError_t launch_trace_stats(cudaStream_t s, const DScene& scene, const vkp_ray* rays, uint32_t n, int any_hit, uint32_t* stats)
{
    if (n == 0) return cudaSuccess;
    unsigned blocks = (n + 127u) / 128u;
    if (any_hit) k_trace_stats<true><<<blocks, 128, 0, s>>>(scene, rays, n, stats);
    else k_trace_stats<false><<<blocks, 128, 0, s>>>(scene, rays, n, stats);
    return cudaGetLastError();
}

static int own_tiles(int W, int H, int tw, int th, int rank, int world, int* tiles_x_out)
{
    int tiles_x = (W + tw - 1) / tw, tiles_y = (H + th - 1) / th;
    if (tiles_x_out) *tiles_x_out = tiles_x;
    int total = tiles_x * tiles_y;
    return total > rank ? (total - rank + world - 1) / world : 0;
}

static cudaError_t tiles_copy(cudaStream_t s, void* plane, void* packed, int elem_bytes, int W, int H, int tw, int th, int rank, int world, int to_plane)
{
    int tiles_x; int n = own_tiles(W, H, tw, th, rank, world, &tiles_x);
    if (n == 0) return cudaSuccess;
    switch (elem_bytes) {
    case 1: k_tiles_copy<uint8_t><<<n, 128, 0, s>>>((uint8_t*)plane, (uint8_t*)packed, W, H, tw, th, tiles_x, rank, world, to_plane); break;
    case 4: k_tiles_copy<uint32_t><<<n, 128, 0, s>>>((uint32_t*)plane, (uint32_t*)packed, W, H, tw, th, tiles_x, rank, world, to_plane); break;
    case 8: k_tiles_copy<uint2><<<n, 128, 0, s>>>((uint2*)plane, (uint2*)packed, W, H, tw, th, tiles_x, rank, world, to_plane); break;
    case 16: k_tiles_copy<uint4><<<n, 128, 0, s>>>((uint4*)plane, (uint4*)packed, W, H, tw, th, tiles_x, rank, world, to_plane); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
cudaError_t launch_tiles_pack(cudaStream_t s, const void* plane, void* dst, int elem_bytes, int W, int H, int tw, int th, int rank, int world)
{
    return tiles_copy(s, const_cast<void*>(plane), dst, elem_bytes, W, H, tw, th, rank, world, 0);
}
cudaError_t launch_tiles_unpack(cudaStream_t s, void* plane, const void* src, int elem_bytes, int W, int H, int tw, int th, int rank, int world)
{
    return tiles_copy(s, plane, const_cast<void*>(src), elem_bytes, W, H, tw, th, rank, world, 1);
}
cudaError_t launch_build_luts(cudaStream_t s, float* lut_pow22, uint8_t* lut_post)
{
    k_build_luts<<<1, 256, 0, s>>>(lut_pow22, lut_post);
    return cudaGetLastError();
}
cudaError_t launch_eval_math(cudaStream_t s, int func, const float* x, const float* y, uint32_t n, float* out)
{
    if (n == 0) return cudaSuccess;
    k_eval_math<<<(n + 255u) / 256u, 256, 0, s>>>(func, x, y, n, out);
    return cudaGetLastError();
}
cudaError_t launch_stage_words(cudaStream_t s, void* dst, const void* src_or_null, size_t bytes)
{
    const uint32_t words = (uint32_t)(bytes / 4);
    if (!words) return cudaSuccess;
    unsigned blocks = (words + 255u) / 256u; if (blocks > 592u) blocks = 592u;
    k_stage_words<<<blocks, 256, 0, s>>>((uint32_t*)dst, (const uint32_t*)src_or_null, words);
    return cudaGetLastError();
}

template <int WORDS>
static cudaError_t stage_inline_t(cudaStream_t s, uint32_t* dst, const uint32_t* src, uint32_t words)
{
    StagePayload<WORDS> p;
    memcpy(p.w, src, sizeof(uint32_t) * words);
    k_stage_inline<WORDS><<<1, WORDS < 256 ? WORDS : 256, 0, s>>>(dst, p, words);
    return cudaGetLastError();
}

// bytes % 4 == 0 and bytes <= VKP_STAGE_INLINE_MAX: host data -> device memory through kernel parameters
cudaError_t launch_stage_inline(cudaStream_t s, void* dst, const void* src, size_t bytes)
{
    uint32_t words = (uint32_t)(bytes / 4);
    const uint32_t* w = (const uint32_t*)src;
    uint32_t* d = (uint32_t*)dst;
    while (words) {
        const uint32_t n = words > 960u ? 960u : words;
        cudaError_t e = n <= 64u ? stage_inline_t<64>(s, d, w, n) : n <= 256u ? stage_inline_t<256>(s, d, w, n) : stage_inline_t<960>(s, d, w, n);
        if (e != cudaSuccess) return e;
        words -= n; w += n; d += n;
    }
    return cudaSuccess;
}

cudaError_t launch_sync_signal(cudaStream_t s, uint32_t* flag, uint32_t value)
{
    k_sync_signal<<<1, 1, 0, s>>>(flag, value);
    return cudaGetLastError();
}
cudaError_t launch_sync_wait(cudaStream_t s, const uint32_t* flags, int count, uint32_t value, uint32_t* timeout_flag)
{
    if (count <= 0) return cudaSuccess;
    k_sync_wait<<<1, 64, 0, s>>>(flags, count, value, timeout_flag);
    return cudaGetLastError();
}

cudaError_t launch_prefetch_l2(cudaStream_t s, const void* p, size_t bytes)
{
    if (!p || !bytes) return cudaSuccess;
    k_prefetch_l2<<<592, 256, 0, s>>>(static_cast<const char*>(p), bytes);
    return cudaGetLastError();
}

cudaError_t launch_l2_flush(cudaStream_t s, void* buf, size_t bytes)
{
    k_l2_flush<<<1184, 256, 0, s>>>((uint4*)buf, bytes / 16);
    return cudaGetLastError();
}
