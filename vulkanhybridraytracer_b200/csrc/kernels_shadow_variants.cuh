// kernels_shadow_variants.cuh -- the shadow pass (shadowRaygen.rgen:28-117) restructured for many-sample frames: classify + compact + trace,
// and lanes refilled from the ray pool.  Both are bit-identical to k_shadow, both measured slower than the persistent warp-granular grid
// on BASELINE configs[4] and are off unless VKP_SHADOW_COMPACT / VKP_SHADOW_STREAM ask for them.
// Textually part of kernels_frame.cu (included inside its anonymous namespace, after k_shadow).
#pragma once

// ---- many-sample soft shadows: classify, compact, trace (round 2) -------------------------------------------------------------------
// With 16 lights x 16 samples (BASELINE configs[4]) ncu showed k_shadow at 14 active lanes per instruction: a pixel faces half of a ring
// of lights, and a lane whose (pixel, light) pair traces nothing idles while its neighbours loop over their samples.  So the pass is split:
// k_shadow_classify, one thread per (pixel, light), does the cheap part -- G-buffer, light geometry -- writes the result of every pair
// that needs no ray (glass, facing away, out of range) and COMPACTS the others into a per-light list (ballot + popc prefix, one atomic
// per warp); k_shadow_items, a persistent grid, pulls 32 listed pairs per warp and runs the sample loops with every lane busy.
// Same arithmetic per pair as shadow_patch: bit-identical planes and ray counts.
__global__ void __launch_bounds__(VKP_TRAV_BLOCK) k_shadow_classify(const __grid_constant__ DScene S, const __grid_constant__ DFrame F, uint32_t* __restrict__ queue,
                                                                    unsigned int* __restrict__ qstate, const size_t qstride)
{
    const int l = (int)blockIdx.y;
    const Pixel px = pixel_of_thread(F);
    bool lit = false;
    size_t i = 0;
    if (px.valid) {
        const int W = F.p.width, H = F.p.height;
        i = (size_t)px.y * W + px.x;
        const size_t plane = (size_t)W * H;
        const float4 pos4 = unpack_half4(F.position[i]);
        const float4 nrm4 = unpack_half4(F.normal[i]);
        const int mode = (int)material_by_float_index(S, pos4.w).shadingMetallicRoughness[0];
        if (mode == 4) F.shadow[l * plane + i] = 255;
        else {
            const float3 N = normalize3(xyz(nrm4) * 2.0f - f3s(1.0f));
            const LightGeom g = light_geom(S.lights[l], xyz(pos4), N);
            float shadowFactor = 0.0f;
            if (g.NdotL > 0.0f) {
                if (g.dist < g.maxd) lit = true;
                else {
                    for (int s = 0; s < F.p.shadow_samples; s++) shadowFactor += 1.0f;      // beyond maxDistance: unshadowed (shadowRaygen.rgen:81-83)
                    shadowFactor /= (float)F.p.shadow_samples;
                }
            }
            if (!lit) F.shadow[l * plane + i] = vkp_unorm8(shadowFactor);
        }
    }
    const unsigned int m = __ballot_sync(0xffffffffu, lit);
    if (m == 0u) return;
    const int lane = (int)(threadIdx.x & 31u), leader = __ffs(m) - 1;
    unsigned int base = 0u;
    if (lane == leader) base = atomicAdd(&qstate[l], (unsigned int)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (lit) queue[(size_t)l * qstride + base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)i;
}

template <int MINB>
__global__ void __launch_bounds__(VKP_TRAV_BLOCK, MINB) k_shadow_items(const __grid_constant__ DScene S, const __grid_constant__ DFrame F, const uint32_t* __restrict__ queue,
                                                                        unsigned int* __restrict__ qstate, const size_t qstride)
{
    __shared__ uint32_t s_trav[VKP_TRAV_SMEM_WORDS];
    uint32_t* const sstack = s_trav + threadIdx.x;
    const int lane = (int)(threadIdx.x & 31u);
    const int W = F.p.width;
    const size_t plane = (size_t)W * F.p.height;
    const int shadowSamples = F.p.shadow_samples;
    const bool advance = F.p.faithful_rng == 0;
    unsigned int nrays = 0;
    TravCount tc = {0u, 0u, 0u};
    TravCount* const tcp = F.count_trav ? &tc : nullptr;
    for (int l = 0; l < S.n_lights; l++) {
        const unsigned int n = qstate[l];
        const vkp_light light = S.lights[l];
        while (true) {
            unsigned int base = 0u;
            if (lane == 0) base = atomicAdd(&qstate[32 + l], 32u);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= n) break;
            const unsigned int k = base + (unsigned int)lane;
            if (k < n) {
                const size_t i = (size_t)queue[(size_t)l * qstride + k];
                const int py = (int)(i / (size_t)W), pxl = (int)(i - (size_t)py * W);
                const float4 pos4 = unpack_half4(F.position[i]);
                const float4 nrm4 = unpack_half4(F.normal[i]);
                const float3 position = xyz(pos4);
                const float3 N = normalize3(xyz(nrm4) * 2.0f - f3s(1.0f));
                const LightGeom g = light_geom(light, position, N);
                uint32_t seed = pixel_seed(F, i, pxl, py);
                if (advance) {
                    for (int kk = 0; kk < l; kk++) {
                        const LightGeom gk = light_geom(S.lights[kk], position, N);
                        if (gk.NdotL > 0.0f && gk.dist < gk.maxd)
                            for (int r = 0; r < 2 * shadowSamples; r++) seed = 1664525u * seed + 1013904223u;   // two rnd() per sampleDisk
                    }
                }
                float shadowFactor = 0.0f;
                for (int s = 0; s < shadowSamples; s++) {
                    const float3 dir = sample_disk(light, position, g.L, seed, advance);
                    HitRec h; nrays++;
                    if (!trace_ray<true, false, VKP_KEEPW_MANY != 0>(S, sstack, position + dir * 1e-1f, dir, 0.001f, g.dist + 1.0f, h, nullptr, tcp)) shadowFactor += 1.0f;
                }
                shadowFactor /= (float)shadowSamples;
                F.shadow[l * plane + i] = vkp_unorm8(shadowFactor);
            }
            __syncwarp();
        }
    }
    add_counter(F.counters, 1, nrays);
    if (F.count_trav) add_trav(F.counters, tc, 1u << 28);
}

// ---- k_shadow_stream: the same pass with LANES REFILLED FROM THE RAY POOL (north_star: "persistent-thread ... traversal", Aila & Laine's
// persistent while-while) ----------------------------------------------------------------------------------------------------------
// k_shadow_items still ran at 14 active lanes per instruction: the idle lanes are not unlit pixels, they are lanes whose any-hit ray has
// ended while a neighbour's is still walking -- the warp leaves trace_ray when its slowest ray does.  Here a lane never waits for the
// warp: its traversal state lives across an outer loop, and whenever its ray ends it takes the next sample of its (pixel, light) pair --
// or the next pair from the list -- while the other lanes go on walking.  Per outer iteration: (A) lanes without a pair draw new ones
// (one atomic per warp), (B) lanes without a ray make their next one or store their pair's result, (C) up to VKP_STREAM_STEPS node
// pairs per lane, then whatever the lane stands on (leaf, instance entry, return to the TLAS, end of ray).  An any-hit answer does not
// depend on the order of the walk and every sample's ray is generated by the same arithmetic: bit-identical planes and ray counts.
#ifndef VKP_STREAM_STEPS
#define VKP_STREAM_STEPS 6
#endif
#ifndef VKP_STREAM_REFILL
#define VKP_STREAM_REFILL 12
#endif
template <int MINB>
__global__ void __launch_bounds__(VKP_TRAV_BLOCK, MINB) k_shadow_stream(const __grid_constant__ DScene S, const __grid_constant__ DFrame F, const uint32_t* __restrict__ queue,
                                                                         unsigned int* __restrict__ qstate, const size_t qstride)
{
    __shared__ uint32_t s_trav[VKP_TRAV_SMEM_WORDS];
    const unsigned int full = 0xffffffffu;
    const int lane = (int)(threadIdx.x & 31u);
    const unsigned int lt = (1u << lane) - 1u;
    const int W = F.p.width;
    const size_t plane = (size_t)W * F.p.height;
    const int shadowSamples = F.p.shadow_samples;
    const bool advance = F.p.faithful_rng == 0;
    unsigned int nrays = 0;
    TravCount tc = {0u, 0u, 0u};
    TravCount* const tcp = F.count_trav ? &tc : nullptr;
    uint32_t scol = (uint32_t)__cvta_generic_to_shared(s_trav + threadIdx.x);
    asm volatile("" : "+r"(scol));
    uint32_t ovf[VKP_TRAV_STACK - VKP_SSTACK];
#define SPUSH(x) do { const uint32_t _v = (x); if (sp < VKP_SSTACK) asm volatile("st.shared.u32 [%0], %1;" :: "r"(scol + (uint32_t)sp * (4u * VKP_TRAV_BLOCK)), "r"(_v)); \
                      else ovf[sp - VKP_SSTACK] = _v; sp++; } while (0)
#define SPOP(dst) do { --sp; if (sp < VKP_SSTACK) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(dst) : "r"(scol + (uint32_t)sp * (4u * VKP_TRAV_BLOCK))); \
                       else dst = ovf[sp - VKP_SSTACK]; } while (0)
    for (int l = 0; l < S.n_lights; l++) {
        const unsigned int n = qstate[l];
        const vkp_light light = S.lights[l];
        bool exhausted = false;                         // the list of this light has been handed out
        // the lane's pair
        bool has_item = false;
        size_t pix = 0;
        float3 position = f3s(0.0f), gL = f3s(0.0f);
        float gdist = 0.0f;
        uint32_t seed = 0u;
        int s_next = 0;
        float shadowFactor = 0.0f;
        // the lane's ray
        bool has_ray = false;
        float3 o = f3s(0.0f), d = f3s(0.0f), ro = f3s(0.0f), rd = f3s(0.0f);
        float tmax = 0.0f;
        SlabRay slab = slab_setup(f3(1.0f, 1.0f, 1.0f), f3(1.0f, 1.0f, 1.0f));
        const float4* nodes = S.tlas_nodes2;
        const float4* tris = nullptr;
        uint32_t cur = VKP_STACK_SENTINEL;
        int inst = -1, sp = 0;
        while (true) {
            // ---- (A) new pairs for the lanes that have none ------------------------------------------------------------------------
            const unsigned int want = __ballot_sync(full, !has_item && !exhausted);
            if (want) {
                unsigned int base = 0u;
                if (lane == __ffs(want) - 1) base = atomicAdd(&qstate[32 + l], (unsigned int)__popc(want));
                base = __shfl_sync(full, base, __ffs(want) - 1);
                if (!has_item && !exhausted) {
                    const unsigned int k = base + (unsigned int)__popc(want & lt);
                    if (k < n) {
                        pix = (size_t)queue[(size_t)l * qstride + k];
                        const int py = (int)(pix / (size_t)W), pxl = (int)(pix - (size_t)py * W);
                        const float4 pos4 = unpack_half4(F.position[pix]);
                        const float4 nrm4 = unpack_half4(F.normal[pix]);
                        position = xyz(pos4);
                        const float3 N = normalize3(xyz(nrm4) * 2.0f - f3s(1.0f));
                        const LightGeom g = light_geom(light, position, N);
                        gL = g.L; gdist = g.dist;
                        seed = pixel_seed(F, pix, pxl, py);
                        if (advance) {
                            for (int kk = 0; kk < l; kk++) {
                                const LightGeom gk = light_geom(S.lights[kk], position, N);
                                if (gk.NdotL > 0.0f && gk.dist < gk.maxd)
                                    for (int r = 0; r < 2 * shadowSamples; r++) seed = 1664525u * seed + 1013904223u;   // two rnd() per sampleDisk
                            }
                        }
                        s_next = 0; shadowFactor = 0.0f; has_item = true;
                    } else exhausted = true;
                }
            }
            // ---- (B) next sample of the pair, or its result ---------------------------------------------------------------------------
            // (generating a ray -- sampleDisk: a square root, a sine / cosine pair, three normalisations -- costs about as much as walking
            // it, so it is not done for one lane at a time: the lanes that need a ray wait until VKP_STREAM_REFILL of them do, or nobody walks)
            const unsigned int idle = __ballot_sync(full, has_item && !has_ray);
            const bool refill = __popc(idle) >= VKP_STREAM_REFILL || !__any_sync(full, has_ray);
            if (refill && has_item && !has_ray) {
                if (s_next < shadowSamples) {
                    const float3 dir = sample_disk(light, position, gL, seed, advance);
                    s_next++; nrays++;
                    o = position + dir * 1e-1f; d = dir; tmax = gdist + 1.0f;
                    if (S.n_instances == 0) shadowFactor += 1.0f;                                // nothing to hit
                    else if (needs_general_path(d)) {
                        HitRec h;
                        if (!trace_ray_general<true>(S, o, d, 0.001f, tmax, h, tcp)) shadowFactor += 1.0f;
                    } else {
                        ro = o; rd = d; slab = slab_setup(o, d);
                        nodes = S.tlas_nodes2; tris = nullptr; inst = -1; sp = 0;
                        SPUSH(VKP_STACK_SENTINEL);
                        cur = S.tlas_root_ref2;
                        has_ray = true;
                    }
                } else {
                    shadowFactor /= (float)shadowSamples;
                    F.shadow[l * plane + pix] = vkp_unorm8(shadowFactor);
                    has_item = false;
                }
            }
            if (!__any_sync(full, has_item)) { if (__all_sync(full, exhausted)) break; else continue; }
            // ---- (C) walk: a few node pairs, then what the lane stands on ------------------------------------------------------------------
            if (has_ray) {
                for (int step = 0; step < VKP_STREAM_STEPS && !(cur & VKP_LEAF_BIT); step++) {
                    const float4* c = nodes + 2u * (size_t)cur;
                    if (tcp) tcp->nodes++;
                    const float4 lo0 = __ldg(c + 0), hi0 = __ldg(c + 1), lo1 = __ldg(c + 2), hi1 = __ldg(c + 3);
                    float t0, t1;
                    const bool h0 = box_test_fast(lo0, hi0, slab, 0.001f, tmax, t0);
                    const bool h1 = box_test_fast(lo1, hi1, slab, 0.001f, tmax, t1);
                    const uint32_t r0 = __float_as_uint(lo0.w), r1 = __float_as_uint(lo1.w);
                    if (h0 && h1) {
                        const bool swap = t1 < t0;
                        SPUSH(swap ? r0 : r1);
                        cur = swap ? r1 : r0;
                    } else if (h0) cur = r0;
                    else if (h1) cur = r1;
                    else SPOP(cur);
                }
                if (cur & VKP_LEAF_BIT) {
                    if (cur == VKP_STACK_SENTINEL) { shadowFactor += 1.0f; has_ray = false; }           // walked everything: unshadowed
                    else if (cur == VKP_STACK_RETURN) {
                        inst = -1; ro = o; rd = d; slab = slab_setup(o, d); nodes = S.tlas_nodes2;
                        SPOP(cur);
                    } else if (inst < 0) {
                        inst = (int)((cur & 0x7fffffffu) >> 3);
                        if (tcp) tcp->entries++;
                        const float4* rec = reinterpret_cast<const float4*>(S.inst_trav + inst);
                        const float4 q0 = __ldg(rec), q1 = __ldg(rec + 1), q2 = __ldg(rec + 2), pp = __ldg(rec + 3), rr = __ldg(rec + 4);
                        enter_instance(q0, q1, q2, o, d, ro, rd);
                        if (needs_general_path(rd)) {
                            // a rotated instance met along an axis (rare): this ray again, out of line, from the start
                            HitRec h;
                            if (!trace_ray_general<true>(S, o, d, 0.001f, tmax, h, tcp)) shadowFactor += 1.0f;
                            has_ray = false;
                        } else {
                            slab = slab_setup(ro, rd);
                            nodes = ptr_from_words(pp.x, pp.y); tris = ptr_from_words(pp.z, pp.w);
                            SPUSH(VKP_STACK_RETURN);
                            cur = __float_as_uint(rr.x);
                        }
                    } else {
                        const uint32_t first = (cur & 0x7fffffffu) >> 3, count = (cur & 7u) + 1u;
                        const float4* tp = tris + 3u * (size_t)first;
                        float4 ta[4], tb[4], tcv[4];
#pragma unroll
                        for (uint32_t k = 0; k < 4; k++)
                            if (k < count) { ta[k] = __ldg(tp + 3 * k); tb[k] = __ldg(tp + 3 * k + 1); tcv[k] = __ldg(tp + 3 * k + 2); }
                        bool blocked = false;
#pragma unroll
                        for (uint32_t k = 0; k < 4; k++) {
                            if (k < count && !blocked) {
                                if (tcp) tcp->tris++;
                                float t, u, v;
                                blocked = tri_test(ro, rd, f3(ta[k].x, ta[k].y, ta[k].z), f3(tb[k].x, tb[k].y, tb[k].z), f3(tcv[k].x, tcv[k].y, tcv[k].z), 0.001f, tmax, t, u, v);
                            }
                        }
                        if (blocked) has_ray = false;                                                 // any hit ends the ray: shadowed
                        else SPOP(cur);
                    }
                }
            }
        }
    }
#undef SPUSH
#undef SPOP
    add_counter(F.counters, 1, nrays);
    if (F.count_trav) add_trav(F.counters, tc, 1u << 28);
}

// blur.comp:35-82 for one pixel and one light, with runtime W/H/NLIGHTS and an integer motion fetch; returns the new
// R8 history value (pixels outside the faithful dispatch area keep their history)
__device__ __forceinline__ uint8_t denoise_pixel(const DFrame& F, int l, int x, int y, size_t i, size_t plane)
{
    const int W = F.p.width, H = F.p.height;
    const int frame = F.p.denoise_frame;
    const int Wc = F.p.denoise_cover_remainder ? W : (W / 16) * 16;
    const int Hc = F.p.denoise_cover_remainder ? H : (H / 16) * 16;
    const uint8_t* in = F.shadow + l * plane;
    const uint8_t* hist = F.denoised + l * plane;
    if (!(x < Wc && y < Hc)) return hist[i];
    float pixelColor;
    if (frame > 0) {
        float center = unit255(in[i]);
        float old = unit255(hist[i]);
        float a = 1.0f / (float)((float)frame + 1.0f);
        pixelColor = mix1(old, center, a);
    } else {
        uint32_t m = F.motion[i];
        float rx = vkp_f16_to_f32((uint16_t)(m & 0xffffu)) * 2.0f - 1.0f;
        float lastx = (float)x + rx * (float)W, lasty = (float)y + rx * (float)H;   // blur.comp:53 (r.x twice)
        if (F.p.reproject_fix) {
            // motion = previous - current position in NDC (see k_gbuffer), NDC spans 2 over the screen: where this surface point was last frame
            float ry = vkp_f16_to_f32((uint16_t)(m >> 16)) * 2.0f - 1.0f;
            lastx = floorf(((float)x + 0.5f) + (rx * 0.5f) * (float)W);
            lasty = floorf(((float)y + 0.5f) + (ry * 0.5f) * (float)H);
        }
        float center = unit255(in[i]);
        float minColor = center, maxColor = center;
#pragma unroll
        for (int yy = -1; yy <= 1; yy++)
#pragma unroll
            for (int xx = -1; xx <= 1; xx++) {
                if (xx == 0 && yy == 0) continue;
                int qx = x + xx, qy = y + yy;
                float c = (qx < 0 || qy < 0 || qx >= W || qy >= H) ? 0.0f : unit255(in[(size_t)qy * W + qx]);
                minColor = fminf(minColor, c); maxColor = fmaxf(maxColor, c);
            }
        float cx = fminf(fmaxf(lastx, -2147483000.0f), 2147483000.0f), cy = fminf(fmaxf(lasty, -2147483000.0f), 2147483000.0f);
        if (cx != cx) cx = -1.0f;
        if (cy != cy) cy = -1.0f;
        int ix = (int)cx, iy = (int)cy;
        float old = (ix < 0 || iy < 0 || ix >= W || iy >= H) ? 0.0f : unit255(hist[(size_t)iy * W + ix]);
        old = fmaxf(minColor, old);
        old = fminf(maxColor, old);
        pixelColor = mix1(old, center, 0.4f);
    }
    return vkp_unorm8(pixelColor);
}

