// kernels_chain.cuh -- the reflection / refraction chain of hybridRaygen.rgen:153-170 + hybridHit.rchit + miss.rmiss, in every schedule
// that was built and measured (vkp_params.wavefront 2 .. 8): path compaction (k_chain_scan / chain_enqueue),
// k_chain (persistent, one thread per path), k_chain_q (persistent wavefront), k_chain_warp (one warp per path), the staged first bounce
// (k_chain_first / _shadow0 / _shade0 / _rest) and k_chain_quad (four lanes per path).  Every variant stores the same bits.
// Textually part of kernels_frame.cu (included inside its anonymous namespace, between the filter kernels and k_resolve).
#pragma once

// hybridRaygen.rgen:27-186
// first reflection / refraction ray of a mirror / glass pixel (hybridRaygen.rgen:123-148): depends on the G-buffer only
__device__ __forceinline__ void first_bounce(int shadingMode, const vkp_material& mat, float3 position, float3 N, float3 V, float3& origin, float3& direction)
{
    if (shadingMode == 3) {
        direction = reflect3(normalize3(-V), N);
        origin = position;
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
    }
}

// The bounce chain of hybridRaygen.rgen:153-170 depends only on the G-buffer, not on the shadow / denoise passes, so in
// chain mode 2 it runs on a second, high-priority stream CONCURRENTLY with k_shadow + k_denoise; the per-bounce colours are
// stored (<= 8 per path) and k_resolve<2> multiplies them into rayColor in the reference's order.
//
// k_chain_scan compacts the mirror / glass pixels (3 % of the default frame) into a path list: one atomic per warp, slots by
// ballot + popc prefix.  k_chain is a PERSISTENT grid fed from that queue: every warp takes the next path with one atomic and
// walks it with ALL 32 lanes -- the shading is computed redundantly (one instruction stream, broadcast loads), every ray of the
// path is traced cooperatively (warp_trace, vkp_bvh.cuh).  A path is a chain of up to 8 dependent closest-hit rays and its longest
// ray set the duration of the whole kernel when one thread walked it alone (round 1: 270 us for 3 % of the pixels).
__global__ void __launch_bounds__(128) k_chain_scan(const __grid_constant__ DScene S, const __grid_constant__ DFrame F)
{
    const Pixel px = pixel_of_thread(F);
    bool path = false;
    size_t i = 0;
    if (px.valid) {
        i = (size_t)px.y * F.p.width + px.x;
        const float matIdx = vkp_f16_to_f32((uint16_t)(F.position[i].y >> 16));
        path = (int)material_by_float_index(S, matIdx).shadingMetallicRoughness[0] != 0 && F.p.max_bounces > 0;
    }
    chain_enqueue(F, i, path);
}

// path compaction: one atomic per warp, slots by ballot + popc prefix.  Every lane of the warp calls it.
__device__ __forceinline__ void chain_enqueue(const DFrame& F, size_t i, bool path)
{
    const int lane = threadIdx.x & 31;
    if (F.chain_split) {
        // schedule 7: two path lists in one array.  Paths that went beyond their first bounce in the previous frame (chain_long, one byte
        // per pixel, written by the chain kernels) are queued from slot 0 upwards for k_chain -- the few long paths start first --
        // the others from the last slot downwards for the staged kernels.  A wrong prediction costs time, never a bit.
        const bool lng = path && F.chain_long[i] >= 2;
        const bool sht = path && !lng;
        const unsigned ml = __ballot_sync(0xffffffffu, lng), ms = __ballot_sync(0xffffffffu, sht);
        unsigned bl = 0, bs = 0;
        if (ml && lane == __ffs(ml) - 1) bl = atomicAdd(&F.path_count[0], (unsigned)__popc(ml));
        if (ms && lane == __ffs(ms) - 1) bs = atomicAdd(&F.path_count[4], (unsigned)__popc(ms));
        if (ml) bl = __shfl_sync(0xffffffffu, bl, __ffs(ml) - 1);
        if (ms) bs = __shfl_sync(0xffffffffu, bs, __ffs(ms) - 1);
        if (lng) {
            const unsigned slot = bl + __popc(ml & ((1u << lane) - 1u)); F.chain_slot[i] = (int)slot; F.chain_pix[slot] = (int)i;
            // the handful of paths that made >= VKP_CHAIN_VL_BOUNCES bounces last frame set the end of the whole chain (every bounce is
            // four dependent rays): they get a quad of lanes each (k_chain_quad) and k_chain skips them
            const bool vl = F.chain_vl_list && F.chain_long[i] >= VKP_CHAIN_VL_BOUNCES;
            F.chain_kind[slot] = vl ? 1 : 0;
            if (vl) F.chain_vl_list[atomicAdd(&F.path_count[5], 1u)] = (int)slot;
        }
        if (sht) { const unsigned slot = F.chain_cap - 1u - (bs + __popc(ms & ((1u << lane) - 1u))); F.chain_slot[i] = (int)slot; F.chain_pix[slot] = (int)i; }
        return;
    }
    const unsigned mask = __ballot_sync(0xffffffffu, path);
    if (mask == 0u) return;
    const int leader = __ffs(mask) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(&F.path_count[0], (unsigned)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (path) {
        const unsigned slot = base + __popc(mask & ((1u << lane) - 1u));
        F.chain_slot[i] = (int)slot;
        F.chain_pix[slot] = (int)i;
    }
}

// the staged kernels' path list: slots [0, path_count[0]) (schedule 4) or, counted downwards from the last slot, path_count[4] of them (7)
__device__ __forceinline__ unsigned int staged_count(const DFrame& F) { return F.chain_split ? F.path_count[4] : F.path_count[0]; }
__device__ __forceinline__ unsigned int staged_slot(const DFrame& F, unsigned int k) { return F.chain_split ? F.chain_cap - 1u - k : k; }

// the first reflection / refraction ray of a queued path, from the G-buffer (origin already offset along the direction)
__device__ __forceinline__ void chain_first_ray(const DScene& S, const DFrame& F, size_t i, float3& ro, float3& direction)
{
    const float4 pos4 = unpack_half4(F.position[i]);
    const vkp_material& mat = material_by_float_index(S, pos4.w);
    const int shadingMode = (int)mat.shadingMetallicRoughness[0];
    const float4 nrm4 = unpack_half4(F.normal[i]);
    const float3 position = xyz(pos4);
    const float3 N = normalize3(xyz(nrm4) * 2.0f - f3s(1.0f));
    const float3 camPosition = f3(F.cam.inv_view[12], F.cam.inv_view[13], F.cam.inv_view[14]);
    const float3 V = normalize3(camPosition - position);
    float3 origin = f3s(0.0f); direction = f3s(0.0f);
    if (S.n_lights > 0) first_bounce(shadingMode, mat, position, N, V, origin, direction);   // upstream sets them inside the light loop
    ro = origin + direction * 1e-2f;
}

// One THREAD per queued path (chain mode 2, the default): all paths are in flight at once -- a path is a chain of dependent
// rays, so what counts is how many of them overlap, and one thread per path keeps every path of the frame resident.
__global__ void __launch_bounds__(VKP_TRAV_BLOCK, VKP_LB_CHAIN) k_chain(const __grid_constant__ DScene S, const __grid_constant__ DFrame F)
{
    __shared__ uint32_t s_trav[VKP_TRAV_SMEM_WORDS];
    uint32_t* const sstack = s_trav + threadIdx.x;
    const unsigned int n_paths = F.path_count[0];
    unsigned int n_refl = 0, n_sec = 0;
    TravCount tc = {0u, 0u, 0u};
    TravCount* const tcp = F.count_trav ? &tc : nullptr;      // traversal accounting only on request (vkp_set_profiling(ctx, 2))
    // A PERSISTENT grid fed from the path queue (north_star: "persistent-thread ... traversal"): every warp takes the next 32 paths with
    // one atomic until the queue is empty.  The grid is a few CTAs per SM, all resident from the start.  That matters for the OTHER
    // stream: while CTAs of this high-priority kernel are waiting for a slot, the block scheduler places nothing of k_shadow, so a
    // one-thread-per-path grid (three waves at this register count) kept the main stream out for half the chain's duration.
    const int lane = (int)(threadIdx.x & 31u);
    while (true) {
        unsigned int base = 0u;
        if (lane == 0) base = atomicAdd(&F.path_count[1], 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_paths) break;
        const unsigned int slot = base + (unsigned int)lane;
        if (slot < n_paths && !(F.chain_vl_list && F.chain_kind[slot])) {
        const size_t i = (size_t)F.chain_pix[slot];
        const int py = (int)(i / (size_t)F.p.width), pxl = (int)(i - (size_t)py * F.p.width);
        const float4 pos4 = unpack_half4(F.position[i]);
        float4* rec = F.chain_buf + (size_t)slot * VKP_CHAIN_MAX;
        const vkp_material& mat = material_by_float_index(S, pos4.w);
        const int shadingMode = (int)mat.shadingMetallicRoughness[0];
        const float4 nrm4 = unpack_half4(F.normal[i]);
        const float3 position = xyz(pos4);
        const float3 N = normalize3(xyz(nrm4) * 2.0f - f3s(1.0f));
        const float3 camPosition = f3(F.cam.inv_view[12], F.cam.inv_view[13], F.cam.inv_view[14]);
        const float3 V = normalize3(camPosition - position);
        float3 origin = f3s(0.0f), direction = f3s(0.0f);
        if (S.n_lights > 0) first_bounce(shadingMode, mat, position, N, V, origin, direction);   // upstream sets them inside the light loop
        Payload prd;
        prd.seed = pixel_seed(F, i, pxl, py);
        int n = 0;
        for (int depth = 0; depth < F.p.max_bounces; depth++) {
            float3 ro = origin + direction * 1e-2f;
            HitRec h; n_refl++;
#if VKP_CHAIN_CALL
            const bool hit = trace_ray_call<false>(S, sstack, ro, direction, 0.001f, 1000.0f, h, tcp);
            if (hit) shade_hit<false, 0, false, true>(S, sstack, F.p, ro, direction, h, prd, n_sec, tcp); else shade_miss(S, direction, prd);
#else
            const bool hit = trace_ray<false>(S, sstack, ro, direction, 0.001f, 1000.0f, h, nullptr, tcp);
            if (hit) shade_hit<false>(S, sstack, F.p, ro, direction, h, prd, n_sec, tcp); else shade_miss(S, direction, prd);
#endif
            rec[n] = make_float4(prd.color.x, prd.color.y, prd.color.z, 0.0f);
            n++;
            if (!hit || !(prd.scatter > 0.0f)) break;
            origin = prd.origin; direction = prd.direction;
        }
        F.chain_len[slot] = n;
        if (F.chain_long) F.chain_long[i] = (uint8_t)n;
        }
        __syncwarp();
    }
    add_counter(F.counters, 2, n_refl);
    add_counter(F.counters, 3, n_sec);
    if (F.count_trav) add_trav(F.counters, tc);
}

// ---- chain schedule 6: a persistent WAVEFRONT (north_star: "persistent-thread ... traversal ... ray compaction with warp ballot/prefix") ----
// k_chain keeps a path in its thread for up to 8 bounces, and 91 % of the paths end after the first: a warp carries on for the few
// lanes that go on, holding 32 x 168 registers for one or two live paths (61 us per 32 paths where a single bounce takes ~25).
// Here ONE BOUNCE is the unit of work.  A persistent grid pulls items from a queue -- items [0, n_paths) are the first bounces of the
// compacted path list, the rest are continuations -- traces and shades one bounce per lane, and a path that scatters is APPENDED to
// the queue (one atomic per warp, slots by ballot + popc prefix) instead of looping: warps always work on full batches.
//   path_count[0] n_paths   [1] head (next item to hand out)   [2] continuations reserved so far   [3] items completed
// A continuation becomes visible through a per-item tag written with st.release after its data (consumers that got the index
// early spin on ld.acquire).  The grid ends when every item ever created has been completed: continuations are reserved before
// their parent is counted, so completed == created means nothing is in flight.  If the continuation array is full the lane simply
// goes on with its path itself.  Same arithmetic as k_chain: every stored colour is bit-identical.
__global__ void __launch_bounds__(VKP_TRAV_BLOCK, VKP_LB_CHAIN) k_chain_q(const __grid_constant__ DScene S, const __grid_constant__ DFrame F, const uint32_t tag,
                                                                          const unsigned int cont_cap)
{
    __shared__ uint32_t s_trav[VKP_TRAV_SMEM_WORDS];
    uint32_t* const sstack = s_trav + threadIdx.x;
    const unsigned int full = 0xffffffffu;
    const int lane = (int)(threadIdx.x & 31u);
    const unsigned int lt = (1u << lane) - 1u;
    const unsigned int n_paths = F.path_count[0];
    unsigned int* const head = &F.path_count[1];
    unsigned int* const conts = &F.path_count[2];
    unsigned int* const completed = &F.path_count[3];
    PathItem* const cq = F.path_q[1];
    unsigned int n_refl = 0, n_sec = 0;
    TravCount tc = {0u, 0u, 0u};
    TravCount* const tcp = F.count_trav ? &tc : nullptr;
    // A warp CLAIMS 32 consecutive item indices with one atomicAdd (no compare-and-swap: 1776 warps retrying on one word took
    // milliseconds) and works through its claim as the items come into existence: [cur, end) is what is left of it.
    unsigned int cur = 0u, end = 0u;
    while (true) {
        // ---- the warp's next batch --------------------------------------------------------------------------------------------
        unsigned int base = 0u, cnt = 0u;
        if (lane == 0) {
            unsigned int backoff = 64u;
            while (true) {
                if (cur == end) { cur = atomicAdd(head, 32u); end = cur + 32u; }
                const unsigned int t = n_paths + ld_volatile_u32(conts);              // items created so far
                if (cur < t) { base = cur; cnt = (end < t ? end : t) - cur; cur += cnt; break; }
                const unsigned int done = ld_volatile_u32(completed), t2 = n_paths + ld_volatile_u32(conts);
                if (done == t2 && t2 == t) break;                      // everything ever created is completed: nothing can appear any more
                __nanosleep(backoff);
                if (backoff < 1024u) backoff <<= 1;
            }
        }
        base = __shfl_sync(full, base, 0); cnt = __shfl_sync(full, cnt, 0);
        if (cnt == 0u) break;
        // ---- one bounce per lane -------------------------------------------------------------------------------------------------
        bool active = (unsigned int)lane < cnt;
        unsigned int slot = 0u;
        int depth = 0;
        float3 ro = f3s(0.0f), direction = f3s(0.0f);
        Payload prd; prd.seed = 0u;
        if (active) {
            const unsigned int k = base + (unsigned int)lane;
            if (k < n_paths) {
                slot = k;
                const size_t i = (size_t)F.chain_pix[slot];
                const int py = (int)(i / (size_t)F.p.width), pxl = (int)(i - (size_t)py * F.p.width);
                chain_first_ray(S, F, i, ro, direction);
                prd.seed = pixel_seed(F, i, pxl, py);
            } else {
                PathItem* it = cq + (k - n_paths);
                while (ld_acquire_gpu(reinterpret_cast<const uint32_t*>(&it->d.w)) != tag) __nanosleep(40);      // reserved, not yet written
                const float4 a = it->a, b = it->b, c = it->c;
                slot = __float_as_uint(a.w); depth = __float_as_int(c.x);
                direction = f3(b.x, b.y, b.z);
                ro = f3(a.x, a.y, a.z) + direction * 1e-2f;
                prd.seed = __float_as_uint(b.w);
            }
        }
        while (__any_sync(full, active)) {
            bool again = false;
            if (active) {
                HitRec h; n_refl++;
                const bool hit = trace_ray<false>(S, sstack, ro, direction, 0.001f, 1000.0f, h, nullptr, tcp);
                if (hit) shade_hit<false>(S, sstack, F.p, ro, direction, h, prd, n_sec, tcp); else shade_miss(S, direction, prd);
                F.chain_buf[(size_t)slot * VKP_CHAIN_MAX + depth] = make_float4(prd.color.x, prd.color.y, prd.color.z, 0.0f);
                again = hit && (prd.scatter > 0.0f) && depth + 1 < F.p.max_bounces;
                if (!again) { F.chain_len[slot] = depth + 1; active = false; }
            }
            // the paths that go on: appended to the queue, compacted
            const unsigned int m = __ballot_sync(full, again);
            if (m) {
                unsigned int at = 0xffffffffu;
                if (lane == __ffs(m) - 1) {
                    const unsigned int want = (unsigned int)__popc(m);
                    if (ld_volatile_u32(conts) + want + 4096u <= cont_cap) at = atomicAdd(conts, want);     // (slack: other warps reserve concurrently)
                }
                at = __shfl_sync(full, at, __ffs(m) - 1);
                if (again) {
                    if (at != 0xffffffffu) {
                        PathItem* it = cq + at + __popc(m & lt);
                        it->a = make_float4(prd.origin.x, prd.origin.y, prd.origin.z, __uint_as_float(slot));
                        it->b = make_float4(prd.direction.x, prd.direction.y, prd.direction.z, __uint_as_float(prd.seed));
                        it->c = make_float4(__int_as_float(depth + 1), 0.0f, 0.0f, 0.0f);
                        st_release_gpu(reinterpret_cast<uint32_t*>(&it->d.w), tag);
                        active = false;
                    } else {
                        // no room: this lane walks the rest of its path itself
                        depth++;
                        direction = prd.direction;
                        ro = prd.origin + direction * 1e-2f;
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) { __threadfence(); atomicAdd(completed, cnt); }
    }
    add_counter(F.counters, 2, n_refl);
    add_counter(F.counters, 3, n_sec);
    if (F.count_trav) add_trav(F.counters, tc);
}

// One WARP per queued path (chain mode 3): a PERSISTENT grid fed from the path queue -- every warp takes the next path with one
// atomic and walks it with all 32 lanes: the shading is computed redundantly (one instruction stream, broadcast loads), every ray
// of the path is traced cooperatively (warp_trace, vkp_bvh.cuh).  Built to shorten the critical path of a path; measured on the
// 1080p default frame it is 8x SLOWER than one thread per path (2.1 ms against 0.27): 62 000 paths x 32 lanes want 2 M threads
// where a B200 holds 57 000 at this kernel's register count, so the paths run in 35 consecutive rounds instead of all at once.
// Kept selectable (vkp_params.wavefront = 3) and bit-identical; DESIGN.md section 4 has the numbers.
__global__ void __launch_bounds__(VKP_TRAV_BLOCK, VKP_LB_CHAIN) k_chain_warp(const __grid_constant__ DScene S, const __grid_constant__ DFrame F)
{
    __shared__ uint32_t s_wq[VKP_WQ_SMEM_WORDS];
    const WarpQueue Q = warp_queue_of(s_wq);
    const int lane = threadIdx.x & 31;
    const unsigned int n_paths = F.path_count[0];
    unsigned int n_refl = 0, n_sec = 0;
    TravCount tc = {0u, 0u, 0u};
    TravCount* const tcp = F.count_trav ? &tc : nullptr;      // traversal accounting only on request (vkp_set_profiling(ctx, 2))
    while (true) {
        unsigned int slot = 0u;
        if (lane == 0) slot = atomicAdd(&F.path_count[1], 1u);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot >= n_paths) break;
        const size_t i = (size_t)F.chain_pix[slot];
        const int py = (int)(i / (size_t)F.p.width), pxl = (int)(i - (size_t)py * F.p.width);
        const float4 pos4 = unpack_half4(F.position[i]);
        float4* rec = F.chain_buf + (size_t)slot * VKP_CHAIN_MAX;
        const vkp_material& mat = material_by_float_index(S, pos4.w);
        const int shadingMode = (int)mat.shadingMetallicRoughness[0];
        const float4 nrm4 = unpack_half4(F.normal[i]);
        const float3 position = xyz(pos4);
        const float3 N = normalize3(xyz(nrm4) * 2.0f - f3s(1.0f));
        const float3 camPosition = f3(F.cam.inv_view[12], F.cam.inv_view[13], F.cam.inv_view[14]);
        const float3 V = normalize3(camPosition - position);
        float3 origin = f3s(0.0f), direction = f3s(0.0f);
        if (S.n_lights > 0) first_bounce(shadingMode, mat, position, N, V, origin, direction);   // upstream sets them inside the light loop
        Payload prd;
        prd.seed = pixel_seed(F, i, pxl, py);
        int n = 0;
        for (int depth = 0; depth < F.p.max_bounces; depth++) {
            float3 ro = origin + direction * 1e-2f;
            HitRec h; n_refl++;
            const bool hit = warp_trace<false>(S, Q, ro, direction, 0.001f, 1000.0f, h, tcp);
            if (hit) shade_hit<false, 2>(S, nullptr, F.p, ro, direction, h, prd, n_sec, tcp, &Q); else shade_miss(S, direction, prd);
            if (lane == 0) rec[n] = make_float4(prd.color.x, prd.color.y, prd.color.z, 0.0f);
            n++;
            if (!hit || !(prd.scatter > 0.0f)) break;
            origin = prd.origin; direction = prd.direction;
        }
        if (lane == 0) { F.chain_len[slot] = n; if (F.chain_long) F.chain_long[i] = (uint8_t)n; }
    }
    add_counter(F.counters, 2, lane == 0 ? n_refl : 0u);       // the rays are the warp's: count them once
    add_counter(F.counters, 3, lane == 0 ? n_sec : 0u);
    if (F.count_trav) add_trav(F.counters, tc);
}

// ---- staged chain (mode 4): the FIRST bounce of every path as a wavefront ---------------------------------------------------------
// 8 % of the 1080p default frame's pixels start a path and 87 % of all reflection rays are first bounces.  One thread per path
// (k_chain) runs them at the register count of the hit shader -- 12 resident warps per SM, three waves of latency-bound traversals --
// and that kernel, not k_shadow, set the length of the overlapped stage.  Here the first bounce is split at its two trace calls so
// that each traversal runs in a lean kernel at full occupancy, and what continues is compacted before it goes on:
//   k_chain_first    one thread per path: first_bounce() + closest hit                -> path record (ray, hit)
//   k_chain_shadow0  one thread per (path, light): hybridHit.rchit:99-131's any-hit ray -> blocked-light mask of the record
//   k_chain_shade0   one thread per path: hybridHit.rchit / miss.rmiss with the mask    -> colour 0; paths that scatter are
//                    appended to a second queue (ballot + popc prefix, one atomic per warp)
//   k_chain_rest     one thread per continuing path (5 % of the first bounces): bounces 1 .. max_bounces-1 as in k_chain
// Records live in path_q[0] (by slot) and path_q[1] (continuing paths); path_count[2] counts the latter.  Same arithmetic in the
// same order as k_chain: every stored colour is bit-identical.
//   record.a = ray origin (already offset), t of the hit or -1     record.b = direction, u     record.c = v, instance, primitive, mask
__device__ __forceinline__ void chain_first_paths(const DScene& S, const DFrame& F, uint32_t* __restrict__ sstack, unsigned int k0, unsigned int stride,
                                                  unsigned int& n_refl, TravCount* tcp)
{
    const unsigned int n_paths = staged_count(F);
    for (unsigned int k = k0; k < n_paths; k += stride) {
        const unsigned int slot = staged_slot(F, k);
        const size_t i = (size_t)F.chain_pix[slot];
        float3 ro, direction;
        chain_first_ray(S, F, i, ro, direction);
        HitRec h; n_refl++;
        const bool hit = trace_ray<false>(S, sstack, ro, direction, 0.001f, 1000.0f, h, nullptr, tcp);
        PathItem* rec = F.path_q[0] + slot;
        rec->a = make_float4(ro.x, ro.y, ro.z, hit ? h.t : -1.0f);
        rec->b = make_float4(direction.x, direction.y, direction.z, h.u);
        rec->c = make_float4(h.v, __int_as_float(h.inst), __int_as_float(h.prim), __uint_as_float(0u));
        if (F.first_blocks) st_release_gpu(reinterpret_cast<uint32_t*>(&rec->d.w), F.chain_tag);      // schedule 8: "record complete" for the blocks further down the same grid
    }
}

__global__ void __launch_bounds__(VKP_TRAV_BLOCK, VKP_LB_CHAIN_TRACE) k_chain_first(const __grid_constant__ DScene S, const __grid_constant__ DFrame F)
{
    __shared__ uint32_t s_trav[VKP_TRAV_SMEM_WORDS];
    uint32_t* const sstack = s_trav + threadIdx.x;
    unsigned int n_refl = 0;
    TravCount tc = {0u, 0u, 0u};
    TravCount* const tcp = F.count_trav ? &tc : nullptr;
    chain_first_paths(S, F, sstack, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, n_refl, tcp);
    add_counter(F.counters, 2, n_refl);
    if (F.count_trav) add_trav(F.counters, tc);
}

// world position, shading normal and shading mode of a hit, as shade_hit computes them
__device__ __forceinline__ void hit_normal_pos(const DScene& S, float3 rayO, float3 rayD, const HitRec& h, float3& N, float3& worldPos, int& shadingMode)
{
    const DInstance& I = S.instances[h.inst];
    const float b0 = 1.0f - h.u - h.v, b1 = h.u, b2 = h.v;
    const uint32_t* ind = I.indices + 3u * (uint32_t)h.prim;
    const vkp_vertex& v0 = I.vertices[__ldg(ind)]; const vkp_vertex& v1 = I.vertices[__ldg(ind + 1)]; const vkp_vertex& v2 = I.vertices[__ldg(ind + 2)];
    const float3 normal = f3(v0.normal[0], v0.normal[1], v0.normal[2]) * b0 + f3(v1.normal[0], v1.normal[1], v1.normal[2]) * b1 +
                          f3(v2.normal[0], v2.normal[1], v2.normal[2]) * b2;
    N = normalize3(xform_normal_by_inverse(I.inv_model, normal));
    worldPos = rayO + rayD * h.t;
    const vkp_material& mat = S.materials[max(0, min(I.material_id, S.n_materials - 1))];
    shadingMode = (int)mat.shadingMetallicRoughness[0];
}

// hybridHit.rchit:99-131 for ONE light of ONE hit: a diffuse hit facing a light within its range traces one any-hit ray towards it.
// seed0 = the path's RNG state when the hit shader starts; with the advancing RNG the sample of light l starts from the state the
// lights before it leave behind (two LCG steps per light that traces).  Returns "traced and blocked".
__device__ __forceinline__ bool secondary_blocked(const DScene& S, const DFrame& F, uint32_t* __restrict__ sstack, float3 rayO, float3 rayD, const HitRec& h,
                                                  int l, uint32_t seed0, unsigned int& n_sec, TravCount* tcp)
{
    float3 N, worldPos; int shadingMode;
    hit_normal_pos(S, rayO, rayD, h, N, worldPos, shadingMode);
    if (shadingMode != 0) return false;                               // mirror / glass hits trace no shadow rays (hybridHit.rchit:133-183)
    const vkp_light light = S.lights[l];
    const LightGeom g = light_geom(light, worldPos, N);
    if (!(g.NdotL > 0.0f) || !(g.dist < g.maxd)) return false;
    const bool advance = F.p.faithful_rng == 0;
    uint32_t seed = seed0;
    if (advance) {
        for (int k = 0; k < l; k++) {
            const LightGeom gk = light_geom(S.lights[k], worldPos, N);
            if (gk.NdotL > 0.0f && gk.dist < gk.maxd) { seed = 1664525u * seed + 1013904223u; seed = 1664525u * seed + 1013904223u; }
        }
    }
    const float3 dir = sample_disk(light, worldPos, g.L, seed, advance);
    HitRec sh; n_sec++;
    return trace_ray<true>(S, sstack, worldPos + dir * 1e-2f, dir, 0.001f, g.dist + 1.0f, sh, nullptr, tcp);
}

__device__ __forceinline__ void chain_shadow0_paths(const DScene& S, const DFrame& F, uint32_t* __restrict__ sstack, int l, unsigned int k0, unsigned int stride,
                                                    unsigned int& n_sec, TravCount* tcp, bool wait_ready)
{
    const unsigned int n_paths = staged_count(F);
    for (unsigned int k = k0; k < n_paths; k += stride) {
        const unsigned int slot = staged_slot(F, k);
        PathItem* rec = F.path_q[0] + slot;
        if (wait_ready) while (ld_acquire_gpu(reinterpret_cast<const uint32_t*>(&rec->d.w)) != F.chain_tag) __nanosleep(100);
        const float4 a = rec->a;
        if (a.w < 0.0f) continue;                                       // the first ray missed: sky
        const float4 b = rec->b, c = rec->c;
        HitRec h; h.t = a.w; h.u = b.w; h.v = c.x; h.inst = __float_as_int(c.y); h.prim = __float_as_int(c.z);
        const size_t i = (size_t)F.chain_pix[slot];
        const int py = (int)(i / (size_t)F.p.width), pxl = (int)(i - (size_t)py * F.p.width);
        if (secondary_blocked(S, F, sstack, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), h, l, pixel_seed(F, i, pxl, py), n_sec, tcp))
            atomicOr(reinterpret_cast<unsigned int*>(&rec->c.w), 1u << l);
    }
}

// blockIdx.y = light: one thread per (path, light)
__global__ void __launch_bounds__(VKP_TRAV_BLOCK, VKP_LB_CHAIN_TRACE) k_chain_shadow0(const __grid_constant__ DScene S, const __grid_constant__ DFrame F)
{
    __shared__ uint32_t s_trav[VKP_TRAV_SMEM_WORDS];
    uint32_t* const sstack = s_trav + threadIdx.x;
    unsigned int n_sec = 0;
    TravCount tc = {0u, 0u, 0u};
    TravCount* const tcp = F.count_trav ? &tc : nullptr;
    chain_shadow0_paths(S, F, sstack, (int)blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, n_sec, tcp, false);
    add_counter(F.counters, 3, n_sec);
    if (F.count_trav) add_trav(F.counters, tc);
}

// ---- the long paths of schedule 7, FOUR LANES per path -------------------------------------------------------------------------------
// A path that goes on for 8 bounces is 8 x (one closest-hit ray + one any-hit ray per light + the hit shader), all dependent: ~230 us for
// the longest path of the 1080p default frame, whoever else is on the GPU.  The per-light shadow rays of a hit are independent of each
// other (hybridHit.rchit:99-131), so here a quad of lanes walks one path: all four trace the reflection ray and run the hit shader
// redundantly (one instruction stream, broadcast loads), and in between lane q traces the shadow ray of light q, q + 4, ...; the quad
// exchanges the answers with a ballot and the shader takes them as its precomputed mask (shade_hit<PRE>).  Only the few long paths run
// like this (4 x the threads); every stored colour is bit-identical to k_chain's.  A/B knob VKP_CHAIN_LONG_QUAD=1, off by default.
__global__ void __launch_bounds__(VKP_TRAV_BLOCK, VKP_LB_CHAIN) k_chain_quad(const __grid_constant__ DScene S, const __grid_constant__ DFrame F)
{
    __shared__ uint32_t s_trav[VKP_TRAV_SMEM_WORDS];
    uint32_t* const sstack = s_trav + threadIdx.x;
    const unsigned int full = 0xffffffffu;
    const int lane = (int)(threadIdx.x & 31u), q = lane & 3, quad = lane >> 2;
    // the very long paths of the split (chain_vl_list, head path_count[6]) -- or, with VKP_CHAIN_LONG_QUAD=1, the whole long list
    const bool vlist = F.chain_vl_list != nullptr && !F.chain_quad_all;
    const unsigned int n_paths = vlist ? F.path_count[5] : F.path_count[0];
    unsigned int n_refl = 0, n_sec = 0, n_unused = 0;
    TravCount tc = {0u, 0u, 0u};
    TravCount* const tcp = F.count_trav ? &tc : nullptr;
    while (true) {
        unsigned int base = 0u;
        if (lane == 0) base = atomicAdd(&F.path_count[vlist ? 6 : 1], 8u);
        base = __shfl_sync(full, base, 0);
        if (base >= n_paths) break;
        const unsigned int item = base + (unsigned int)quad;
        bool active = item < n_paths;
        const unsigned int slot = !active ? 0u : vlist ? (unsigned int)F.chain_vl_list[item] : item;
        size_t i = 0;
        float3 origin = f3s(0.0f), direction = f3s(0.0f), ro = f3s(0.0f);
        Payload prd; prd.seed = 0u;
        int n = 0;
        if (active) {
            i = (size_t)F.chain_pix[slot];
            const int py = (int)(i / (size_t)F.p.width), pxl = (int)(i - (size_t)py * F.p.width);
            chain_first_ray(S, F, i, ro, direction);
            prd.seed = pixel_seed(F, i, pxl, py);
            active = F.p.max_bounces > 0;
        }
        while (__any_sync(full, active)) {
            HitRec h; h.t = -1.0f; h.u = 0.0f; h.v = 0.0f; h.inst = -1; h.prim = -1;
            bool hit = false;
            if (active) {
                if (n > 0) ro = origin + direction * 1e-2f;
                if (q == 0) n_refl++;
                hit = trace_ray<false>(S, sstack, ro, direction, 0.001f, 1000.0f, h, nullptr, q == 0 ? tcp : nullptr);
            }
            uint32_t pre = 0u;
            for (int l0 = 0; l0 < S.n_lights; l0 += 4) {
                const int l = l0 + q;
                bool blocked = false;
                if (active && hit && l < S.n_lights) blocked = secondary_blocked(S, F, sstack, ro, direction, h, l, prd.seed, n_sec, tcp);
                const unsigned int bal = __ballot_sync(full, blocked);
                pre |= ((bal >> (quad * 4)) & 0xfu) << l0;
            }
            if (active) {
                if (hit) shade_hit<false, 0, true>(S, nullptr, F.p, ro, direction, h, prd, n_unused, nullptr, nullptr, pre); else shade_miss(S, direction, prd);
                if (q == 0) F.chain_buf[(size_t)slot * VKP_CHAIN_MAX + n] = make_float4(prd.color.x, prd.color.y, prd.color.z, 0.0f);
                n++;
                if (!hit || !(prd.scatter > 0.0f) || n >= F.p.max_bounces) {
                    if (q == 0) { F.chain_len[slot] = n; if (F.chain_long) F.chain_long[i] = (uint8_t)n; }
                    active = false;
                } else { origin = prd.origin; direction = prd.direction; }
            }
        }
    }
    add_counter(F.counters, 2, n_refl);
    add_counter(F.counters, 3, n_sec);
    if (F.count_trav) add_trav(F.counters, tc);
}

__global__ void __launch_bounds__(VKP_TRAV_BLOCK, VKP_LB_CHAIN) k_chain_shade0(const __grid_constant__ DScene S, const __grid_constant__ DFrame F)
{
    const unsigned int n_paths = staged_count(F);
    const unsigned int n_round = (n_paths + 31u) & ~31u;              // whole warps stay in the loop so that the ballots are convergent
    unsigned int n_sec = 0;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_round; k += gridDim.x * blockDim.x) {
        bool again = false;
        PathItem item;
        if (k < n_paths) {
            const unsigned int slot = staged_slot(F, k);
            const PathItem* rec = F.path_q[0] + slot;
            const float4 a = rec->a, b = rec->b, c = rec->c;
            const size_t i = (size_t)F.chain_pix[slot];
            const int py = (int)(i / (size_t)F.p.width), pxl = (int)(i - (size_t)py * F.p.width);
            Payload prd;
            prd.seed = pixel_seed(F, i, pxl, py);
            const float3 ro = f3(a.x, a.y, a.z), direction = f3(b.x, b.y, b.z);
            const bool hit = !(a.w < 0.0f);
            if (hit) {
                HitRec h; h.t = a.w; h.u = b.w; h.v = c.x; h.inst = __float_as_int(c.y); h.prim = __float_as_int(c.z);
                shade_hit<false, 0, true>(S, nullptr, F.p, ro, direction, h, prd, n_sec, nullptr, nullptr, __float_as_uint(c.w));
            } else shade_miss(S, direction, prd);
            F.chain_buf[(size_t)slot * VKP_CHAIN_MAX] = make_float4(prd.color.x, prd.color.y, prd.color.z, 0.0f);
            again = hit && (prd.scatter > 0.0f) && F.p.max_bounces > 1;
            if (F.chain_long) F.chain_long[i] = again ? 2 : 1;          // (k_chain_rest writes the real count of the paths that go on)
            if (!again) F.chain_len[slot] = 1;
            else {
                item.a = make_float4(prd.origin.x, prd.origin.y, prd.origin.z, __uint_as_float(slot));
                item.b = make_float4(prd.direction.x, prd.direction.y, prd.direction.z, __uint_as_float(prd.seed));
            }
        }
        // the continuing paths, compacted (only a and b of the item are used)
        const unsigned mask = __ballot_sync(0xffffffffu, again);
        if (mask) {
            const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
            unsigned base = 0;
            if (lane == leader) base = atomicAdd(&F.path_count[2], (unsigned)__popc(mask));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (again) { PathItem* q = F.path_q[1] + base + __popc(mask & ((1u << lane) - 1u)); q->a = item.a; q->b = item.b; }
        }
    }
}

__global__ void __launch_bounds__(VKP_TRAV_BLOCK, VKP_LB_CHAIN) k_chain_rest(const __grid_constant__ DScene S, const __grid_constant__ DFrame F)
{
    __shared__ uint32_t s_trav[VKP_TRAV_SMEM_WORDS];
    uint32_t* const sstack = s_trav + threadIdx.x;
    const unsigned int n_cont = F.path_count[2];
    unsigned int n_refl = 0, n_sec = 0;
    TravCount tc = {0u, 0u, 0u};
    TravCount* const tcp = F.count_trav ? &tc : nullptr;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_cont; k += gridDim.x * blockDim.x) {
        const float4 qa = F.path_q[1][k].a, qb = F.path_q[1][k].b;
        const unsigned int slot = __float_as_uint(qa.w);
        float4* rec = F.chain_buf + (size_t)slot * VKP_CHAIN_MAX;
        float3 origin = f3(qa.x, qa.y, qa.z), direction = f3(qb.x, qb.y, qb.z);
        Payload prd; prd.seed = __float_as_uint(qb.w);
        int n = 1;
        for (int depth = 1; depth < F.p.max_bounces; depth++) {
            float3 ro = origin + direction * 1e-2f;
            HitRec h; n_refl++;
            const bool hit = trace_ray<false>(S, sstack, ro, direction, 0.001f, 1000.0f, h, nullptr, tcp);
            if (hit) shade_hit<false>(S, sstack, F.p, ro, direction, h, prd, n_sec, tcp); else shade_miss(S, direction, prd);
            rec[n] = make_float4(prd.color.x, prd.color.y, prd.color.z, 0.0f);
            n++;
            if (!hit || !(prd.scatter > 0.0f)) break;
            origin = prd.origin; direction = prd.direction;
        }
        F.chain_len[slot] = n;
        if (F.chain_long) F.chain_long[(size_t)F.chain_pix[slot]] = (uint8_t)n;
    }
    add_counter(F.counters, 2, n_refl);
    add_counter(F.counters, 3, n_sec);
    if (F.count_trav) add_trav(F.counters, tc);
}

