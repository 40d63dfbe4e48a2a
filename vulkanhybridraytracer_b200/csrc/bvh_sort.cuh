// bvh_sort.cuh -- hand-written device-wide primitives of the BVH build: LSD radix sort of (Morton key, primitive id)
// pairs and exclusive prefix sums.  No CUB / Thrust.
//
// Radix sort: 8-bit digits, one pass = histogram -> scan -> stable scatter.  The scatter ranks keys at warp level:
// __match_any_sync groups the lanes of a warp by digit, the lane's rank inside its group is a popcount of the lower
// lanes, and per-warp digit counts are chained through shared memory so that the order inside a tile -- and therefore
// the whole sort -- is stable.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vkpsort {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;     // keys per CTA
constexpr int RS_WARPS = RS_THREADS / 32;

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint32_t* __restrict__ keys, uint32_t n, int shift, uint32_t nb, uint32_t* __restrict__ ghist)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        uint32_t i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    ghist[threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];     // digit-major: one exclusive scan gives every (digit, block) base
}

// single-CTA exclusive scan (the histogram table has 256 * nb entries; nb <= a few thousand)
template <typename T>
__global__ void __launch_bounds__(1024) k_scan_single(const T* __restrict__ in, T* __restrict__ out, uint32_t n)
{
    __shared__ T warp_sums[32];
    const uint32_t per = (n + 1023u) / 1024u;
    const uint32_t lo = threadIdx.x * per, hi = min(lo + per, n);
    T sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += in[i];
    // block exclusive scan of the per-thread sums
    T v = sum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { T t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    if (lane == 31) warp_sums[warp] = v;
    __syncthreads();
    if (warp == 0) {
        T w = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { T t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
        warp_sums[lane] = w;
    }
    __syncthreads();
    T run = (v - sum) + (warp ? warp_sums[warp - 1] : (T)0);
    for (uint32_t i = lo; i < hi; i++) { T t = in[i]; out[i] = run; run += t; }
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
                                                           uint32_t* __restrict__ vals_out, uint32_t n, int shift, uint32_t nb, const uint32_t* __restrict__ gscan)
{
    __shared__ uint32_t gbase[256];               // global offset of this CTA's first key of each digit
    __shared__ uint32_t run[256];                 // keys of each digit already placed by this CTA
    __shared__ uint32_t cnt[RS_WARPS][256];       // per-round: digit counts per warp -> exclusive offsets
    gbase[threadIdx.x] = gscan[threadIdx.x * nb + blockIdx.x];
    run[threadIdx.x] = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t base = blockIdx.x * RS_TILE;
    for (int r = 0; r < RS_ITEMS; r++) {
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) cnt[w][threadIdx.x] = 0;
        __syncthreads();
        const uint32_t i = base + r * RS_THREADS + threadIdx.x;
        const bool valid = i < n;
        uint32_t key = 0, val = 0, digit = 256u + (uint32_t)lane;      // invalid lanes: a digit nobody shares
        if (valid) { key = keys_in[i]; val = vals_in[i]; digit = (key >> shift) & 255u; }
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        const uint32_t rank = (uint32_t)__popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) cnt[warp][digit] = (uint32_t)__popc(peers);
        __syncthreads();
        {   // thread t owns digit t: chain the warps in order (stability) on top of what earlier rounds placed
            uint32_t acc = run[threadIdx.x];
#pragma unroll
            for (int w = 0; w < RS_WARPS; w++) { uint32_t c = cnt[w][threadIdx.x]; cnt[w][threadIdx.x] = acc; acc += c; }
            run[threadIdx.x] = acc;
        }
        __syncthreads();
        if (valid) {
            const uint32_t pos = gbase[digit] + cnt[warp][digit] + rank;
            keys_out[pos] = key; vals_out[pos] = val;
        }
        __syncthreads();
    }
}

// ---- generic exclusive scan over n elements: tile-local scan + scan of tile sums + add ------------------------------------
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 4;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

template <typename T>
__global__ void __launch_bounds__(SC_THREADS) k_scan_tiles(const T* __restrict__ in, T* __restrict__ out, uint32_t n, T* __restrict__ tile_sums)
{
    __shared__ T warp_sums[SC_THREADS / 32];
    const uint32_t base = blockIdx.x * SC_TILE + threadIdx.x * SC_ITEMS;
    T x[SC_ITEMS];
    T sum = 0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) { x[k] = (base + k < n) ? in[base + k] : (T)0; sum += x[k]; }
    T v = sum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { T t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    if (lane == 31) warp_sums[warp] = v;
    __syncthreads();
    T woff = 0;
    for (int w = 0; w < warp; w++) woff += warp_sums[w];
    T run = woff + (v - sum);
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) { if (base + k < n) out[base + k] = run; run += x[k]; }
    if (threadIdx.x == SC_THREADS - 1) tile_sums[blockIdx.x] = run;
}

template <typename T>
__global__ void __launch_bounds__(SC_THREADS) k_scan_add(T* __restrict__ out, uint32_t n, const T* __restrict__ tile_offsets)
{
    const T off = tile_offsets[blockIdx.x];
    const uint32_t base = blockIdx.x * SC_TILE + threadIdx.x * SC_ITEMS;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) if (base + k < n) out[base + k] += off;
}

inline size_t scan_scratch_bytes(uint32_t n, size_t elem) { return ((size_t)(n + SC_TILE - 1) / SC_TILE + 1) * elem * 2; }

// exclusive sum of in[0..n) into out (in != out).  `scratch` >= scan_scratch_bytes(n, sizeof(T)).  3 launches.
template <typename T>
inline cudaError_t exclusive_scan(cudaStream_t s, const T* in, T* out, uint32_t n, void* scratch, int* launches)
{
    if (n == 0) return cudaSuccess;
    const uint32_t tiles = (n + SC_TILE - 1) / SC_TILE;
    T* sums = reinterpret_cast<T*>(scratch);
    T* offs = sums + tiles + 1;
    k_scan_tiles<T><<<tiles, SC_THREADS, 0, s>>>(in, out, n, sums);
    k_scan_single<T><<<1, 1024, 0, s>>>(sums, offs, tiles);
    k_scan_add<T><<<tiles, SC_THREADS, 0, s>>>(out, n, offs);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

inline size_t sort_scratch_bytes(uint32_t n) { return (size_t)256 * ((n + RS_TILE - 1) / RS_TILE) * sizeof(uint32_t) * 2; }

// stable LSD radix sort of (key, value) pairs on bits [0, end_bit).  Result lands in keys_out / vals_out; the inputs are
// clobbered (ping-pong).  `scratch` >= sort_scratch_bytes(n).
inline cudaError_t sort_pairs(cudaStream_t s, uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out, uint32_t n, int end_bit,
                              void* scratch, int* launches)
{
    if (n == 0) return cudaSuccess;
    const uint32_t nb = (n + RS_TILE - 1) / RS_TILE;
    uint32_t* ghist = reinterpret_cast<uint32_t*>(scratch);
    uint32_t* gscan = ghist + (size_t)256 * nb;
    const int passes = (end_bit + 7) / 8;
    uint32_t *ki = keys_in, *vi = vals_in, *ko = keys_out, *vo = vals_out;
    for (int p = 0; p < passes; p++) {
        const int shift = 8 * p;
        k_rs_hist<<<nb, RS_THREADS, 0, s>>>(ki, n, shift, nb, ghist);
        k_scan_single<uint32_t><<<1, 1024, 0, s>>>(ghist, gscan, 256u * nb);
        k_rs_scatter<<<nb, RS_THREADS, 0, s>>>(ki, vi, ko, vo, n, shift, nb, gscan);
        uint32_t* t = ki; ki = ko; ko = t; t = vi; vi = vo; vo = t;
        if (launches) *launches += 3;
    }
    // ki / vi now name the buffers holding the sorted data (the inputs again after an even number of passes)
    if (ki != keys_out) {
        cudaError_t e = cudaMemcpyAsync(keys_out, ki, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return e;
        e = cudaMemcpyAsync(vals_out, vi, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

} // namespace vkpsort
