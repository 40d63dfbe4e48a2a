// bvh_build.cu -- acceleration-structure build on the GPU: what vkCmdBuildAccelerationStructuresKHR does for the
// reference (renderer.cpp:1747 buildBlas, :1853 buildTlas / MODE_UPDATE refit).
//
// Morton-code LBVH (Karras 2012): primitive boxes -> centroid bounds -> 30-bit Morton keys -> radix sort (own warp-level
// LSD sort, bvh_sort.cuh) ->
// binary radix tree (one thread per internal node) -> bottom-up AABB fit with atomic arrival flags ->
// collapse sub-trees of <= leaf_max primitives into leaves and emit the 32-byte sibling-pair node array
// described in vkp_bvh.cuh.  Refit re-runs only fit + emit on the stored topology.
#include "vkp_internal.h"

#include "bvh_sort.cuh"
#include <math_constants.h>
#include <cstdlib>

namespace {

// loop state of the device-driven PLOC rounds (below).  The kernels that consume a finished hierarchy take a `gate`: queued behind
// rounds that turned out too few (n_cur != 1) they must not walk a half-built forest -- its parent links are stale memory.
struct PlocState { int n_cur, node_base, cur, rounds; };
__device__ __forceinline__ bool gate_closed(const PlocState* gate) { return gate && gate->n_cur != 1; }

__device__ __forceinline__ int float_to_ordered(float f)
{
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// ---- primitive boxes -----------------------------------------------------------------------------------------
// BLAS input of Node::node_to_geometry (vk_mesh.cpp:424-470): R32G32B32 positions at stride sizeof(Vertex) = 44,
// uint32 indices starting at firstIndex
__global__ void k_triangle_boxes(const vkp_vertex* __restrict__ verts, const uint32_t* __restrict__ indices, uint32_t n, float* __restrict__ boxes)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t i0 = indices[3 * i], i1 = indices[3 * i + 1], i2 = indices[3 * i + 2];
    const float* a = verts[i0].position; const float* b = verts[i1].position; const float* c = verts[i2].position;
    float* o = boxes + 6 * (size_t)i;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        o[k] = fminf(fminf(a[k], b[k]), c[k]);
        o[3 + k] = fmaxf(fmaxf(a[k], b[k]), c[k]);
    }
}

// TLAS input of Renderer::object_to_instance (renderer.cpp:1895-1916): world box of the instance = box of the 8
// transformed corners of its BLAS root, padded by 1e-5 relative (the object-space hit is exact inside the
// object box, the world box only to rounding)
__global__ void k_instance_boxes(const DInstance* __restrict__ inst, uint32_t n, float* __restrict__ boxes)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DInstance& I = inst[i];
    const float4* root = reinterpret_cast<const float4*>(I.nodes);
    float4 lo = root[0], hi = root[1];
    float mn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, mx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    for (int c = 0; c < 8; c++) {
        float3 p = make_float3((c & 1) ? hi.x : lo.x, (c & 2) ? hi.y : lo.y, (c & 4) ? hi.z : lo.z);
        float3 w = xform_point(I.model, p);
        mn[0] = fminf(mn[0], w.x); mn[1] = fminf(mn[1], w.y); mn[2] = fminf(mn[2], w.z);
        mx[0] = fmaxf(mx[0], w.x); mx[1] = fmaxf(mx[1], w.y); mx[2] = fmaxf(mx[2], w.z);
    }
    float* o = boxes + 6 * (size_t)i;
    for (int k = 0; k < 3; k++) {
        float e = 1e-5f * fmaxf(fmaxf(fabsf(mn[k]), fabsf(mx[k])), 1e-3f);
        o[k] = mn[k] - e; o[3 + k] = mx[k] + e;
    }
}

// ---- centroid bounds + Morton keys ------------------------------------------------------------------------
__global__ void k_init_bounds(int* b)
{
    if (threadIdx.x < 3) b[threadIdx.x] = 0x7fffffff;
    else if (threadIdx.x < 6) b[threadIdx.x] = (int)0x80000000;
}

__global__ void k_centroid_bounds(const float* __restrict__ boxes, uint32_t n, int* __restrict__ bounds)
{
    float mn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, mx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* b = boxes + 6 * (size_t)i;
#pragma unroll
        for (int k = 0; k < 3; k++) { float c = 0.5f * (b[k] + b[3 + k]); mn[k] = fminf(mn[k], c); mx[k] = fmaxf(mx[k], c); }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        for (int off = 16; off; off >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], off));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], off));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) { atomicMin(&bounds[k], float_to_ordered(mn[k])); atomicMax(&bounds[3 + k], float_to_ordered(mx[k])); }
    }
}

__device__ __forceinline__ uint32_t expand_bits10(uint32_t v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void k_morton(const float* __restrict__ boxes, uint32_t n, const int* __restrict__ bounds, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* b = boxes + 6 * (size_t)i;
    uint32_t q[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float lo = ordered_to_float(bounds[k]), hi = ordered_to_float(bounds[3 + k]);
        float c = 0.5f * (b[k] + b[3 + k]);
        float ext = hi - lo;
        float t = ext > 0.0f ? (c - lo) / ext : 0.0f;
        t = fminf(fmaxf(t * 1024.0f, 0.0f), 1023.0f);
        q[k] = (uint32_t)t;
    }
    keys[i] = (expand_bits10(q[0]) << 2) | (expand_bits10(q[1]) << 1) | expand_bits10(q[2]);
    vals[i] = i;
}

// ---- Karras binary radix tree -------------------------------------------------------------------------------
__device__ __forceinline__ int delta(const uint32_t* __restrict__ keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    uint32_t a = keys[i], b = keys[j];
    if (a == b) return 32 + __clz((uint32_t)i ^ (uint32_t)j);   // tie-break on the index: balanced split of duplicates
    return __clz(a ^ b);
}

__global__ void k_hierarchy(const uint32_t* __restrict__ keys, int n, int2* __restrict__ child, int* __restrict__ parent, int2* __restrict__ range)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int first = min(i, j), last = max(i, j);
    int left = (first == gamma) ? ~gamma : gamma;
    int right = (last == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    child[i] = make_int2(left, right);
    range[i] = make_int2(first, last);
    // parents: internals in [0, n-1), leaves in [n-1, 2n-1)
    if (left >= 0) parent[left] = i; else parent[(n - 1) + (~left)] = i;
    if (right >= 0) parent[right] = i; else parent[(n - 1) + (~right)] = i;
    if (i == 0) parent[0] = -1;
}

// ---- bottom-up fit ------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_box(const float* __restrict__ boxes, const uint32_t* __restrict__ items, const float4* ibox,
                                         int c, float4& lo, float4& hi)
{
    if (c < 0) {
        const float* b = boxes + 6 * (size_t)items[~c];
        lo = make_float4(b[0], b[1], b[2], 0.f); hi = make_float4(b[3], b[4], b[5], 0.f);
    } else {
        lo = __ldcg(&ibox[2 * (size_t)c]); hi = __ldcg(&ibox[2 * (size_t)c + 1]);
    }
}

__global__ void k_fit(const float* __restrict__ boxes, const uint32_t* __restrict__ items, int n, const int2* __restrict__ child,
                      const int* __restrict__ parent, float4* ibox, int* flags, int* height_out, int root, const PlocState* __restrict__ gate)
{
    int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n || gate_closed(gate)) return;
    int p = parent[(n - 1) + leaf];
    int h = 1;
    while (p >= 0) {
        __threadfence();
        int old = atomicAdd(&flags[p], 1);
        if (old == 0) return;                // the sibling sub-tree is not finished yet
        int2 c = child[p];
        float4 l0, h0, l1, h1;
        load_box(boxes, items, ibox, c.x, l0, h0);
        load_box(boxes, items, ibox, c.y, l1, h1);
        // heights ride in the w components
        float hl = c.x < 0 ? 0.f : l0.w, hr = c.y < 0 ? 0.f : l1.w;
        float hh = fmaxf(hl, hr) + 1.0f;
        float4 lo = make_float4(fminf(l0.x, l1.x), fminf(l0.y, l1.y), fminf(l0.z, l1.z), hh);
        float4 hi = make_float4(fmaxf(h0.x, h1.x), fmaxf(h0.y, h1.y), fmaxf(h0.z, h1.z), 0.f);
        __stcg(&ibox[2 * (size_t)p], lo); __stcg(&ibox[2 * (size_t)p + 1], hi);
        h = (int)hh;
        if (p == root) *height_out = h;
        p = parent[p];
    }
}

// ---- collapse + emit ----------------------------------------------------------------------------------------
__global__ void k_mark_keep(const int2* __restrict__ range, int n_internal, int leaf_max, uint32_t* __restrict__ keep)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_internal) return;
    int2 r = range[i];
    keep[i] = (r.y - r.x + 1) > leaf_max ? 1u : 0u;
}

__global__ void k_emit_nodes(const float* __restrict__ boxes, const uint32_t* __restrict__ items, int n, const int2* __restrict__ child,
                             const int2* __restrict__ range, const float4* __restrict__ ibox, const uint32_t* __restrict__ keep,
                             const uint32_t* __restrict__ rank, vkp_bvh_node* __restrict__ nodes, int root, int ref_items, const PlocState* __restrict__ gate)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n - 1 || gate_closed(gate)) return;
    if (k == root) {
        float4 lo = ibox[2 * (size_t)root], hi = ibox[2 * (size_t)root + 1];
        vkp_bvh_node r;
        r.bmin[0] = lo.x; r.bmin[1] = lo.y; r.bmin[2] = lo.z; r.bmax[0] = hi.x; r.bmax[1] = hi.y; r.bmax[2] = hi.z;
        r.ref = keep[root] ? (1u + 2u * rank[root]) : (0x80000000u | ((uint32_t)range[root].x << 3) | (uint32_t)(range[root].y - range[root].x));
        r.aux = (uint32_t)(range[root].y - range[root].x + 1);
        nodes[0] = r;
    }
    if (!keep[k]) return;
    uint32_t slot = 1u + 2u * rank[k];
    int2 c = child[k];
    int cc[2] = {c.x, c.y};
#pragma unroll
    for (int s = 0; s < 2; s++) {
        vkp_bvh_node nd;
        int ci = cc[s];
        if (ci < 0) {
            int pos = ~ci;
            const float* b = boxes + 6 * (size_t)items[pos];
            nd.bmin[0] = b[0]; nd.bmin[1] = b[1]; nd.bmin[2] = b[2]; nd.bmax[0] = b[3]; nd.bmax[1] = b[4]; nd.bmax[2] = b[5];
            // TLAS (ref_items): a single-instance leaf names the instance itself, saving traversal one dependent load
            nd.ref = 0x80000000u | ((ref_items ? items[pos] : (uint32_t)pos) << 3);
            nd.aux = 1u;
        } else {
            float4 lo = ibox[2 * (size_t)ci], hi = ibox[2 * (size_t)ci + 1];
            nd.bmin[0] = lo.x; nd.bmin[1] = lo.y; nd.bmin[2] = lo.z; nd.bmax[0] = hi.x; nd.bmax[1] = hi.y; nd.bmax[2] = hi.z;
            int2 r = range[ci];
            nd.ref = keep[ci] ? (1u + 2u * rank[ci]) : (0x80000000u | ((uint32_t)r.x << 3) | (uint32_t)(r.y - r.x));
            nd.aux = (uint32_t)(r.y - r.x + 1);
        }
        nodes[slot + s] = nd;
    }
}

__global__ void k_emit_single(const float* __restrict__ boxes, int n, vkp_bvh_node* __restrict__ nodes)
{
    vkp_bvh_node r;
    if (n == 1) {
        for (int k = 0; k < 3; k++) { r.bmin[k] = boxes[k]; r.bmax[k] = boxes[3 + k]; }
        r.ref = 0x80000000u; r.aux = 1u;
    } else {
        for (int k = 0; k < 3; k++) { r.bmin[k] = CUDART_INF_F; r.bmax[k] = -CUDART_INF_F; }
        r.ref = 0x80000000u; r.aux = 0u;
    }
    nodes[0] = r;
}

__global__ void k_emit_triangles(const vkp_vertex* __restrict__ verts, const uint32_t* __restrict__ indices, const uint32_t* __restrict__ items,
                                 uint32_t n, float4* __restrict__ tris, const PlocState* __restrict__ gate)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || gate_closed(gate)) return;
    uint32_t prim = items[i];
    uint32_t i0 = indices[3 * prim], i1 = indices[3 * prim + 1], i2 = indices[3 * prim + 2];
    const float* a = verts[i0].position; const float* b = verts[i1].position; const float* c = verts[i2].position;
    tris[3 * (size_t)i + 0] = make_float4(a[0], a[1], a[2], __int_as_float((int)prim));
    tris[3 * (size_t)i + 1] = make_float4(b[0], b[1], b[2], 0.f);
    tris[3 * (size_t)i + 2] = make_float4(c[0], c[1], c[2], 0.f);
}

// ---- PLOC: parallel locally-ordered clustering (Meister & Bittner 2018) --------------------------------------------
// Bottom-up agglomerative build over the Morton-sorted leaves: every cluster looks for the neighbour within +-R slots
// whose union has the smallest surface area, mutual nearest neighbours merge, the cluster list is compacted with a prefix
// sum, repeat until one cluster is left.  Tree quality is close to a full SAH sweep and far better than the Karras radix
// tree for long thin triangles (sphere poles) and grazing rays -- which is what sets the frame kernels' tail latency.
#ifndef PLOC_R
#define PLOC_R 16                   /* search radius; 8 / 32 / 64 measured: see DESIGN.md section 4c */
#endif
#define PLOC_BS 256

__device__ __forceinline__ float half_area(float4 lo, float4 hi)
{
    float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}

__global__ void k_ploc_init(const float* __restrict__ boxes, const uint32_t* __restrict__ items, int n, int* __restrict__ cid, float4* __restrict__ cbox, uint32_t* __restrict__ ccnt)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* b = boxes + 6 * (size_t)items[i];
    cid[i] = ~i;
    cbox[2 * (size_t)i] = make_float4(b[0], b[1], b[2], 0.f);
    cbox[2 * (size_t)i + 1] = make_float4(b[3], b[4], b[5], 0.f);
    ccnt[i] = 1u;
}

// ---- PLOC rounds driven from the device ---------------------------------------------------------------------------------------------
// The cluster count of the next round is only known once a round has run.  Round 1 read it back (two 8-byte copies and a stream
// synchronisation per round: 138 round trips for the default scene).  Here the loop state lives in device memory -- two copies,
// round r reads state[r & 1] and leaves state[(r + 1) & 1] -- every kernel of a round sizes its work from it (grid-stride over
// chunks; a launch that finds one cluster left returns at once), and the host queues a fixed number of rounds without ever waiting.
// A round is 4 launches: nearest neighbours, flags + tile-local prefix sums, one-CTA scan of the tile sums (which also writes the
// next state), merge + compaction.

__global__ void k_ploc_state_init(PlocState* st, int n)
{
    st[0].n_cur = n; st[0].node_base = 0; st[0].cur = 0; st[0].rounds = 0;
    st[1] = st[0];
}

__global__ void __launch_bounds__(PLOC_BS) k_ploc_nn_dev(const PlocState* __restrict__ st, const float4* __restrict__ box0, const float4* __restrict__ box1, int* __restrict__ nn)
{
    __shared__ float4 slo[PLOC_BS + 2 * PLOC_R], shi[PLOC_BS + 2 * PLOC_R];
    const int n = st->n_cur;
    if (n <= 1) return;
    const float4* __restrict__ cbox = st->cur ? box1 : box0;
    for (int chunk = blockIdx.x; chunk * PLOC_BS < n; chunk += gridDim.x) {
        const int base = chunk * PLOC_BS - PLOC_R;
        for (int t = threadIdx.x; t < PLOC_BS + 2 * PLOC_R; t += PLOC_BS) {
            int j = base + t;
            if (j >= 0 && j < n) { slo[t] = cbox[2 * (size_t)j]; shi[t] = cbox[2 * (size_t)j + 1]; }
        }
        __syncthreads();
        const int i = chunk * PLOC_BS + threadIdx.x;
        if (i < n) {
            const float4 lo = slo[threadIdx.x + PLOC_R], hi = shi[threadIdx.x + PLOC_R];
            float best = CUDART_INF_F; int bj = -1;
            for (int d = -PLOC_R; d <= PLOC_R; d++) {
                int j = i + d;
                if (d == 0 || j < 0 || j >= n) continue;
                float4 l2 = slo[threadIdx.x + PLOC_R + d], h2 = shi[threadIdx.x + PLOC_R + d];
                float4 ul = make_float4(fminf(lo.x, l2.x), fminf(lo.y, l2.y), fminf(lo.z, l2.z), 0.f);
                float4 uh = make_float4(fmaxf(hi.x, h2.x), fmaxf(hi.y, h2.y), fmaxf(hi.z, h2.z), 0.f);
                float a = half_area(ul, uh);
                if (a < best) { best = a; bj = j; }          // ties keep the smaller j: deterministic
            }
            nn[i] = bj;
        }
        __syncthreads();
    }
}

// high word = "i is the lower index of a mutual pair" (creates a node), low word = "i survives into the next list"
__device__ __forceinline__ unsigned long long ploc_flags_of(const int* __restrict__ nn, int i)
{
    const int j = nn[i];
    const bool mutual = j >= 0 && nn[j] == i;
    const bool lower = mutual && i < j, upper = mutual && i > j;
    return ((unsigned long long)(lower ? 1u : 0u) << 32) | (unsigned long long)(upper ? 0u : 1u);
}

constexpr int PLOC_SC_THREADS = 256, PLOC_SC_ITEMS = 4, PLOC_SC_TILE = PLOC_SC_THREADS * PLOC_SC_ITEMS;

__global__ void __launch_bounds__(PLOC_SC_THREADS) k_ploc_scan_tiles(const PlocState* __restrict__ st, const int* __restrict__ nn, unsigned long long* __restrict__ local,
                                                                    unsigned long long* __restrict__ tile_sums)
{
    __shared__ unsigned long long warp_sums[PLOC_SC_THREADS / 32];
    const int n = st->n_cur;
    if (n <= 1) return;
    const int tiles = (n + PLOC_SC_TILE - 1) / PLOC_SC_TILE;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int base = tile * PLOC_SC_TILE + threadIdx.x * PLOC_SC_ITEMS;
        unsigned long long x[PLOC_SC_ITEMS], sum = 0;
#pragma unroll
        for (int k = 0; k < PLOC_SC_ITEMS; k++) { x[k] = (base + k < n) ? ploc_flags_of(nn, base + k) : 0ull; sum += x[k]; }
        unsigned long long v = sum;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
        if (lane == 31) warp_sums[warp] = v;
        __syncthreads();
        unsigned long long woff = 0;
        for (int w = 0; w < warp; w++) woff += warp_sums[w];
        unsigned long long run = woff + (v - sum);
#pragma unroll
        for (int k = 0; k < PLOC_SC_ITEMS; k++) { if (base + k < n) local[base + k] = run; run += x[k]; }
        if (threadIdx.x == PLOC_SC_THREADS - 1) tile_sums[tile] = run;
        __syncthreads();
    }
}

// one CTA: exclusive scan of the tile sums -> tile offsets; the totals give the next round's state
__global__ void __launch_bounds__(1024) k_ploc_scan_state(const PlocState* __restrict__ st, PlocState* __restrict__ st_next, const unsigned long long* __restrict__ tile_sums,
                                                          unsigned long long* __restrict__ tile_offs)
{
    __shared__ unsigned long long warp_sums[32];
    const PlocState cur = *st;
    if (cur.n_cur <= 1) { if (threadIdx.x == 0) *st_next = cur; return; }
    const uint32_t n = (uint32_t)((cur.n_cur + PLOC_SC_TILE - 1) / PLOC_SC_TILE);
    const uint32_t per = (n + 1023u) / 1024u;
    const uint32_t lo = min(threadIdx.x * per, n), hi = min(lo + per, n);
    unsigned long long sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += tile_sums[i];
    unsigned long long v = sum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    if (lane == 31) warp_sums[warp] = v;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
        warp_sums[lane] = w;
    }
    __syncthreads();
    unsigned long long run = (v - sum) + (warp ? warp_sums[warp - 1] : 0ull);
    for (uint32_t i = lo; i < hi; i++) { unsigned long long t = tile_sums[i]; tile_offs[i] = run; run += t; }
    if (threadIdx.x == 1023) {
        const unsigned long long tot = warp_sums[31];            // high word: merges (new nodes), low word: survivors (next list)
        PlocState nx;
        nx.n_cur = (int)(tot & 0xffffffffull); nx.node_base = cur.node_base + (int)(tot >> 32); nx.cur = cur.cur ^ 1; nx.rounds = cur.rounds + 1;
        *st_next = nx;
    }
}

__global__ void k_ploc_merge_dev(const PlocState* __restrict__ st, const int* __restrict__ nn, const unsigned long long* __restrict__ local,
                                 const unsigned long long* __restrict__ tile_offs,
                                 int* __restrict__ cid0, float4* __restrict__ box0, uint32_t* __restrict__ cnt0,
                                 int* __restrict__ cid1, float4* __restrict__ box1, uint32_t* __restrict__ cnt1, int n_leaves,
                                 int2* __restrict__ child, int* __restrict__ parent, float4* __restrict__ ibox, uint32_t* __restrict__ icnt)
{
    const int n = st->n_cur;
    if (n <= 1) return;
    const int node_base = st->node_base;
    const bool flip = st->cur != 0;
    const int* __restrict__ cid_in = flip ? cid1 : cid0; const float4* __restrict__ cbox_in = flip ? box1 : box0; const uint32_t* __restrict__ ccnt_in = flip ? cnt1 : cnt0;
    int* __restrict__ cid_out = flip ? cid0 : cid1; float4* __restrict__ cbox_out = flip ? box0 : box1; uint32_t* __restrict__ ccnt_out = flip ? cnt0 : cnt1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long f = ploc_flags_of(nn, i);
        if (!(f & 1ull)) continue;                                 // the upper half of a merged pair disappears
        const unsigned long long sc = local[i] + tile_offs[i / PLOC_SC_TILE];
        const int slot = (int)(sc & 0xffffffffull);
        if (f >> 32) {
            const int j = nn[i];
            const int node = node_base + (int)(sc >> 32);
            const int a = cid_in[i], b = cid_in[j];
            float4 la = cbox_in[2 * (size_t)i], ha = cbox_in[2 * (size_t)i + 1], lb = cbox_in[2 * (size_t)j], hb = cbox_in[2 * (size_t)j + 1];
            float hgt = fmaxf(a < 0 ? 0.f : la.w, b < 0 ? 0.f : lb.w) + 1.0f;        // heights ride in lo.w like k_fit
            float4 lo = make_float4(fminf(la.x, lb.x), fminf(la.y, lb.y), fminf(la.z, lb.z), hgt);
            float4 hi = make_float4(fmaxf(ha.x, hb.x), fmaxf(ha.y, hb.y), fmaxf(ha.z, hb.z), 0.f);
            child[node] = make_int2(a, b);
            ibox[2 * (size_t)node] = lo; ibox[2 * (size_t)node + 1] = hi;
            const uint32_t cnt = ccnt_in[i] + ccnt_in[j];
            icnt[node] = cnt;
            if (a >= 0) parent[a] = node; else parent[(n_leaves - 1) + (~a)] = node;
            if (b >= 0) parent[b] = node; else parent[(n_leaves - 1) + (~b)] = node;
            cid_out[slot] = node; cbox_out[2 * (size_t)slot] = lo; cbox_out[2 * (size_t)slot + 1] = hi; ccnt_out[slot] = cnt;
        } else {
            cid_out[slot] = cid_in[i]; cbox_out[2 * (size_t)slot] = cbox_in[2 * (size_t)i]; cbox_out[2 * (size_t)slot + 1] = cbox_in[2 * (size_t)i + 1];
            ccnt_out[slot] = ccnt_in[i];
        }
    }
}

// The last rounds of PLOC -- a few thousand clusters down to one -- are many (half of all rounds) and tiny: as separate launches they
// are pure launch latency.  ONE CTA finishes them all: nearest neighbours, flags, block-wide prefix sum (in chunks of 1024 with a
// running carry, so any cluster count works) and merge, round after round, with __syncthreads() in between, until one cluster is
// left.  The hierarchy is therefore complete when this kernel ends, however many clusters the queued multi-CTA rounds left over.
__global__ void __launch_bounds__(1024) k_ploc_finish(const PlocState* __restrict__ st, PlocState* __restrict__ st_out,
                                                      int* cid0, float4* box0, uint32_t* cnt0, int* cid1, float4* box1, uint32_t* cnt1,
                                                      int* nn, int n_leaves, int2* child, int* parent, float4* ibox, uint32_t* icnt)
{
    __shared__ unsigned long long warp_sums[32];
    __shared__ unsigned long long s_total;
    int n = st->n_cur, node_base = st->node_base, cur = st->cur, rounds = st->rounds;
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    while (n > 1) {
        const int* cid_in = cur ? cid1 : cid0; const float4* cbox = cur ? box1 : box0; const uint32_t* ccnt_in = cur ? cnt1 : cnt0;
        int* cid_out = cur ? cid0 : cid1; float4* cbox_out = cur ? box0 : box1; uint32_t* ccnt_out = cur ? cnt0 : cnt1;
        for (int i = tid; i < n; i += 1024) {
            const float4 lo = cbox[2 * (size_t)i], hi = cbox[2 * (size_t)i + 1];
            float best = CUDART_INF_F; int bj = -1;
            for (int d = -PLOC_R; d <= PLOC_R; d++) {
                const int j = i + d;
                if (d == 0 || j < 0 || j >= n) continue;
                const float4 l2 = cbox[2 * (size_t)j], h2 = cbox[2 * (size_t)j + 1];
                float4 ul = make_float4(fminf(lo.x, l2.x), fminf(lo.y, l2.y), fminf(lo.z, l2.z), 0.f);
                float4 uh = make_float4(fmaxf(hi.x, h2.x), fmaxf(hi.y, h2.y), fmaxf(hi.z, h2.z), 0.f);
                const float a = half_area(ul, uh);
                if (a < best) { best = a; bj = j; }
            }
            nn[i] = bj;
        }
        __syncthreads();
        unsigned long long carry = 0ull;
        for (int c0 = 0; c0 < n; c0 += 1024) {
            const int i = c0 + tid;
            // (everything this kernel reads was written by it a moment ago: plain coherent loads, nothing through the read-only path)
            unsigned long long f = 0ull;
            if (i < n) {
                const int j = nn[i];
                const bool mutual = j >= 0 && nn[j] == i;
                f = ((unsigned long long)((mutual && i < j) ? 1u : 0u) << 32) | (unsigned long long)((mutual && i > j) ? 0u : 1u);
            }
            unsigned long long v = f;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
            if (lane == 31) warp_sums[warp] = v;
            __syncthreads();
            if (warp == 0) {
                unsigned long long w = warp_sums[lane];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
                warp_sums[lane] = w;
                if (lane == 31) s_total = w;
            }
            __syncthreads();
            const unsigned long long sc = carry + (v - f) + (warp ? warp_sums[warp - 1] : 0ull);
            if (i < n && (f & 1ull)) {
                const int slot = (int)(sc & 0xffffffffull);
                if (f >> 32) {
                    const int j = nn[i];
                    const int node = node_base + (int)(sc >> 32);
                    const int a = cid_in[i], b = cid_in[j];
                    float4 la = cbox[2 * (size_t)i], ha = cbox[2 * (size_t)i + 1], lb = cbox[2 * (size_t)j], hb = cbox[2 * (size_t)j + 1];
                    float hgt = fmaxf(a < 0 ? 0.f : la.w, b < 0 ? 0.f : lb.w) + 1.0f;
                    float4 lo = make_float4(fminf(la.x, lb.x), fminf(la.y, lb.y), fminf(la.z, lb.z), hgt);
                    float4 hi = make_float4(fmaxf(ha.x, hb.x), fmaxf(ha.y, hb.y), fmaxf(ha.z, hb.z), 0.f);
                    child[node] = make_int2(a, b);
                    ibox[2 * (size_t)node] = lo; ibox[2 * (size_t)node + 1] = hi;
                    const uint32_t cnt = ccnt_in[i] + ccnt_in[j];
                    icnt[node] = cnt;
                    if (a >= 0) parent[a] = node; else parent[(n_leaves - 1) + (~a)] = node;
                    if (b >= 0) parent[b] = node; else parent[(n_leaves - 1) + (~b)] = node;
                    cid_out[slot] = node; cbox_out[2 * (size_t)slot] = lo; cbox_out[2 * (size_t)slot + 1] = hi; ccnt_out[slot] = cnt;
                } else {
                    cid_out[slot] = cid_in[i]; cbox_out[2 * (size_t)slot] = cbox[2 * (size_t)i]; cbox_out[2 * (size_t)slot + 1] = cbox[2 * (size_t)i + 1];
                    ccnt_out[slot] = ccnt_in[i];
                }
            }
            carry += s_total;
            __syncthreads();
        }
        node_base += (int)(carry >> 32); n = (int)(carry & 0xffffffffull); cur ^= 1; rounds++;
        __syncthreads();
    }
    if (tid == 0) { PlocState nx; nx.n_cur = n; nx.node_base = node_base; nx.cur = cur; nx.rounds = rounds; *st_out = nx; }
}

__global__ void k_set_int(int* p, int v) { *p = v; }

// what the host wants to know about a finished build, gathered into one record (copied to pinned memory, read after ONE synchronisation)
__global__ void k_build_result(const PlocState* __restrict__ st, const uint32_t* __restrict__ rank, const uint32_t* __restrict__ keep, uint32_t n_internal,
                               const vkp_bvh_node* __restrict__ nodes, const int* __restrict__ height, BvhBuildResult* __restrict__ out)
{
    BvhBuildResult r;
    r.n_nodes = 1u + 2u * (rank[n_internal - 1] + keep[n_internal - 1]);
    r.root_ref = nodes[0].ref;
    r.height = *height;
    r.ploc_left = st ? st->n_cur : 1;
    r.ploc_rounds = st ? st->rounds : 0;
    r.valid = 1;
    *out = r;
}

// depth-first position of a node = sum over its ancestors (where it hangs on the right) of the left sibling's leaf count
__device__ __forceinline__ uint32_t ploc_first(int self_is_leaf, int self, int n_leaves, const int2* __restrict__ child, const int* __restrict__ parent,
                                               const uint32_t* __restrict__ icnt, int root)
{
    uint32_t first = 0;
    int cur = self_is_leaf ? ~self : self;                     // encoded like child entries
    int p = self_is_leaf ? parent[(n_leaves - 1) + self] : (self == root ? -1 : parent[self]);
    while (p >= 0) {
        int2 c = child[p];
        if (c.y == cur) first += c.x < 0 ? 1u : icnt[c.x];
        cur = p;
        p = p == root ? -1 : parent[p];
    }
    return first;
}

__global__ void k_ploc_ranges(int n_leaves, const int2* __restrict__ child, const int* __restrict__ parent, const uint32_t* __restrict__ icnt, int root,
                              int2* __restrict__ range, uint32_t* __restrict__ newpos, const PlocState* __restrict__ gate)
{
    if (gate_closed(gate)) return;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_leaves - 1) {
        uint32_t f = ploc_first(0, t, n_leaves, child, parent, icnt, root);
        range[t] = make_int2((int)f, (int)(f + icnt[t] - 1u));
    }
    if (t < n_leaves) newpos[t] = ploc_first(1, t, n_leaves, child, parent, icnt, root);
}

// leaves are re-ordered depth-first so that every sub-tree owns a contiguous range of `items` (needed to collapse
// sub-trees into multi-triangle leaves); child / parent entries that name leaves follow the permutation
__global__ void k_ploc_permute(int n_leaves, const uint32_t* __restrict__ newpos, const uint32_t* __restrict__ items_in, uint32_t* __restrict__ items_out,
                               int2* __restrict__ child, const int* __restrict__ parent_in, int* __restrict__ parent_out, const PlocState* __restrict__ gate)
{
    if (gate_closed(gate)) return;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_leaves) {
        items_out[newpos[t]] = items_in[t];
        parent_out[(n_leaves - 1) + newpos[t]] = parent_in[(n_leaves - 1) + t];
    }
    if (t < n_leaves - 1) {
        int2 c = child[t];
        if (c.x < 0) c.x = ~(int)newpos[~c.x];
        if (c.y < 0) c.y = ~(int)newpos[~c.y];
        child[t] = c;
        parent_out[t] = parent_in[t];
    }
}

inline unsigned blocks_for(uint32_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

cudaError_t fit_and_emit(cudaStream_t s, BuildScratch& sc, Bvh& B, int* launches, bool read_height = true)
{
    const uint32_t n = B.n_prims;
    cudaError_t e;
    if (n <= 1) {
        k_emit_single<<<1, 1, 0, s>>>(sc.boxes.as<float>(), (int)n, B.nodes.as<vkp_bvh_node>());
        if (launches) *launches += 1;
        B.n_nodes = 1; B.height = 0;
        return cudaGetLastError();
    }
    // (clears by kernel, not cudaMemsetAsync: keeps the per-frame refit off the copy engines, see launch_stage_words)
    if ((e = launch_stage_words(s, B.flags.p, nullptr, sizeof(int) * (n - 1))) != cudaSuccess) return e;
    int* height_dev = sc.misc.as<int>();
    if ((e = launch_stage_words(s, height_dev, nullptr, sizeof(int))) != cudaSuccess) return e;
    k_fit<<<blocks_for(n, 256), 256, 0, s>>>(sc.boxes.as<float>(), B.items.as<uint32_t>(), (int)n, B.child.as<int2>(), B.parent.as<int>(),
                                             B.ibox.as<float4>(), B.flags.as<int>(), height_dev, B.root, static_cast<const PlocState*>(B.gate));
    k_emit_nodes<<<blocks_for(n - 1, 256), 256, 0, s>>>(sc.boxes.as<float>(), B.items.as<uint32_t>(), (int)n, B.child.as<int2>(), B.range.as<int2>(),
                                                        B.ibox.as<float4>(), B.keep.as<uint32_t>(), B.rank.as<uint32_t>(), B.nodes.as<vkp_bvh_node>(), B.root, B.leaf_max == 1 ? 1 : 0, static_cast<const PlocState*>(B.gate));
    if (launches) *launches += 2;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    // a refit keeps the topology, hence the height: no D2H (a copy into pageable memory would make the per-frame path wait for the device)
    if (!read_height) return cudaSuccess;
    return cudaMemcpyAsync(&B.height, height_dev, sizeof(int), cudaMemcpyDeviceToHost, s);
}

} // namespace

cudaError_t bvh_triangle_boxes(cudaStream_t s, const vkp_vertex* verts, const uint32_t* indices, uint32_t n_tris, float* boxes)
{
    if (n_tris == 0) return cudaSuccess;
    k_triangle_boxes<<<blocks_for(n_tris, 256), 256, 0, s>>>(verts, indices, n_tris, boxes);
    return cudaGetLastError();
}

cudaError_t bvh_instance_boxes(cudaStream_t s, const DInstance* inst, uint32_t n, float* boxes)
{
    if (n == 0) return cudaSuccess;
    k_instance_boxes<<<blocks_for(n, 128), 128, 0, s>>>(inst, n, boxes);
    return cudaGetLastError();
}

// multi-CTA rounds the host queues: enough to bring the cluster count to ~2000 at a pessimistic 20 % reduction per round (triangle
// meshes merge 25-40 %); whatever is left -- however much -- is finished by k_ploc_finish
static int ploc_round_budget(uint32_t n)
{
    if (const char* e = getenv("VKP_PLOC_ROUNDS")) { int v = atoi(e); if (v >= 0) return v; }      // test hook: leave (nearly) everything to the finish kernel
    int r = 0;
    for (double x = (double)n; x > 2048.0; x *= 0.8) r++;
    return r;
}

static void ploc_queue_rounds(cudaStream_t s, BuildScratch& sc, Bvh& B, int first_round, int rounds, int* launches)
{
    const int n = (int)B.n_prims;
    PlocState* st = sc.ploc_state.as<PlocState>();
    unsigned long long* tile_sums = sc.ploc_tiles.as<unsigned long long>();
    unsigned long long* tile_offs = tile_sums + ((size_t)n + PLOC_SC_TILE - 1) / PLOC_SC_TILE + 1;
    double est = (double)n;
    for (int k = 0; k < first_round; k++) est *= 0.85;
    for (int r = first_round; r < first_round + rounds; r++) {
        // grids follow a generous estimate of the cluster count; the kernels loop (grid-stride) if the round is larger than that
        const uint32_t m = (uint32_t)(est < 4096.0 ? 4096.0 : est);
        const PlocState* sin = st + (r & 1); PlocState* sout = st + ((r + 1) & 1);
        k_ploc_nn_dev<<<blocks_for(m, PLOC_BS), PLOC_BS, 0, s>>>(sin, sc.ploc_box[0].as<float4>(), sc.ploc_box[1].as<float4>(), sc.ploc_nn.as<int>());
        k_ploc_scan_tiles<<<blocks_for(m, PLOC_SC_TILE), PLOC_SC_THREADS, 0, s>>>(sin, sc.ploc_nn.as<int>(), sc.ploc_scan.as<unsigned long long>(), tile_sums);
        k_ploc_scan_state<<<1, 1024, 0, s>>>(sin, sout, tile_sums, tile_offs);
        k_ploc_merge_dev<<<blocks_for(m, 256), 256, 0, s>>>(sin, sc.ploc_nn.as<int>(), sc.ploc_scan.as<unsigned long long>(), tile_offs,
                                                          sc.ploc_cid[0].as<int>(), sc.ploc_box[0].as<float4>(), sc.ploc_cnt[0].as<uint32_t>(),
                                                          sc.ploc_cid[1].as<int>(), sc.ploc_box[1].as<float4>(), sc.ploc_cnt[1].as<uint32_t>(), n,
                                                          B.child.as<int2>(), B.parent.as<int>(), B.ibox.as<float4>(), sc.ploc_icnt.as<uint32_t>());
        est *= 0.85;
        if (launches) *launches += 4;
    }
}

// wait == false: the rounds of ploc_round_budget() are queued and nothing is waited for; whether they sufficed is in the build
// result (BvhBuildResult::ploc_left == 1).  wait == true: batches of rounds until one cluster is left, one read-back per batch.
static cudaError_t ploc_hierarchy(cudaStream_t s, BuildScratch& sc, Bvh& B, int* launches, int* host_syncs, bool wait, int* rounds_queued)
{
    const int n = (int)B.n_prims;
    cudaError_t e;
    k_ploc_init<<<blocks_for(n, 256), 256, 0, s>>>(sc.boxes.as<float>(), B.items.as<uint32_t>(), n, sc.ploc_cid[0].as<int>(), sc.ploc_box[0].as<float4>(), sc.ploc_cnt[0].as<uint32_t>());
    k_ploc_state_init<<<1, 1, 0, s>>>(sc.ploc_state.as<PlocState>(), n);
    if (launches) *launches += 2;
    int done = ploc_round_budget((uint32_t)n);
    ploc_queue_rounds(s, sc, B, 0, done, launches);
    {
        PlocState* st = sc.ploc_state.as<PlocState>();
        k_ploc_finish<<<1, 1024, 0, s>>>(st + (done & 1), st + ((done + 1) & 1),
                                         sc.ploc_cid[0].as<int>(), sc.ploc_box[0].as<float4>(), sc.ploc_cnt[0].as<uint32_t>(),
                                         sc.ploc_cid[1].as<int>(), sc.ploc_box[1].as<float4>(), sc.ploc_cnt[1].as<uint32_t>(),
                                         sc.ploc_nn.as<int>(), n, B.child.as<int2>(), B.parent.as<int>(), B.ibox.as<float4>(), sc.ploc_icnt.as<uint32_t>());
        done += 1;                                      // the state index moved on by one
        if (launches) *launches += 1;
    }
    while (wait) {
        PlocState st;
        if ((e = cudaMemcpyAsync(&st, sc.ploc_state.as<PlocState>() + (done & 1), sizeof st, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
        if (host_syncs) *host_syncs += 1;
        if (st.n_cur <= 1) break;
        if (done > 2 * n + 64) return cudaErrorUnknown;         // cannot happen: the globally closest pair is always mutual, every round merges
        ploc_queue_rounds(s, sc, B, done, 8, launches);
        done += 8;
    }
    if (rounds_queued) *rounds_queued = done;
    B.gate = sc.ploc_state.as<PlocState>() + (done & 1);        // the state the last queued round left
    // n - 1 nodes are created in the order of their indices and the last merge makes the root
    const int root = n - 2;
    B.root = root;
    // depth-first leaf order, contiguous ranges, permuted leaf references
    k_ploc_ranges<<<blocks_for(n, 256), 256, 0, s>>>(n, B.child.as<int2>(), B.parent.as<int>(), sc.ploc_icnt.as<uint32_t>(), root, B.range.as<int2>(), sc.ploc_nn.as<uint32_t>(), static_cast<const PlocState*>(B.gate));
    k_ploc_permute<<<blocks_for(n, 256), 256, 0, s>>>(n, sc.ploc_nn.as<uint32_t>(), B.items.as<uint32_t>(), sc.vals_alt.as<uint32_t>(), B.child.as<int2>(), B.parent.as<int>(),
                                                       sc.ploc_parent.as<int>(), static_cast<const PlocState*>(B.gate));
    if ((e = cudaMemcpyAsync(B.items.p, sc.vals_alt.p, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToDevice, s)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(B.parent.p, sc.ploc_parent.p, sizeof(int) * (2 * (size_t)n - 1), cudaMemcpyDeviceToDevice, s)) != cudaSuccess) return e;
    k_set_int<<<1, 1, 0, s>>>(B.parent.as<int>() + root, -1);      // the root ends every bottom-up walk (k_fit, refit)
    if (launches) *launches += 3;
    return cudaGetLastError();
}

// every scratch array a build of n primitives needs, allocated (or grown) up front: nothing is allocated or freed -- cudaFree
// synchronises the device -- between the builds of a batch
cudaError_t bvh_reserve(BuildScratch& sc, uint32_t n, int builder, bool temps)
{
    cudaError_t e;
    const size_t m = n ? n : 1, ni = n > 1 ? n - 1 : 1;
    if ((e = sc.boxes.ensure(sizeof(float) * 6 * m)) != cudaSuccess) return e;
    if ((e = sc.misc.ensure(256)) != cudaSuccess) return e;
    if ((e = sc.keys.ensure(sizeof(uint32_t) * m)) != cudaSuccess) return e;
    if ((e = sc.keys_alt.ensure(sizeof(uint32_t) * m)) != cudaSuccess) return e;
    if ((e = sc.vals.ensure(sizeof(uint32_t) * m)) != cudaSuccess) return e;
    if ((e = sc.bounds.ensure(64)) != cudaSuccess) return e;
    {
        size_t need = vkpsort::sort_scratch_bytes(n), need2 = vkpsort::scan_scratch_bytes(n, sizeof(unsigned long long));
        if ((e = sc.cub_tmp.ensure((need > need2 ? need : need2) + 256)) != cudaSuccess) return e;
    }
    if (builder == 1) {
        for (int k = 0; k < 2; k++) {
            if ((e = sc.ploc_cid[k].ensure(sizeof(int) * m)) != cudaSuccess) return e;
            if ((e = sc.ploc_box[k].ensure(sizeof(float4) * 2 * m)) != cudaSuccess) return e;
            if ((e = sc.ploc_cnt[k].ensure(sizeof(uint32_t) * m)) != cudaSuccess) return e;
        }
        if ((e = sc.ploc_nn.ensure(sizeof(int) * m)) != cudaSuccess) return e;
        if ((e = sc.ploc_scan.ensure(sizeof(unsigned long long) * m)) != cudaSuccess) return e;
        if ((e = sc.ploc_icnt.ensure(sizeof(uint32_t) * m)) != cudaSuccess) return e;
        if ((e = sc.vals_alt.ensure(sizeof(uint32_t) * m)) != cudaSuccess) return e;
        if ((e = sc.ploc_parent.ensure(sizeof(int) * 2 * m)) != cudaSuccess) return e;
        if ((e = sc.ploc_state.ensure(2 * sizeof(PlocState))) != cudaSuccess) return e;
        if ((e = sc.ploc_tiles.ensure(sizeof(unsigned long long) * 2 * ((m + PLOC_SC_TILE - 1) / PLOC_SC_TILE + 2))) != cudaSuccess) return e;
    }
    if (temps) {
        if ((e = sc.t_child.ensure(sizeof(int2) * ni)) != cudaSuccess) return e;
        if ((e = sc.t_range.ensure(sizeof(int2) * ni)) != cudaSuccess) return e;
        if ((e = sc.t_parent.ensure(sizeof(int) * (ni + m))) != cudaSuccess) return e;
        if ((e = sc.t_ibox.ensure(sizeof(float4) * 2 * ni)) != cudaSuccess) return e;
        if ((e = sc.t_flags.ensure(sizeof(int) * ni)) != cudaSuccess) return e;
        if ((e = sc.t_rank.ensure(sizeof(uint32_t) * ni)) != cudaSuccess) return e;
        if ((e = sc.t_keep.ensure(sizeof(uint32_t) * ni)) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// async_result (pinned host memory, written by the last kernel of the build) != NULL: nothing is waited for; the caller synchronises
// the stream once -- for as many builds as it likes -- and completes each Bvh with bvh_build_finish().  NULL: waits and finishes here.
// temps_in_scratch: the topology arrays only the build itself (and a refit) needs are views of the scratch, valid until its next build.
cudaError_t bvh_build(cudaStream_t s, BuildScratch& sc, uint32_t n, int leaf_max, Bvh& B, int* launches, int builder, int* host_syncs,
                      BvhBuildResult* async_result, bool temps_in_scratch)
{
    cudaError_t e;
    B.n_prims = n; B.leaf_max = leaf_max;
    const uint32_t ni = n > 1 ? n - 1 : 0;
    if ((e = bvh_reserve(sc, n, n > 1 ? builder : 0, temps_in_scratch && n > 1)) != cudaSuccess) return e;
    if ((e = B.nodes.ensure(sizeof(vkp_bvh_node) * (2 * (size_t)ni + 1))) != cudaSuccess) return e;
    if ((e = B.items.ensure(sizeof(uint32_t) * (size_t)(n ? n : 1))) != cudaSuccess) return e;
    if (n <= 1) {
        if (n == 1 && (e = cudaMemsetAsync(B.items.p, 0, sizeof(uint32_t), s)) != cudaSuccess) return e;
        B.root_ref = 0x80000000u;     // a single leaf: first item 0, one item (an empty BLAS holds one degenerate triangle)
        if (async_result) { async_result->valid = 0; }
        return fit_and_emit(s, sc, B, launches);
    }
    if (temps_in_scratch) {
        B.child.view(sc.t_child.p, sizeof(int2) * ni); B.range.view(sc.t_range.p, sizeof(int2) * ni); B.parent.view(sc.t_parent.p, sizeof(int) * (ni + (size_t)n));
        B.ibox.view(sc.t_ibox.p, sizeof(float4) * 2 * ni); B.flags.view(sc.t_flags.p, sizeof(int) * ni); B.rank.view(sc.t_rank.p, sizeof(uint32_t) * ni);
        B.keep.view(sc.t_keep.p, sizeof(uint32_t) * ni);
    } else {
        if ((e = B.child.ensure(sizeof(int2) * ni)) != cudaSuccess) return e;
        if ((e = B.range.ensure(sizeof(int2) * ni)) != cudaSuccess) return e;
        if ((e = B.parent.ensure(sizeof(int) * (ni + (size_t)n))) != cudaSuccess) return e;
        if ((e = B.ibox.ensure(sizeof(float4) * 2 * ni)) != cudaSuccess) return e;
        if ((e = B.flags.ensure(sizeof(int) * ni)) != cudaSuccess) return e;
        if ((e = B.rank.ensure(sizeof(uint32_t) * ni)) != cudaSuccess) return e;
        if ((e = B.keep.ensure(sizeof(uint32_t) * ni)) != cudaSuccess) return e;
    }

    k_init_bounds<<<1, 32, 0, s>>>(sc.bounds.as<int>());
    unsigned rb = blocks_for(n, 256); if (rb > 592) rb = 592;
    k_centroid_bounds<<<rb, 256, 0, s>>>(sc.boxes.as<float>(), n, sc.bounds.as<int>());
    k_morton<<<blocks_for(n, 256), 256, 0, s>>>(sc.boxes.as<float>(), n, sc.bounds.as<int>(), sc.keys.as<uint32_t>(), sc.vals.as<uint32_t>());
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((e = vkpsort::sort_pairs(s, sc.keys.as<uint32_t>(), sc.vals.as<uint32_t>(), sc.keys_alt.as<uint32_t>(), B.items.as<uint32_t>(), n, 30, sc.cub_tmp.p,
                                 launches)) != cudaSuccess) return e;
    B.root = 0; B.gate = nullptr;
    int rounds = 0;
    if (builder == 1) {
        if ((e = ploc_hierarchy(s, sc, B, launches, host_syncs, async_result == nullptr, &rounds)) != cudaSuccess) return e;
    } else {
        k_hierarchy<<<blocks_for(ni, 256), 256, 0, s>>>(sc.keys_alt.as<uint32_t>(), (int)n, B.child.as<int2>(), B.parent.as<int>(), B.range.as<int2>());
    }
    k_mark_keep<<<blocks_for(ni, 256), 256, 0, s>>>(B.range.as<int2>(), (int)ni, leaf_max, B.keep.as<uint32_t>());
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((e = vkpsort::exclusive_scan<uint32_t>(s, B.keep.as<uint32_t>(), B.rank.as<uint32_t>(), ni, sc.cub_tmp.p, launches)) != cudaSuccess) return e;
    if (launches) *launches += 5;
    // PLOC already knows every box, but the shared fit pass is cheap and keeps one code path (also used by refit)
    if ((e = fit_and_emit(s, sc, B, launches, false)) != cudaSuccess) return e;
    // number of emitted nodes (1 + 2 * kept internals), root reference, height, PLOC state: one record, one wait
    BvhBuildResult* res = async_result;
    if (!res) {
        if (!sc.result_host && (e = cudaMallocHost(&sc.result_host, sizeof(BvhBuildResult))) != cudaSuccess) return e;
        res = static_cast<BvhBuildResult*>(sc.result_host);
    }
    res->valid = 0;
    k_build_result<<<1, 1, 0, s>>>(builder == 1 ? sc.ploc_state.as<PlocState>() + (rounds & 1) : nullptr, B.rank.as<uint32_t>(), B.keep.as<uint32_t>(), ni,
                                   B.nodes.as<vkp_bvh_node>(), sc.misc.as<int>(), res);
    if (launches) *launches += 1;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (async_result) return cudaSuccess;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    if (host_syncs) *host_syncs += 1;
    return bvh_build_finish(B, *res) ? cudaSuccess : cudaErrorUnknown;
}

// after the stream has passed the build: false = the queued PLOC rounds did not suffice (build again with async_result == NULL)
bool bvh_build_finish(Bvh& B, const BvhBuildResult& r)
{
    if (B.n_prims <= 1) return true;                  // fit_and_emit filled it in
    if (!r.valid || r.ploc_left != 1) { B.gate = nullptr; return false; }
    B.n_nodes = r.n_nodes; B.root_ref = r.root_ref; B.height = r.height; B.ploc_rounds = r.ploc_rounds;
    B.gate = nullptr;             // the scratch state it named belongs to the next build
    return true;
}

cudaError_t bvh_refit(cudaStream_t s, BuildScratch& sc, Bvh& B, int* launches)
{
    return fit_and_emit(s, sc, B, launches, false);
}

cudaError_t bvh_emit_triangles(cudaStream_t s, const vkp_vertex* verts, const uint32_t* indices, Bvh& B)
{
    cudaError_t e;
    if ((e = B.tris.ensure(sizeof(float4) * 3 * (size_t)(B.n_prims ? B.n_prims : 1))) != cudaSuccess) return e;
    if (B.n_prims == 0) return cudaMemsetAsync(B.tris.p, 0, sizeof(float4) * 3, s);   // one degenerate triangle: det == 0, never hit
    k_emit_triangles<<<blocks_for(B.n_prims, 256), 256, 0, s>>>(verts, indices, B.items.as<uint32_t>(), B.n_prims, B.tris.as<float4>(), static_cast<const PlocState*>(B.gate));
    return cudaGetLastError();
}

// test hook (vkp_selftest_sort): sort host keys with the build's own radix sort and scan them with its own prefix sum
cudaError_t bvh_selftest_sort(cudaStream_t s, const uint32_t* keys_host, uint32_t n, int end_bit, uint32_t* keys_sorted_host, uint32_t* vals_sorted_host,
                              unsigned long long* scan_host)
{
    if (n == 0) return cudaSuccess;
    DevBuf k0, v0, k1, v1, tmp, in64, out64;
    cudaError_t e;
    std::vector<uint32_t> iota(n);
    for (uint32_t i = 0; i < n; i++) iota[i] = i;
    std::vector<unsigned long long> wide(n);
    for (uint32_t i = 0; i < n; i++) wide[i] = ((unsigned long long)(keys_host[i] & 1u) << 32) | (unsigned long long)(keys_host[i] & 3u);
    size_t need = vkpsort::sort_scratch_bytes(n), need2 = vkpsort::scan_scratch_bytes(n, 8);
    if ((e = k0.ensure(4 * (size_t)n)) != cudaSuccess || (e = v0.ensure(4 * (size_t)n)) != cudaSuccess || (e = k1.ensure(4 * (size_t)n)) != cudaSuccess ||
        (e = v1.ensure(4 * (size_t)n)) != cudaSuccess || (e = tmp.ensure(need > need2 ? need : need2)) != cudaSuccess || (e = in64.ensure(8 * (size_t)n)) != cudaSuccess ||
        (e = out64.ensure(8 * (size_t)n)) != cudaSuccess) return e;
    cudaMemcpyAsync(k0.p, keys_host, 4 * (size_t)n, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(v0.p, iota.data(), 4 * (size_t)n, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(in64.p, wide.data(), 8 * (size_t)n, cudaMemcpyHostToDevice, s);
    if ((e = vkpsort::sort_pairs(s, k0.as<uint32_t>(), v0.as<uint32_t>(), k1.as<uint32_t>(), v1.as<uint32_t>(), n, end_bit, tmp.p, nullptr)) != cudaSuccess) return e;
    cudaMemcpyAsync(keys_sorted_host, k1.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(vals_sorted_host, v1.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, s);
    if ((e = vkpsort::exclusive_scan<unsigned long long>(s, in64.as<unsigned long long>(), out64.as<unsigned long long>(), n, tmp.p, nullptr)) != cudaSuccess) return e;
    cudaMemcpyAsync(scan_host, out64.p, 8 * (size_t)n, cudaMemcpyDeviceToHost, s);
    e = cudaStreamSynchronize(s);
    k0.release(); v0.release(); k1.release(); v1.release(); tmp.release(); in64.release(); out64.release();
    return e;
}
