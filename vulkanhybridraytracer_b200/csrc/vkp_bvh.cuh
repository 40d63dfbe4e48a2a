// vkp_bvh.cuh -- two-level BVH traversal (what the Vulkan driver + RT cores do behind traceRayEXT:
// shadowRaygen.rgen:78, hybridRaygen.rgen:158, hybridHit.rchit:123, raygen.rgen:49, closesthit.rchit:111).
//
// Layout (DESIGN.md "BVH layout"): 32-byte nodes {bmin, ref, bmax, aux}.  Node 0 is the root.  An internal
// node's `ref` is the index of its left child; the right child is ref+1, so one 64-byte aligned fetch brings
// both children.  A leaf's ref has bit 31 set: bits[30:3] = first item, bits[2:0] = count-1.  BLAS items are
// triangles stored as 3 x float4 in leaf order (v0 | primitive id, v1, v2); TLAS items are instance ids.
//
// Intersection contract: Moeller-Trumbore in object space, no culling, tmin < t < tmax, u >= 0, v >= 0,
// u + v <= 1; closest hit = min t, then lower instance id, then lower primitive id.  The arithmetic order
// of tri_test() is the contract -- do not re-associate it.
#pragma once
#include "vkp_device.cuh"

#define VKP_LEAF_BIT 0x80000000u
#define VKP_STACK_SENTINEL 0xffffffffu
#define VKP_STACK_RETURN 0xfffffffeu
// VKP_WIDE: 0 = binary nodes (default), 1 = 4-wide nodes.  Measured on B200: the wide layout halves the node
// fetches per ray but costs more slab tests and registers; it is neutral on the latency-bound 1080p Hybrid frame and
// 10-26 % slower on the throughput-bound workloads (cfg3 / cfg4 / cfg5), so the binary layout stays the default.
#ifndef VKP_WIDE
#define VKP_WIDE 0
#endif
#if VKP_WIDE
#define VKP_TRAV_STACK 128
#else
#define VKP_TRAV_STACK 64
#endif
#define VKP_W_EMPTY_REF 0xfffffffdu

struct HitRec {
    float t, u, v;
    int inst, prim;
};

// what a thread's rays actually requested from memory: node fetches (64 B each in the binary layout), triangle tests (48 B)
// and instance entries (80 B); summed per frame into the device counters for the "requested bytes" figure of bench.py
struct TravCount { unsigned int nodes, tris, entries; };

VKP_D float3 safe_inv_dir(float3 d)
{
    // a zero component would give inf and 0*inf = NaN in the slab test; a huge finite reciprocal keeps the
    // test conservative (inside the slab -> unbounded interval, outside -> empty)
    const float big = 1e30f;
    float3 r;
    r.x = (fabsf(d.x) > 1e-30f) ? 1.0f / d.x : copysignf(big, d.x);
    r.y = (fabsf(d.y) > 1e-30f) ? 1.0f / d.y : copysignf(big, d.y);
    r.z = (fabsf(d.z) > 1e-30f) ? 1.0f / d.z : copysignf(big, d.z);
    return r;
}

// conservative slab test in (b - o) * inv form (relative error only) with slack on both ends.  PAR == true is the
// rare path for rays with a (near-)zero direction component: such a ray is inside a slab for all t or for none,
// including when it lies exactly in a face plane of the box (0 * inf in exact arithmetic).
template <bool PAR>
VKP_D bool box_test(const float4 lo, const float4 hi, float3 o, float3 d, float3 inv, float tmin, float tmax, float& tnear)
{
    float x0 = (lo.x - o.x) * inv.x, x1 = (hi.x - o.x) * inv.x;
    float y0 = (lo.y - o.y) * inv.y, y1 = (hi.y - o.y) * inv.y;
    float z0 = (lo.z - o.z) * inv.z, z1 = (hi.z - o.z) * inv.z;
    if (PAR) {
        // A ray parallel to a slab is inside it for every t or for none.  "None" has to leave here: no (t0, t1) pair survives
        // the min / max below as an empty interval.  (It used to be encoded as (+inf, -inf), which the min / max turned back
        // into "unbounded": such rays -- one per few frames, when a light sample cancels a coordinate exactly -- were never
        // culled on that axis, walked thousands of nodes and set the tail of the whole kernel.)  The slack keeps the test
        // conservative against the rounding of the triangle test for rays that run along a face of the box.
        const float inf = __int_as_float(0x7f800000);
        tnear = inf;
        if (fabsf(d.x) <= 1e-30f) { const float s = 1e-5f * fmaxf(fabsf(o.x), fmaxf(fabsf(lo.x), fabsf(hi.x))); if (lo.x - o.x > s || o.x - hi.x > s) return false; x0 = -inf; x1 = inf; }
        if (fabsf(d.y) <= 1e-30f) { const float s = 1e-5f * fmaxf(fabsf(o.y), fmaxf(fabsf(lo.y), fabsf(hi.y))); if (lo.y - o.y > s || o.y - hi.y > s) return false; y0 = -inf; y1 = inf; }
        if (fabsf(d.z) <= 1e-30f) { const float s = 1e-5f * fmaxf(fabsf(o.z), fmaxf(fabsf(lo.z), fabsf(hi.z))); if (lo.z - o.z > s || o.z - hi.z > s) return false; z0 = -inf; z1 = inf; }
    }
    float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tmin));
    float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));
    tnear = tn;
    return tn * 0.999999f <= tf * 1.000001f;
}
// The common-case slab test: t = b * (1/d) - o * (1/d) as one FMA per plane, the near / far plane of each axis chosen by the sign
// of the direction (no min / max per axis), and the rounding of the precomputed o * (1/d) -- the only absolute error of this
// form -- folded into two copies of it: oin = o/d + e for the near planes, oif = o/d - e for the far ones, e = |o/d| * 2^-22.
// 18 instructions per box instead of 26.  Not used when a direction component is (near) zero: that is box_test<true>.
#ifndef VKP_SLAB_FMA
#define VKP_SLAB_FMA 0      /* measured 2-3 % slower per frame than the (b - o) * inv form: fewer instructions, more live registers */
#endif
struct SlabRay { float3 inv, oin, oif; bool nx, ny, nz; };
VKP_D SlabRay slab_setup(float3 o, float3 inv)
{
    SlabRay R;
    R.inv = inv;
    const float ox = o.x * inv.x, oy = o.y * inv.y, oz = o.z * inv.z;
    const float ex = fabsf(ox) * 2.4e-7f, ey = fabsf(oy) * 2.4e-7f, ez = fabsf(oz) * 2.4e-7f;
    R.oin = f3(ox + ex, oy + ey, oz + ez);
    R.oif = f3(ox - ex, oy - ey, oz - ez);
    R.nx = inv.x < 0.0f; R.ny = inv.y < 0.0f; R.nz = inv.z < 0.0f;
    return R;
}
VKP_D bool box_test_fast(const float4 lo, const float4 hi, const SlabRay& R, float tmin, float tmax, float& tnear)
{
    const float ax = __fmaf_rn(R.nx ? hi.x : lo.x, R.inv.x, -R.oin.x), bx = __fmaf_rn(R.nx ? lo.x : hi.x, R.inv.x, -R.oif.x);
    const float ay = __fmaf_rn(R.ny ? hi.y : lo.y, R.inv.y, -R.oin.y), by = __fmaf_rn(R.ny ? lo.y : hi.y, R.inv.y, -R.oif.y);
    const float az = __fmaf_rn(R.nz ? hi.z : lo.z, R.inv.z, -R.oin.z), bz = __fmaf_rn(R.nz ? lo.z : hi.z, R.inv.z, -R.oif.z);
    const float tn = fmaxf(fmaxf(ax, ay), fmaxf(az, tmin));
    const float tf = fminf(fminf(bx, by), fminf(bz, tmax));
    tnear = tn;
    return tn <= tf * 1.000001f;
}

VKP_D bool has_parallel_axis(float3 d) { return fabsf(d.x) <= 1e-30f || fabsf(d.y) <= 1e-30f || fabsf(d.z) <= 1e-30f; }
VKP_D bool box_test(const float4 lo, const float4 hi, float3 o, float3 d, float3 inv, bool par, float tmin, float tmax, float& tnear)
{
    return par ? box_test<true>(lo, hi, o, d, inv, tmin, tmax, tnear) : box_test<false>(lo, hi, o, d, inv, tmin, tmax, tnear);
}

// the intersection contract
VKP_D bool tri_test(float3 o, float3 d, float3 v0, float3 v1, float3 v2, float tmin, float tmax, float& t, float& u, float& v)
{
    float3 e1 = v1 - v0, e2 = v2 - v0;
    float3 p = cross3(d, e2);
    float det = dot3(e1, p);
    if (det == 0.0f) return false;
    float inv = 1.0f / det;
    float3 s = o - v0;
    u = dot3(s, p) * inv;
    if (!(u >= 0.0f) || u > 1.0f) return false;
    float3 q = cross3(s, e1);
    v = dot3(d, q) * inv;
    if (!(v >= 0.0f) || u + v > 1.0f) return false;
    t = dot3(e2, q) * inv;
    return t > tmin && t < tmax;
}

VKP_D bool hit_better(float t, int inst, int prim, const HitRec& h)
{
    if (h.inst < 0) return true;
    if (t < h.t) return true;
    if (t > h.t) return false;
    if (inst != h.inst) return inst < h.inst;
    return prim < h.prim;
}

// ANY == true: gl_RayFlagsTerminateOnFirstHitEXT | SkipClosestHitShaderEXT (shadow rays): returns on the first hit
#ifdef VKP_NOINLINE_TRACE
#define VKP_TRACE_FN __device__ __noinline__
#else
#define VKP_TRACE_FN __device__ __forceinline__
#endif
// LIMITED == true (one-sample shadow rays, k_shadow): after `max_steps` child-pair fetches the thread reserves a slot in the
// spill queue and gives the ray up -- best.prim = VKP_TRACE_SPILLED, best.inst = slot -- for the warp-cooperative tail kernel
// (warp_any_hit below); when the queue is full it simply carries on.  A kernel ends with its slowest thread, and the slowest
// shadow rays walk 50-100x the average number of nodes.
#define VKP_TRACE_SPILLED (-2)
template <bool ANY, bool STATS = false, bool LIMITED = false>
VKP_TRACE_FN bool trace_ray(const DScene& S, float3 o, float3 d, float tmin, float tmax, HitRec& best, uint32_t* stats = nullptr, TravCount* tcount = nullptr,
                            int max_steps = 0, unsigned int* spill_count = nullptr, unsigned int spill_cap = 0u)
{
    uint32_t n_steps = 0, n_tris = 0;      // STATS: node pairs fetched, triangles tested
    int n_limited = 0;
    best.t = -1.0f; best.u = 0.0f; best.v = 0.0f; best.inst = -1; best.prim = -1;
    if (S.n_instances == 0) return false;     // empty TLAS: its root box is inverted, which a slab test cannot reject
    float limit = tmax;                       // current far bound (closest so far)

    uint32_t stack[VKP_TRAV_STACK];
    int sp = 0;
    stack[sp++] = VKP_STACK_SENTINEL;

    float3 ro = o, rd = d;
    const float3 winv = safe_inv_dir(d);
    float3 inv = winv;
    const bool wpar = has_parallel_axis(d);
    bool par = wpar;
#if !VKP_WIDE
    const SlabRay wslab = slab_setup(o, winv);
    SlabRay slab = wslab;
#endif
#if VKP_WIDE
    const float4* nodes = S.tlas_nodes4;
    uint32_t cur = S.tlas_root_ref;
#else
    const float4* nodes = S.tlas_nodes2;
    uint32_t cur = S.tlas_root_ref2;
#endif
    const float4* tris = nullptr;
    int inst = -1;

    // "while-while" traversal (Aila & Laine) over 4-wide nodes: all lanes of a warp walk internal nodes together and only
    // then process their leaves together.  One trip = one 128-byte node fetch and four slab tests; the hit children other
    // than the nearest go on the stack, farthest first.
    while (true) {
#if !VKP_WIDE
        while (!(cur & VKP_LEAF_BIT)) {
            if (LIMITED) {
                if (++n_limited == max_steps) {
                    const unsigned int slot = atomicAdd(spill_count, 1u);
                    if (slot < spill_cap) { best.prim = VKP_TRACE_SPILLED; best.inst = (int)slot; return false; }
                }
            }
            const float4* c = nodes + 2u * (size_t)cur;      // both children: 64 contiguous bytes
            if (STATS) n_steps++;
            if (tcount) tcount->nodes++;
            float4 lo0 = __ldg(c + 0), hi0 = __ldg(c + 1), lo1 = __ldg(c + 2), hi1 = __ldg(c + 3);
            float t0, t1;
            bool h0, h1;
#if !VKP_SLAB_FMA
            if (true) {
                h0 = box_test(lo0, hi0, ro, rd, inv, par, tmin, limit, t0);
                h1 = box_test(lo1, hi1, ro, rd, inv, par, tmin, limit, t1);
            } else {
#else
            if (par) {      // rare: a direction component is zero
                h0 = box_test<true>(lo0, hi0, ro, rd, inv, tmin, limit, t0);
                h1 = box_test<true>(lo1, hi1, ro, rd, inv, tmin, limit, t1);
            } else {
#endif
                h0 = box_test_fast(lo0, hi0, slab, tmin, limit, t0);
                h1 = box_test_fast(lo1, hi1, slab, tmin, limit, t1);
            }
            uint32_t r0 = __float_as_uint(lo0.w), r1 = __float_as_uint(lo1.w);
            if (h0 && h1) {
                bool swap = t1 < t0;
                stack[sp++] = swap ? r0 : r1;
                cur = swap ? r1 : r0;
            } else if (h0) cur = r0;
            else if (h1) cur = r1;
            else cur = stack[--sp];
        }
#else
        while (!(cur & VKP_LEAF_BIT)) {
            const float4* c = nodes + 8u * (size_t)cur;
            if (STATS) n_steps++;
            if (tcount) tcount->nodes++;
            const float4 lox = __ldg(c + 0), loy = __ldg(c + 1), loz = __ldg(c + 2), hix = __ldg(c + 3), hiy = __ldg(c + 4), hiz = __ldg(c + 5), rf = __ldg(c + 6);
            const float inf = __int_as_float(0x7f800000);
            float t0, t1, t2, t3;
            const uint32_t r0 = __float_as_uint(rf.x), r1 = __float_as_uint(rf.y), r2 = __float_as_uint(rf.z), r3 = __float_as_uint(rf.w);
            bool h0 = box_test(make_float4(lox.x, loy.x, loz.x, 0.f), make_float4(hix.x, hiy.x, hiz.x, 0.f), ro, rd, inv, par, tmin, limit, t0);
            bool h1 = box_test(make_float4(lox.y, loy.y, loz.y, 0.f), make_float4(hix.y, hiy.y, hiz.y, 0.f), ro, rd, inv, par, tmin, limit, t1);
            bool h2 = box_test(make_float4(lox.z, loy.z, loz.z, 0.f), make_float4(hix.z, hiy.z, hiz.z, 0.f), ro, rd, inv, par, tmin, limit, t2) && r2 != VKP_W_EMPTY_REF;
            bool h3 = box_test(make_float4(lox.w, loy.w, loz.w, 0.f), make_float4(hix.w, hiy.w, hiz.w, 0.f), ro, rd, inv, par, tmin, limit, t3) && r3 != VKP_W_EMPTY_REF;
            // sort the (distance, ref) pairs ascending with a 5-comparator network; misses sink to the end as +inf
            float d0 = h0 ? t0 : inf, d1 = h1 ? t1 : inf, d2 = h2 ? t2 : inf, d3 = h3 ? t3 : inf;
            uint32_t q0 = r0, q1 = r1, q2 = r2, q3 = r3;
#define VKP_CSWAP(da, qa, db, qb) { if (db < da) { float td = da; da = db; db = td; uint32_t tq = qa; qa = qb; qb = tq; } }
            VKP_CSWAP(d0, q0, d1, q1) VKP_CSWAP(d2, q2, d3, q3) VKP_CSWAP(d0, q0, d2, q2) VKP_CSWAP(d1, q1, d3, q3) VKP_CSWAP(d1, q1, d2, q2)
#undef VKP_CSWAP
            if (d3 < inf) stack[sp++] = q3;
            if (d2 < inf) stack[sp++] = q2;
            if (d1 < inf) stack[sp++] = q1;
            cur = (d0 < inf) ? q0 : stack[--sp];
        }
#endif
        if (cur == VKP_STACK_SENTINEL) break;
        if (cur == VKP_STACK_RETURN) {
            inst = -1; ro = o; rd = d; inv = winv; par = wpar;
#if !VKP_WIDE
            slab = wslab;
#endif
#if VKP_WIDE
            nodes = S.tlas_nodes4;
#else
            nodes = S.tlas_nodes2;
#endif
            cur = stack[--sp];
            continue;
        }
        if (inst < 0) {
            // ---- TLAS leaf: one instance; its id rides in the leaf reference, its record is one 80-byte fetch ----------------
            inst = (int)((cur & 0x7fffffffu) >> 3);
            if (tcount) tcount->entries++;
            const float4* rec = reinterpret_cast<const float4*>(S.inst_trav + inst);
            const float4 r0 = __ldg(rec), r1 = __ldg(rec + 1), r2 = __ldg(rec + 2), pp = __ldg(rec + 3), rr = __ldg(rec + 4);
            // xform_point / xform_dir with the rows of inverse(model): same association as the contract
            ro = f3(vkp_row_point(r0.x, r0.y, r0.z, r0.w, o.x, o.y, o.z), vkp_row_point(r1.x, r1.y, r1.z, r1.w, o.x, o.y, o.z), vkp_row_point(r2.x, r2.y, r2.z, r2.w, o.x, o.y, o.z));
            rd = f3(vkp_dot3(r0.x, r0.y, r0.z, d.x, d.y, d.z), vkp_dot3(r1.x, r1.y, r1.z, d.x, d.y, d.z), vkp_dot3(r2.x, r2.y, r2.z, d.x, d.y, d.z));
            inv = safe_inv_dir(rd);
            par = has_parallel_axis(rd);
#if !VKP_WIDE
            slab = slab_setup(ro, inv);
#endif
            tris = reinterpret_cast<const float4*>(((unsigned long long)__float_as_uint(pp.w) << 32) | (unsigned long long)__float_as_uint(pp.z));
            // the TLAS leaf box is the world box of this instance: go straight to the BLAS root's content
            stack[sp++] = VKP_STACK_RETURN;
#if VKP_WIDE
            nodes = reinterpret_cast<const float4*>(((unsigned long long)__float_as_uint(pp.y) << 32) | (unsigned long long)__float_as_uint(pp.x));
            cur = __float_as_uint(rr.x);
#else
            nodes = reinterpret_cast<const float4*>(((unsigned long long)__float_as_uint(rr.w) << 32) | (unsigned long long)__float_as_uint(rr.z));
            cur = __float_as_uint(rr.y);
#endif
            continue;
        }
        // ---- BLAS leaf: <= 4 triangles; issue all loads first so that their latency overlaps ----------------------------------
        {
            const uint32_t first = (cur & 0x7fffffffu) >> 3, count = (cur & 7u) + 1u;
            const float4* tp = tris + 3u * (size_t)first;
            float4 ta[4], tb[4], tc[4];
#pragma unroll
            for (uint32_t k = 0; k < 4; k++)
                if (k < count) { ta[k] = __ldg(tp + 3 * k); tb[k] = __ldg(tp + 3 * k + 1); tc[k] = __ldg(tp + 3 * k + 2); }
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) {
                if (k < count) {
                    if (STATS) n_tris++;
                    if (tcount) tcount->tris++;
                    float t, u, v;
                    if (tri_test(ro, rd, f3(ta[k].x, ta[k].y, ta[k].z), f3(tb[k].x, tb[k].y, tb[k].z), f3(tc[k].x, tc[k].y, tc[k].z), tmin, tmax, t, u, v)) {
                        int prim = __float_as_int(ta[k].w);
                        if (ANY) { best.t = t; best.u = u; best.v = v; best.inst = inst; best.prim = prim; if (STATS) { stats[0] = n_steps; stats[1] = n_tris; } return true; }
                        if (hit_better(t, inst, prim, best)) {
                            best.t = t; best.u = u; best.v = v; best.inst = inst; best.prim = prim;
                            limit = t;
                        }
                    }
                }
            }
            cur = stack[--sp];
        }
    }
    if (STATS) { stats[0] = n_steps; stats[1] = n_tris; }
    return best.inst >= 0;
}


#if !VKP_WIDE
// ---- warp-cooperative any-hit traversal (the tail of k_shadow) --------------------------------------------------------------------
// One warp, ONE ray.  An any-hit query is the OR over every leaf the ray's interval reaches, so the order of the walk is free:
// each lane keeps its own stack of (node reference, instance) items and walks depth-first; a lane that runs dry takes the OLDEST
// item (the root of the largest unexplored subtree) of a lane that still has some, through two shuffles.  Items carry their
// instance (-1 = TLAS level), so there are no "return to the TLAS" markers and a lane re-derives the object-space ray only when
// the instance of its item changes.  A ray that needs N child-pair fetches finishes in about N / 32 + tree depth trips instead
// of N.  Same box test, same triangle test, same far bound as trace_ray<true>: the answer is the same boolean.
static __device__ __forceinline__ bool warp_any_hit(const DScene& S, float3 o, float3 d, float tmin, float tmax, TravCount* tcount)
{
    const unsigned int full = 0xffffffffu;
    const int lane = (int)(threadIdx.x & 31u);
    const unsigned int lt = (1u << lane) - 1u;
    if (S.n_instances == 0) return false;

    uint2 stack[VKP_TRAV_STACK];
    int base = 0, sp = 0;                    // live items: stack[base .. sp)
    bool has = false;                        // a current item
    uint32_t cur = 0u;
    int cinst = -1;
    int finst = -2;                          // instance the cached object-space ray belongs to (-1 world, -2 none yet)
    float3 ro = o, rd = d;
    const float3 winv = safe_inv_dir(d);
    const bool wpar = has_parallel_axis(d);
    float3 inv = winv;
    bool par = wpar;
    const float4* nodes = S.tlas_nodes2;
    const float4* tris = nullptr;
    if (lane == 0) { has = true; cur = S.tlas_root_ref2; cinst = -1; }

    while (true) {
        if (!has && sp > base) {
            const uint2 e = stack[--sp];
            cur = e.x; cinst = (int)e.y; has = true;
            if (sp == base) sp = base = 0;
        }
        bool found = false;
        if (has) {
            if (cinst != finst) {
                finst = cinst;
                if (cinst < 0) { ro = o; rd = d; inv = winv; par = wpar; nodes = S.tlas_nodes2; tris = nullptr; }
                else {
                    const float4* rec = reinterpret_cast<const float4*>(S.inst_trav + cinst);
                    const float4 r0 = __ldg(rec), r1 = __ldg(rec + 1), r2 = __ldg(rec + 2), pp = __ldg(rec + 3), rr = __ldg(rec + 4);
                    ro = f3(vkp_row_point(r0.x, r0.y, r0.z, r0.w, o.x, o.y, o.z), vkp_row_point(r1.x, r1.y, r1.z, r1.w, o.x, o.y, o.z), vkp_row_point(r2.x, r2.y, r2.z, r2.w, o.x, o.y, o.z));
                    rd = f3(vkp_dot3(r0.x, r0.y, r0.z, d.x, d.y, d.z), vkp_dot3(r1.x, r1.y, r1.z, d.x, d.y, d.z), vkp_dot3(r2.x, r2.y, r2.z, d.x, d.y, d.z));
                    inv = safe_inv_dir(rd);
                    par = has_parallel_axis(rd);
                    tris = reinterpret_cast<const float4*>(((unsigned long long)__float_as_uint(pp.w) << 32) | (unsigned long long)__float_as_uint(pp.z));
                    nodes = reinterpret_cast<const float4*>(((unsigned long long)__float_as_uint(rr.w) << 32) | (unsigned long long)__float_as_uint(rr.z));
                }
            }
            if (!(cur & VKP_LEAF_BIT)) {
                const float4* c = nodes + 2u * (size_t)cur;
                if (tcount) tcount->nodes++;
                const float4 lo0 = __ldg(c + 0), hi0 = __ldg(c + 1), lo1 = __ldg(c + 2), hi1 = __ldg(c + 3);
                float t0, t1;
                const bool h0 = box_test(lo0, hi0, ro, rd, inv, par, tmin, tmax, t0);
                const bool h1 = box_test(lo1, hi1, ro, rd, inv, par, tmin, tmax, t1);
                const uint32_t r0 = __float_as_uint(lo0.w), r1 = __float_as_uint(lo1.w);
                if (h0 && h1) {
                    if (sp == VKP_TRAV_STACK) {      // the window [base, sp) only moves up while items are given away: slide it back down
                        for (int k = base; k < sp; k++) stack[k - base] = stack[k];
                        sp -= base; base = 0;
                    }
                    stack[sp++] = make_uint2(r1, (unsigned int)cinst); cur = r0;
                }
                else if (h0) cur = r0;
                else if (h1) cur = r1;
                else has = false;
            } else if (cinst < 0) {
                // TLAS leaf: the item becomes (root of that instance's BLAS, instance); the record is read on the next trip
                cinst = (int)((cur & 0x7fffffffu) >> 3);
                if (tcount) tcount->entries++;
                const float4 rr = __ldg(reinterpret_cast<const float4*>(S.inst_trav + cinst) + 4);
                cur = __float_as_uint(rr.y);
            } else {
                const uint32_t first = (cur & 0x7fffffffu) >> 3, count = (cur & 7u) + 1u;
                const float4* tp = tris + 3u * (size_t)first;
                float4 ta[4], tb[4], tc[4];
#pragma unroll
                for (uint32_t k = 0; k < 4; k++)
                    if (k < count) { ta[k] = __ldg(tp + 3 * k); tb[k] = __ldg(tp + 3 * k + 1); tc[k] = __ldg(tp + 3 * k + 2); }
#pragma unroll
                for (uint32_t k = 0; k < 4; k++) {
                    if (k < count) {
                        if (tcount) tcount->tris++;
                        float t, u, v;
                        if (tri_test(ro, rd, f3(ta[k].x, ta[k].y, ta[k].z), f3(tb[k].x, tb[k].y, tb[k].z), f3(tc[k].x, tc[k].y, tc[k].z), tmin, tmax, t, u, v)) found = true;
                    }
                }
                has = false;
            }
        }
        if (__any_sync(full, found)) return true;
        const bool dry = !has && sp == base;
        const unsigned int need = __ballot_sync(full, dry);
        if (need == full) return false;                       // nobody has anything left: unoccluded
        if (need) {
            const bool can = has && sp > base;               // keeps its current item, gives the oldest one on its stack
            const unsigned int have = __ballot_sync(full, can);
            if (have) {
                const int n = min(__popc(need), __popc(have));
                const int rank_give = __popc(have & lt), rank_need = __popc(need & lt);
                uint2 give = make_uint2(0u, 0u);
                if (can && rank_give < n) { give = stack[base++]; if (sp == base) sp = base = 0; }
                const bool take = dry && rank_need < n;
                const int src = take ? (int)__fns(have, 0u, rank_need + 1) : lane;
                const uint32_t gx = __shfl_sync(full, give.x, src), gy = __shfl_sync(full, give.y, src);
                if (take) { cur = gx; cinst = (int)gy; has = true; }
            }
        }
    }
}
#endif   // !VKP_WIDE
