// vkp_bvh.cuh -- two-level BVH traversal (what the Vulkan driver + RT cores do behind traceRayEXT:
// shadowRaygen.rgen:78, hybridRaygen.rgen:158, hybridHit.rchit:123, raygen.rgen:49, closesthit.rchit:111).
//
// Layout (DESIGN.md "BVH layout"): 32-byte nodes {bmin, ref, bmax, aux}.  Node 0 is the root.  An internal
// node's `ref` is the index of its left child; the right child is ref+1, so one 64-byte aligned fetch brings
// both children.  A leaf's ref has bit 31 set: bits[30:3] = first item, bits[2:0] = count-1.  BLAS items are
// triangles stored as 3 x float4 in leaf order (v0 | primitive id, v1, v2); TLAS items are instance ids.
//
// Intersection contract: Moeller-Trumbore in object space, no culling, tmin < t < tmax, u >= 0, v >= 0,
// u + v <= 1; closest hit = min t, then lower instance id, then lower primitive id.  The arithmetic of
// tri_test() is the contract (contract v2: dot / cross are the FMA forms of include/vkp_detmath.h) -- do not
// re-associate it.  The box test is NOT part of the contract: it only has to be conservative.
//
// Round-2 traversal:
//  * the stack lives in SHARED memory, one column per thread (s[level * VKP_TRAV_BLOCK], conflict-free), VKP_SSTACK levels
//    deep; deeper levels -- rare -- go to a small local-memory overflow array.  No LDL / STL in the node loop.
//  * the slab test is 6 FMAs per box: t = b * (1/d) - o * (1/d), near / far plane picked by the sign of the direction,
//    the rounding of the precomputed o * (1/d) folded into two copies of it (near planes earlier, far planes later).
//  * nothing of the world-space ray stays in registers while a BLAS is walked: leaving an instance recomputes the world
//    slab from the ray (three IEEE reciprocals, once per instance visit) instead of holding 12 more registers in the loop.
//  * rays with a zero direction component -- in world space or inside an instance -- restart in trace_ray_general,
//    the exact "parallel to a slab" formulation of round 1, out of line.
#pragma once
#include "vkp_device.cuh"

#define VKP_LEAF_BIT 0x80000000u
#define VKP_STACK_SENTINEL 0xffffffffu
#define VKP_STACK_RETURN 0xfffffffeu
#define VKP_TRAV_STACK 64            /* deepest stack a ray may need: checked against the tree heights when the BVHs are built */
#define VKP_TRAV_BLOCK 128           /* threads per CTA of every kernel that traces rays */
#ifndef VKP_SSTACK
#define VKP_SSTACK 16                /* stack levels kept in shared memory: 16 * 128 * 4 B = 8 KB per CTA */
#endif
#define VKP_TRAV_SMEM_WORDS (VKP_SSTACK * VKP_TRAV_BLOCK)
#ifndef VKP_PF_CLOSEST
#define VKP_PF_CLOSEST 0             /* 1: closest-hit rays prefetch the far child they push (measured: no gain) */
#endif
#ifndef VKP_PF_ANY
#define VKP_PF_ANY 0                 /* 1: any-hit rays too */
#endif

struct HitRec {
    float t, u, v;
    int inst, prim;
};

// what a thread's rays actually requested from memory: node fetches (64 B each), triangle tests (48 B)
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
        // the min / max below as an empty interval.  The slack keeps the test conservative against the rounding of the
        // triangle test for rays that run along a face of the box.
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
// form, worth a spatial error of |o| * 2^-24 whatever the direction -- folded into two copies of it: oin = o/d + e for the near
// planes, oif = o/d - e for the far ones, e = |o/d| * 2^-22.  6 FFMA + 6 selects + 2 three-input min / max + 2 min / max + the
// slack multiply and the compare per box.  Only for rays without a zero direction component (1/d finite).
struct SlabRay { float3 inv, oin, oif; bool nx, ny, nz; };
VKP_D SlabRay slab_setup(float3 o, float3 d)
{
    SlabRay R;
    R.inv = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    const float ox = o.x * R.inv.x, oy = o.y * R.inv.y, oz = o.z * R.inv.z;
    const float ex = fabsf(ox) * 2.4e-7f, ey = fabsf(oy) * 2.4e-7f, ez = fabsf(oz) * 2.4e-7f;
    R.oin = f3(ox + ex, oy + ey, oz + ez);
    R.oif = f3(ox - ex, oy - ey, oz - ez);
    R.nx = d.x < 0.0f; R.ny = d.y < 0.0f; R.nz = d.z < 0.0f;
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

// zero (or denormal, or NaN) direction components go to the general path: 1/d is not finite there and the exact
// "parallel to a slab" logic is needed
VKP_D bool needs_general_path(float3 d) { return !(fabsf(d.x) > 1e-30f && fabsf(d.y) > 1e-30f && fabsf(d.z) > 1e-30f); }
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

// object-space ray of an instance: the rows of inverse(model) applied with the contract's matrix-row primitives
VKP_D void enter_instance(const float4 r0, const float4 r1, const float4 r2, float3 o, float3 d, float3& ro, float3& rd)
{
    ro = f3(vkp_row_point(r0.x, r0.y, r0.z, r0.w, o.x, o.y, o.z), vkp_row_point(r1.x, r1.y, r1.z, r1.w, o.x, o.y, o.z), vkp_row_point(r2.x, r2.y, r2.z, r2.w, o.x, o.y, o.z));
    rd = f3(vkp_dot3(r0.x, r0.y, r0.z, d.x, d.y, d.z), vkp_dot3(r1.x, r1.y, r1.z, d.x, d.y, d.z), vkp_dot3(r2.x, r2.y, r2.z, d.x, d.y, d.z));
}
VKP_D const float4* ptr_from_words(float lo, float hi)
{
    return reinterpret_cast<const float4*>(((unsigned long long)__float_as_uint(hi) << 32) | (unsigned long long)__float_as_uint(lo));
}

// ---- the general path (round 1's traversal): any ray, including directions with zero components; local-memory stack ---------------
// ANY == true: gl_RayFlagsTerminateOnFirstHitEXT | SkipClosestHitShaderEXT (shadow rays): returns on the first hit
template <bool ANY>
__device__ __noinline__ bool trace_ray_general(const DScene& S, float3 o, float3 d, float tmin, float tmax, HitRec& best, TravCount* tcount)
{
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
    const float4* nodes = S.tlas_nodes2;
    uint32_t cur = S.tlas_root_ref2;
    const float4* tris = nullptr;
    int inst = -1;
    while (true) {
        while (!(cur & VKP_LEAF_BIT)) {
            const float4* c = nodes + 2u * (size_t)cur;      // both children: 64 contiguous bytes
            if (tcount) tcount->nodes++;
            const float4 lo0 = __ldg(c + 0), hi0 = __ldg(c + 1), lo1 = __ldg(c + 2), hi1 = __ldg(c + 3);
            float t0, t1;
            const bool h0 = box_test(lo0, hi0, ro, rd, inv, par, tmin, limit, t0);
            const bool h1 = box_test(lo1, hi1, ro, rd, inv, par, tmin, limit, t1);
            const uint32_t r0 = __float_as_uint(lo0.w), r1 = __float_as_uint(lo1.w);
            if (h0 && h1) {
                const bool swap = t1 < t0;
                stack[sp++] = swap ? r0 : r1;
                cur = swap ? r1 : r0;
            } else if (h0) cur = r0;
            else if (h1) cur = r1;
            else cur = stack[--sp];
        }
        if (cur == VKP_STACK_SENTINEL) break;
        if (cur == VKP_STACK_RETURN) {
            inst = -1; ro = o; rd = d; inv = winv; par = wpar;
            nodes = S.tlas_nodes2;
            cur = stack[--sp];
            continue;
        }
        if (inst < 0) {
            inst = (int)((cur & 0x7fffffffu) >> 3);
            if (tcount) tcount->entries++;
            const float4* rec = reinterpret_cast<const float4*>(S.inst_trav + inst);
            const float4 r0 = __ldg(rec), r1 = __ldg(rec + 1), r2 = __ldg(rec + 2), pp = __ldg(rec + 3);
            enter_instance(r0, r1, r2, o, d, ro, rd);
            inv = safe_inv_dir(rd);
            par = has_parallel_axis(rd);
            nodes = ptr_from_words(pp.x, pp.y);
            tris = ptr_from_words(pp.z, pp.w);
            stack[sp++] = VKP_STACK_RETURN;
            cur = __float_as_uint(__ldg(rec + 4).x);
            continue;
        }
        {
            const uint32_t first = (cur & 0x7fffffffu) >> 3, count = (cur & 7u) + 1u;
            const float4* tp = tris + 3u * (size_t)first;
            for (uint32_t k = 0; k < count; k++) {
                const float4 ta = __ldg(tp + 3 * k), tb = __ldg(tp + 3 * k + 1), tc = __ldg(tp + 3 * k + 2);
                if (tcount) tcount->tris++;
                float t, u, v;
                if (tri_test(ro, rd, f3(ta.x, ta.y, ta.z), f3(tb.x, tb.y, tb.z), f3(tc.x, tc.y, tc.z), tmin, tmax, t, u, v)) {
                    const int prim = __float_as_int(ta.w);
                    if (ANY) { best.t = t; best.u = u; best.v = v; best.inst = inst; best.prim = prim; return true; }
                    if (hit_better(t, inst, prim, best)) { best.t = t; best.u = u; best.v = v; best.inst = inst; best.prim = prim; limit = t; }
                }
            }
            cur = stack[--sp];
        }
    }
    return best.inst >= 0;
}

// ---- the fast path ---------------------------------------------------------------------------------------------------------------
// `sstack` = this thread's column of the CTA's shared-memory stack: sstack[level * VKP_TRAV_BLOCK], VKP_SSTACK levels.
// STATS: stats[0] = child-pair fetches, stats[1] = triangle tests of this ray (vkp_trace_stats).
// KEEPW: the world-space slab stays in registers while a BLAS is walked instead of being recomputed (three IEEE reciprocals) when the
// walk returns to the TLAS -- 12 more live registers against ~30 instructions per instance visit.  Pays in the issue-bound kernels
// that have registers to spare (many-sample shadows, Raytraced mode), costs in the 64-register latency-bound ones.
template <bool ANY, bool STATS = false, bool KEEPW = false>
__device__ __forceinline__ bool trace_ray(const DScene& S, uint32_t* __restrict__ sstack, float3 o, float3 d, float tmin, float tmax, HitRec& best,
                                          uint32_t* stats = nullptr, TravCount* tcount = nullptr)
{
    uint32_t n_steps = 0, n_tris = 0;
    best.t = -1.0f; best.u = 0.0f; best.v = 0.0f; best.inst = -1; best.prim = -1;
    if (S.n_instances == 0) { if (STATS) { stats[0] = 0u; stats[1] = 0u; } return false; }
    bool general = needs_general_path(d);
    if (!general) {
        float limit = tmax;
        uint32_t ovf[VKP_TRAV_STACK - VKP_SSTACK];          // levels beyond the shared-memory part (rare)
        int sp = 0;
        // the 32-bit shared-memory address of this thread's column, pinned in a register: left alone the compiler re-derives it
        // from %tid and the CTA's shared window at every push and pop (S2R + S2UR + 3 address instructions, seen in the SASS)
        uint32_t scol = (uint32_t)__cvta_generic_to_shared(sstack);
        asm volatile("" : "+r"(scol));
#define VKP_PUSH(x) do { const uint32_t _v = (x); if (sp < VKP_SSTACK) asm volatile("st.shared.u32 [%0], %1;" :: "r"(scol + (uint32_t)sp * (4u * VKP_TRAV_BLOCK)), "r"(_v)); \
                         else ovf[sp - VKP_SSTACK] = _v; sp++; } while (0)
#define VKP_POP(dst) do { --sp; if (sp < VKP_SSTACK) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(dst) : "r"(scol + (uint32_t)sp * (4u * VKP_TRAV_BLOCK))); \
                          else dst = ovf[sp - VKP_SSTACK]; } while (0)
        const unsigned int cnt_inc = tcount ? 1u : 0u;      // traversal accounting (profiling level 2): one add per trip, no branch
        unsigned int cnt_nodes = 0u;
        VKP_PUSH(VKP_STACK_SENTINEL);
        float3 ro = o, rd = d;
        SlabRay slab = slab_setup(o, d);
        SlabRay wslab;
        if (KEEPW) wslab = slab;
        const float4* nodes = S.tlas_nodes2;
        uint32_t cur = S.tlas_root_ref2;
        const float4* tris = nullptr;
        int inst = -1;
        while (true) {
            // "while-while" traversal (Aila & Laine): the lanes of a warp walk internal nodes together and only then their leaves
            while (!(cur & VKP_LEAF_BIT)) {
                const float4* c = nodes + 2u * (size_t)cur;      // both children: 64 contiguous bytes
                if (STATS) n_steps++;
                cnt_nodes += cnt_inc;
                const float4 lo0 = __ldg(c + 0), hi0 = __ldg(c + 1), lo1 = __ldg(c + 2), hi1 = __ldg(c + 3);
                float t0, t1;
                const bool h0 = box_test_fast(lo0, hi0, slab, tmin, limit, t0);
                const bool h1 = box_test_fast(lo1, hi1, slab, tmin, limit, t1);
                const uint32_t r0 = __float_as_uint(lo0.w), r1 = __float_as_uint(lo1.w);
                if (h0 && h1) {
                    const bool swap = t1 < t0;
                    const uint32_t far_ref = swap ? r0 : r1;
                    VKP_PUSH(far_ref);
                    // (A/B knob, off: measured slower) the far child is walked after the whole near subtree -- ask for its child pair (or its
                    // first triangles) now, so that the pop finds it in L1 instead of starting a dependent L2 round trip
                    if (ANY ? (VKP_PF_ANY != 0) : (VKP_PF_CLOSEST != 0)) {
                        const float4* pf = (far_ref & VKP_LEAF_BIT) ? (inst >= 0 ? tris + 3u * (size_t)((far_ref & 0x7fffffffu) >> 3) : nullptr)
                                                                    : nodes + 2u * (size_t)far_ref;
                        if (pf) asm volatile("prefetch.global.L1 [%0];" :: "l"(pf));
                    }
                    cur = swap ? r1 : r0;
                } else if (h0) cur = r0;
                else if (h1) cur = r1;
                else VKP_POP(cur);
            }
            if (cur == VKP_STACK_SENTINEL) break;
            if (cur == VKP_STACK_RETURN) {
                // back in the TLAS: the world slab is recomputed rather than carried through the BLAS walk in registers
                inst = -1; ro = o; rd = d;
                if (KEEPW) slab = wslab; else slab = slab_setup(o, d);
                nodes = S.tlas_nodes2;
                VKP_POP(cur);
                continue;
            }
            if (inst < 0) {
                // ---- TLAS leaf: one instance; its id rides in the leaf reference, its record is one 80-byte fetch ----------------
                inst = (int)((cur & 0x7fffffffu) >> 3);
                if (tcount) tcount->entries++;
                const float4* rec = reinterpret_cast<const float4*>(S.inst_trav + inst);
                const float4 r0 = __ldg(rec), r1 = __ldg(rec + 1), r2 = __ldg(rec + 2), pp = __ldg(rec + 3), rr = __ldg(rec + 4);
                enter_instance(r0, r1, r2, o, d, ro, rd);
                if (needs_general_path(rd)) { general = true; break; }      // restart out of line (rare: a rotated instance met along an axis)
                slab = slab_setup(ro, rd);
                nodes = ptr_from_words(pp.x, pp.y);
                tris = ptr_from_words(pp.z, pp.w);
                // the TLAS leaf box is the world box of this instance: go straight to the BLAS root's content
                VKP_PUSH(VKP_STACK_RETURN);
                cur = __float_as_uint(rr.x);
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
                            const int prim = __float_as_int(ta[k].w);
                            if (ANY) {
                                best.t = t; best.u = u; best.v = v; best.inst = inst; best.prim = prim;
                                if (STATS) { stats[0] = n_steps; stats[1] = n_tris; }
                                if (tcount) tcount->nodes += cnt_nodes;
                                return true;
                            }
                            if (hit_better(t, inst, prim, best)) {
                                best.t = t; best.u = u; best.v = v; best.inst = inst; best.prim = prim;
                                limit = t;
                            }
                        }
                    }
                }
                VKP_POP(cur);
            }
        }
#undef VKP_PUSH
#undef VKP_POP
        if (tcount) tcount->nodes += cnt_nodes;
    }
    if (general) {
        const bool hit = trace_ray_general<ANY>(S, o, d, tmin, tmax, best, tcount);
        if (STATS) { stats[0] = n_steps; stats[1] = n_tris; }
        return hit;
    }
    if (STATS) { stats[0] = n_steps; stats[1] = n_tris; }
    return best.inst >= 0;
}

// ---- warp-cooperative traversal: ONE ray, 32 lanes (the reflection chain, k_chain) ---------------------------------------------------
// A reflection / refraction path is a chain of dependent rays, and its longest single ray (a reflection grazing the 328 k-triangle
// statue walks hundreds of nodes, one dependent fetch after the other) bounds the whole kernel.  Here the warp walks the tree of ONE
// ray together: a per-warp stack of (node reference, instance, entry distance) items lives in shared memory; every trip the lanes
// take the (up to) 32 topmost items, each lane tests the two children of its node (or the <= 4 triangles of its leaf, or enters the
// instance of its TLAS leaf) and the survivors are pushed back with ballot + popc offsets, far children below near children so the
// next trip continues with the nearest ones.  A ray of N node visits finishes in about N / (lanes busy) + depth trips instead of N.
//
// The result is the same as trace_ray's, bit for bit: the closest hit is the minimum over ALL accepted triangles under the total
// order (t, instance, primitive), whatever order they are found in; a box is only skipped when it lies behind a hit already found
// (conservatively, with the slack of the serial test); any-hit is the OR over the reachable leaves.
// Every lane passes the same ray and gets the same answer.  Rays the fast slab test cannot take (a zero direction component, in
// the world or inside an instance) and a stack that would overflow fall back to trace_ray_general, run redundantly by all lanes.
#ifndef VKP_WQ_CAP
#define VKP_WQ_CAP 384                                   /* items per warp: 384 * 12 B = 4.5 KB, 18 KB per 4-warp CTA */
#endif
#define VKP_WQ_SMEM_WORDS (3 * VKP_WQ_CAP * (VKP_TRAV_BLOCK / 32))
struct WarpQueue { uint32_t* ref; int* inst; float* tn; };
VKP_D WarpQueue warp_queue_of(uint32_t* smem_words)      // smem_words: VKP_WQ_SMEM_WORDS words of the CTA
{
    uint32_t* base = smem_words + (threadIdx.x >> 5) * (3 * VKP_WQ_CAP);
    WarpQueue Q; Q.ref = base; Q.inst = reinterpret_cast<int*>(base + VKP_WQ_CAP); Q.tn = reinterpret_cast<float*>(base + 2 * VKP_WQ_CAP);
    return Q;
}

template <bool ANY>
__device__ __forceinline__ bool warp_trace(const DScene& S, const WarpQueue& Q, float3 o, float3 d, float tmin, float tmax, HitRec& best, TravCount* tcount)
{
    const unsigned int full = 0xffffffffu;
    const int lane = (int)(threadIdx.x & 31u);
    const unsigned int lt = (1u << lane) - 1u;
    best.t = -1.0f; best.u = 0.0f; best.v = 0.0f; best.inst = -1; best.prim = -1;
    if (S.n_instances == 0) return false;
    bool fallback = needs_general_path(d);                // warp-uniform: every lane holds the same ray
    bool any_found = false;
    if (!fallback) {
        float limit = tmax;
        int n = 1;
        if (lane == 0) { Q.ref[0] = S.tlas_root_ref2; Q.inst[0] = -1; Q.tn[0] = tmin; }
        __syncwarp();
        int cinst = -1;                                   // the instance the lane's cached ray belongs to (-1 = world)
        float3 ro = o, rd = d;
        SlabRay slab = slab_setup(o, d);
        const float4* nodes = S.tlas_nodes2;
        const float4* tris = nullptr;
        HitRec mine; mine.t = 0.0f; mine.u = 0.0f; mine.v = 0.0f; mine.inst = -1; mine.prim = -1;
        while (n > 0) {
            const int take = n < 32 ? n : 32, base = n - take;
            bool have = lane < take;
            uint32_t ref = 0u; int inst = -1; float tn = 0.0f;
            if (have) { ref = Q.ref[base + lane]; inst = Q.inst[base + lane]; tn = Q.tn[base + lane]; }
            n = base;
            __syncwarp();                                 // the items are read before anyone pushes over them
            if (have && tn > limit * 1.000001f) have = false;      // behind a hit found since the item was pushed
            bool p0 = false, p1 = false, found = false, bad = false;
            uint32_t q0 = 0u, q1 = 0u; int qi1 = inst; float qt0 = 0.0f, qt1 = 0.0f;
            if (have) {
                if (inst != cinst) {
                    cinst = inst;
                    if (inst < 0) { ro = o; rd = d; slab = slab_setup(o, d); nodes = S.tlas_nodes2; tris = nullptr; }
                    else {
                        const float4* rec = reinterpret_cast<const float4*>(S.inst_trav + inst);
                        const float4 r0 = __ldg(rec), r1 = __ldg(rec + 1), r2 = __ldg(rec + 2), pp = __ldg(rec + 3);
                        enter_instance(r0, r1, r2, o, d, ro, rd);
                        if (needs_general_path(rd)) bad = true;
                        else { slab = slab_setup(ro, rd); nodes = ptr_from_words(pp.x, pp.y); tris = ptr_from_words(pp.z, pp.w); }
                    }
                }
                if (!bad) {
                    if (!(ref & VKP_LEAF_BIT)) {
                        const float4* c = nodes + 2u * (size_t)ref;
                        if (tcount) tcount->nodes++;
                        const float4 lo0 = __ldg(c + 0), hi0 = __ldg(c + 1), lo1 = __ldg(c + 2), hi1 = __ldg(c + 3);
                        float t0, t1;
                        const bool h0 = box_test_fast(lo0, hi0, slab, tmin, limit, t0);
                        const bool h1 = box_test_fast(lo1, hi1, slab, tmin, limit, t1);
                        const uint32_t r0 = __float_as_uint(lo0.w), r1 = __float_as_uint(lo1.w);
                        const bool swap = h1 && (!h0 || t1 < t0);           // child 1 is the nearer (or the only) hit
                        p1 = h0 || h1; p0 = h0 && h1;
                        q1 = swap ? r1 : r0; qt1 = swap ? t1 : t0;           // near child: pushed on top
                        q0 = swap ? r0 : r1; qt0 = swap ? t0 : t1;           // far child: below it
                    } else if (inst < 0) {
                        // TLAS leaf: the item becomes the root of that instance's BLAS (its record is read when a lane takes it)
                        qi1 = (int)((ref & 0x7fffffffu) >> 3);
                        if (tcount) tcount->entries++;
                        q1 = __float_as_uint(__ldg(reinterpret_cast<const float4*>(S.inst_trav + qi1) + 4).x);
                        qt1 = tn; p1 = true;
                    } else {
                        const uint32_t first = (ref & 0x7fffffffu) >> 3, count = (ref & 7u) + 1u;
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
                                if (tri_test(ro, rd, f3(ta[k].x, ta[k].y, ta[k].z), f3(tb[k].x, tb[k].y, tb[k].z), f3(tc[k].x, tc[k].y, tc[k].z), tmin, tmax, t, u, v)) {
                                    const int prim = __float_as_int(ta[k].w);
                                    found = true;
                                    if (hit_better(t, inst, prim, mine)) { mine.t = t; mine.u = u; mine.v = v; mine.inst = inst; mine.prim = prim; }
                                }
                            }
                        }
                    }
                }
            }
            if (__any_sync(full, bad)) { fallback = true; break; }
            if (ANY) {
                if (__any_sync(full, found)) { any_found = true; break; }
            } else {
                // every lane learns the closest hit distance so far (t > 0: the bit pattern orders like the value)
                const unsigned int tb = __reduce_min_sync(full, mine.inst >= 0 ? __float_as_uint(mine.t) : 0x7f800000u);
                limit = fminf(limit, __uint_as_float(tb));
            }
            const unsigned int m0 = __ballot_sync(full, p0), m1 = __ballot_sync(full, p1);
            const int c0 = __popc(m0), tot = c0 + __popc(m1);
            if (n + tot > VKP_WQ_CAP) { fallback = true; break; }
            if (p0) { const int k = n + __popc(m0 & lt); Q.ref[k] = q0; Q.inst[k] = inst; Q.tn[k] = qt0; }
            if (p1) { const int k = n + c0 + __popc(m1 & lt); Q.ref[k] = q1; Q.inst[k] = qi1; Q.tn[k] = qt1; }
            n += tot;
            __syncwarp();
        }
        if (!fallback) {
            if (ANY) { best.inst = any_found ? 0 : -1; return any_found; }
            // the winner under the total order (t, instance, primitive)
            const unsigned int tbits = __reduce_min_sync(full, mine.inst >= 0 ? __float_as_uint(mine.t) : 0xffffffffu);
            if (tbits == 0xffffffffu) return false;
            const bool m_t = mine.inst >= 0 && __float_as_uint(mine.t) == tbits;
            const int imin = __reduce_min_sync(full, m_t ? mine.inst : 0x7fffffff);
            const bool m_i = m_t && mine.inst == imin;
            const int pmin = __reduce_min_sync(full, m_i ? mine.prim : 0x7fffffff);
            const unsigned int who = __ballot_sync(full, m_i && mine.prim == pmin);
            const int src = __ffs(who) - 1;
            best.t = __uint_as_float(tbits); best.inst = imin; best.prim = pmin;
            best.u = __shfl_sync(full, mine.u, src); best.v = __shfl_sync(full, mine.v, src);
            return true;
        }
    }
    __syncwarp();
    return trace_ray_general<ANY>(S, o, d, tmin, tmax, best, tcount);       // same ray in every lane: one instruction stream, broadcast loads
}
