// vkp_internal.h -- host-side declarations shared by the translation units of libvkp
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/vkp.h"
#include "vkp_device.cuh"

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    bool owned = true;          // false: a view into somebody else's allocation (the BLAS arena)
    cudaError_t ensure(size_t n)
    {
        if (n <= bytes && owned) return cudaSuccess;
        if (p && owned) cudaFree(p);
        p = nullptr; bytes = 0; owned = true;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
    void release() { if (p && owned) cudaFree(p); p = nullptr; bytes = 0; owned = true; }
    void view(void* q, size_t n) { release(); p = q; bytes = n; owned = false; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// One bounding volume hierarchy (BLAS over triangles or TLAS over instance boxes)
struct Bvh {
    uint32_t n_prims = 0;
    uint32_t n_nodes = 0;
    int leaf_max = 4;
    int height = 0;
    int ploc_rounds = 0;
    const void* gate = nullptr;   // device PlocState the consumers of a freshly built PLOC hierarchy check (NULL: nothing to check)
    int root = 0;        // index of the root among the internal nodes (0 for the Karras tree)
    uint32_t root_ref = 0x80000000u;   // nodes[0].ref of the binary layout
    DevBuf nodes;        // vkp_bvh_node[n_nodes]
    DevBuf items;        // uint32_t[n_prims]: leaf-ordered primitive ids
    DevBuf tris;         // BLAS only: float4[3*n_prims]
    // topology kept for refit (Karras radix tree)
    DevBuf child;        // int2[n_prims-1]   (>=0 internal index, <0: ~leaf position)
    DevBuf parent;       // int[n_prims-1 + n_prims]: parents of internals then of leaves
    DevBuf ibox;         // float4[2*(n_prims-1)]
    DevBuf flags;        // int[n_prims-1]
    DevBuf rank;         // uint32_t[n_prims-1]: exclusive rank among emitted internals
    DevBuf keep;         // uint32_t[n_prims-1]
    DevBuf range;        // int2[n_prims-1]
    void release()
    {
        nodes.release(); items.release(); tris.release(); child.release(); parent.release(); ibox.release();
        flags.release(); rank.release(); keep.release(); range.release();
    }
};

struct BuildScratch {
    DevBuf boxes;       // float[6*n] primitive boxes
    DevBuf keys, keys_alt, vals, vals_alt, cub_tmp, bounds, misc;
    DevBuf ploc_cid[2], ploc_box[2], ploc_cnt[2], ploc_nn, ploc_scan, ploc_icnt, ploc_parent, ploc_state, ploc_tiles;
    DevBuf t_child, t_range, t_parent, t_ibox, t_flags, t_rank, t_keep;     // topology arrays of a build whose Bvh does not keep them (BLAS)
    void* result_host = nullptr;                                             // pinned BvhBuildResult of the synchronous path
    void release()
    {
        DevBuf* all[] = {&boxes, &keys, &keys_alt, &vals, &vals_alt, &cub_tmp, &bounds, &misc, &ploc_cid[0], &ploc_cid[1], &ploc_box[0], &ploc_box[1],
                         &ploc_cnt[0], &ploc_cnt[1], &ploc_nn, &ploc_scan, &ploc_icnt, &ploc_parent, &ploc_state, &ploc_tiles,
                         &t_child, &t_range, &t_parent, &t_ibox, &t_flags, &t_rank, &t_keep};
        for (DevBuf* b : all) b->release();
        if (result_host) { cudaFreeHost(result_host); result_host = nullptr; }
    }
};

// bvh_build.cu
cudaError_t bvh_triangle_boxes(cudaStream_t s, const vkp_vertex* verts, const uint32_t* indices, uint32_t n_tris, float* boxes);
cudaError_t bvh_instance_boxes(cudaStream_t s, const DInstance* inst, uint32_t n, float* boxes);
// what the host learns from a finished build (written by its last kernel into pinned host memory)
struct BvhBuildResult { uint32_t n_nodes, root_ref; int height, ploc_left, ploc_rounds, valid; uint32_t _pad[2]; };
cudaError_t bvh_reserve(BuildScratch& scratch, uint32_t n_prims, int builder, bool temps);
cudaError_t bvh_build(cudaStream_t s, BuildScratch& scratch, uint32_t n_prims, int leaf_max, Bvh& out, int* launches, int builder /*0 LBVH (Karras), 1 PLOC*/,
                      int* host_syncs = nullptr, BvhBuildResult* async_result = nullptr, bool temps_in_scratch = false);
bool bvh_build_finish(Bvh& B, const BvhBuildResult& r);
cudaError_t bvh_refit(cudaStream_t s, BuildScratch& scratch, Bvh& bvh, int* launches);
cudaError_t bvh_selftest_sort(cudaStream_t s, const uint32_t* keys_host, uint32_t n, int end_bit, uint32_t* keys_sorted_host, uint32_t* vals_sorted_host,
                              unsigned long long* scan_host);
cudaError_t bvh_emit_triangles(cudaStream_t s, const vkp_vertex* verts, const uint32_t* indices, Bvh& bvh);

// kernels_frame.cu
struct FrameLaunch {
    cudaStream_t stream;
    DScene scene;
    const DFrame* frame_dev;   // device copy of DFrame
    DFrame frame;              // host copy (for grid sizing)
    int n_own_tiles;
};
cudaError_t launch_set_carveout(int pct);
cudaError_t launch_prefetch_l2(cudaStream_t s, const void* p, size_t bytes);
cudaError_t launch_gbuffer(const FrameLaunch& L);
cudaError_t launch_shadow(const FrameLaunch& L, unsigned int* persist_counter = nullptr, uint32_t* shadow_queue = nullptr, unsigned int* shadow_qstate = nullptr);
cudaError_t launch_denoise(const FrameLaunch& L, uint8_t* atrous_a = nullptr, uint8_t* atrous_b = nullptr);
cudaError_t launch_build_order(cudaStream_t s, uint32_t* cost, uint32_t* order, uint32_t n);
cudaError_t launch_resolve(const FrameLaunch& L, int* launches);
cudaError_t launch_chain(const FrameLaunch& L, cudaStream_t stream, cudaEvent_t after_scan = nullptr, uint32_t tag = 1u, unsigned int cont_cap = 0u);
cudaError_t launch_resolve_paths(const FrameLaunch& L);
cudaError_t launch_chain_staged(const FrameLaunch& L, cudaStream_t stream);
cudaError_t launch_chain_after_first(const FrameLaunch& L, cudaStream_t stream);
cudaError_t launch_chain_first_only(const FrameLaunch& L, cudaStream_t stream);
cudaError_t launch_chain_very_long(const FrameLaunch& L, cudaStream_t stream);
bool shadow_is_persistent(const FrameLaunch& L);
unsigned int chain_first_blocks(const FrameLaunch& L);
cudaError_t launch_post(const FrameLaunch& L);
cudaError_t launch_raytrace(const FrameLaunch& L);
cudaError_t launch_deferred(const FrameLaunch& L);
cudaError_t launch_trace_rays(cudaStream_t s, const DScene& scene, const vkp_ray* rays, uint32_t n, int any_hit, vkp_hit* hits);
cudaError_t launch_trace_stats(cudaStream_t s, const DScene& scene, const vkp_ray* rays, uint32_t n, int any_hit, uint32_t* stats);
cudaError_t launch_tiles_pack(cudaStream_t s, const void* plane, void* dst, int elem_bytes, int W, int H, int tile_w, int tile_h, int rank, int world);
cudaError_t launch_tiles_unpack(cudaStream_t s, void* plane, const void* src, int elem_bytes, int W, int H, int tile_w, int tile_h, int rank, int world);
cudaError_t launch_build_luts(cudaStream_t s, float* lut_pow22, uint8_t* lut_post);
cudaError_t launch_eval_math(cudaStream_t s, int func, const float* x, const float* y, uint32_t n, float* out);
cudaError_t launch_l2_flush(cudaStream_t s, void* buf, size_t bytes);
cudaError_t launch_sync_signal(cudaStream_t s, uint32_t* flag, uint32_t value);
cudaError_t launch_sync_wait(cudaStream_t s, const uint32_t* flags, int count /* <= 64 */, uint32_t value, uint32_t* timeout_flag);
#define VKP_STAGE_INLINE_MAX 15360u     /* four launches of 3840 bytes */
cudaError_t launch_stage_inline(cudaStream_t s, void* dst, const void* src, size_t bytes);     // bytes % 4 == 0, <= VKP_STAGE_INLINE_MAX
cudaError_t launch_stage_words(cudaStream_t s, void* dst, const void* src_or_null, size_t bytes);     // bytes % 4 == 0; src NULL = clear

// vkp_api.cu, for vkp_group.cu
struct vkp_ctx;
cudaStream_t vkp_internal_stream(vkp_ctx* ctx);
cudaEvent_t vkp_internal_pending_copy(vkp_ctx* ctx, int buf);
int vkp_internal_output_buffers(vkp_ctx* ctx, void** buffer0, void** buffer1);
