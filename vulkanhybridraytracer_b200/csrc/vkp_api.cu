// vkp_api.cu -- the extern "C" boundary of libvkp (include/vkp.h): context, device memory, streams, frame graph.
// No CPU fallback lives here: every entry point that computes launches CUDA kernels, and vkp_create fails
// loudly when no sm_100-class device is usable.
#include "vkp_internal.h"
#include "vkp_bvh.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <memory>

static_assert(sizeof(vkp_build_stats) == 80, "vkp_build_stats layout (types.py BuildStats)");

namespace {

struct MeshRec { DevBuf verts, idx; uint32_t nv = 0, ni = 0; };
struct TexRec { DevBuf texels; int w = 0, h = 0; };
struct BlasRec { int mesh; uint32_t first, count; Bvh bvh; };

std::string g_create_error;

} // namespace

struct vkp_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::string err;

    std::vector<std::unique_ptr<MeshRec>> meshes;
    std::vector<std::unique_ptr<TexRec>> textures;
    int sky = -1, irr = -1;
    std::vector<vkp_material> materials;
    std::vector<vkp_light> lights;
    std::vector<vkp_instance> instances;
    std::vector<int> inst_blas;
    std::vector<std::unique_ptr<BlasRec>> blas;
    DevBuf d_materials, d_textures;
    // Per-frame uniforms -- lights, instance records, refitted TLAS boxes -- exist twice, like the reference's FrameData[FRAME_OVERLAP]
    // (renderer.h:34,49): frame k+1's VulkanEngine::update work (uploads + TLAS refit) runs on `update_stream` into the set frame k
    // is not reading, overlapped with frame k; the render stream only waits for `ev_updated`.
    struct FrameUniforms {
        DevBuf lights, inst_trav, instances, tlas_nodes;
        unsigned long long lights_ver = 0, inst_ver = 0;       // version of the host state this set holds (0 = nothing)
        cudaEvent_t last_use = nullptr;                         // end of the last frame that read the set
        bool consumed = false;                                  // a frame has read it since it was last written
    };
    FrameUniforms uni[2];
    int ucur = 0;
    unsigned long long lights_ver = 0, inst_ver = 0;            // host-side versions: bumped by vkp_set_lights / vkp_build_tlas
    std::vector<DInstance> h_inst;                              // host mirrors of the device records of the current instances
    std::vector<DInstTrav> h_trav;
    cudaStream_t update_stream = nullptr;
    cudaEvent_t ev_updated = nullptr;
    bool update_pending = false;
    Bvh tlas;
    BuildScratch scratch, tlas_scratch;
    bool blas_built = false, tlas_built = false, tlas_topology = false, textures_dirty = true;

    vkp_camera cam{};
    vkp_params params{};
    bool have_cam = false, have_params = false;
    int rank = 0, world = 1, tile_w = 16, tile_h = 8;

    DevBuf seed;                      // per-pixel RNG seeds of the frame being rendered (k_gbuffer -> k_shadow / chain / resolve)
    DevBuf position, normal, albedo, motion, material, emissive, hit_ids, final_f32, final_srgb8, shadow, denoised[2];
    int denoised_cur = 0;
    int plane_w = 0, plane_h = 0, plane_lights = 0;
    DevBuf shadow_queue, shadow_qstate;   // many-sample shadow pass: compacted (pixel, light) lists and their counters
    DevBuf chain_vl, chain_kind;          // schedules 7 / 8: list of the very long paths' slots, and the per-slot mark k_chain skips them by
    cudaStream_t chain_stream3 = nullptr;
    cudaEvent_t ev_chain3 = nullptr;
    DevBuf chain_long;                // one byte per pixel: last frame's path went beyond its first bounce (chain schedule 7)
    cudaStream_t chain_stream2 = nullptr;
    cudaEvent_t ev_chain2 = nullptr, ev_shadow = nullptr;
    bool chain2_in_flight = false;
    DevBuf atrous[2];                 // ping-pong planes of the a-trous shadow filter (vkp_params.denoiser = 1)
    DevBuf counters, io_tmp, io_tmp2, flush, path_q[2], path_count, luts, chain_buf, chain_len, chain_slot, chain_pix;
    cudaStream_t chain_stream = nullptr;
    cudaEvent_t ev_gbuffer = nullptr, ev_chain = nullptr, ev_scan = nullptr;
    cudaEvent_t ev_chain_t0 = nullptr, ev_chain_t1 = nullptr;     // profiling: the chain on its own stream (vkp_counters.ms_stage[7], ms_tlas = its end in the frame)
    bool chain_timed = false;
    bool debug_chain_timing = false;  // VKP_DEBUG_CHAIN_TIMING=1: print when the long-path kernel ended (development aid)
    cudaEvent_t ev_chain_long = nullptr;
    uint32_t chain_tag = 0;           // launch counter: the "written" mark of the continuation items of chain schedule 6 (never 0)
    // longest-first CTA schedule of k_gbuffer [0] and k_shadow [1], rebuilt every frame from the previous frame's per-block cycles
    DevBuf cta_cost[2], cta_order[2];
    uint32_t order_n[2] = {0u, 0u};
    bool order_ready[2] = {false, false}, order_pending[2] = {false, false};
    cudaStream_t order_stream = nullptr;
    cudaEvent_t ev_stage_done[2] = {nullptr, nullptr}, ev_order_done[2] = {nullptr, nullptr};
    int schedule = 1;                 // 0 = hardware block order, 1 = longest first (VKP_SCHEDULE=0|1)
    int prefetch = 0;                 // VKP_PREFETCH: 1 = stream the BVH arena into L2 at the start of every frame, 2 = vertex / index buffers too
    cudaEvent_t ev_prefetch = nullptr;
    bool fuse_accumulate = true;      // VKP_FUSE_ACCUMULATE=0: accumulation frames keep their own filter pass (k_denoise_accumulate4)
    bool use_sky_codes = true;        // VKP_SKY_CODES=0: every sky pixel redoes its light arithmetic in k_shadow
    DevBuf sky_shadow;
    bool shadow_wait_scan = false;
    bool scan_in_gbuffer = false;     // schedule 7: path queueing inside k_gbuffer instead of k_chain_scan (VKP_SCAN_IN_GBUFFER=1; measured neutral)
    bool use_seed_plane = true;       // k_gbuffer evaluates TEA once per pixel for the frame's later stages (VKP_SEED_PLANE=0: every stage evaluates it)
    bool chain_in_flight = false;
    // multi-GPU: FINAL_SRGB8 is double-buffered (vkp_set_output_buffer) and may be redirected into rank 0's two buffers over NVLink
    // (vkp_set_output_redirect): frame k+1 is written to the other buffer while rank 0 still consumes frame k
    void* peer_srgb8[2] = {nullptr, nullptr};
    DevBuf final_srgb8_alt;
    int out_sel = 0;
    std::vector<void*> ipc_opened;
    // frame barrier flags (vkp_sync_*): rank 0 owns [0..63] "rank r finished frame f" and [64] "rank 0 released frame f";
    // every rank keeps [65] of its own buffer as the timeout flag of its waits
    DevBuf sync_buf;
    uint32_t* sync_flags = nullptr;       // rank 0's flag array (own memory on rank 0, IPC mapping elsewhere)
    // present with frames in flight (vkp_readback_async): D2H copies run on their own stream behind the frame that produced them
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_frame_done = nullptr, ev_copy_done[3] = {nullptr, nullptr, nullptr};   // [0],[1]: the two sRGB8 buffers, [2]: any other plane
    bool copy_pending[3] = {false, false, false};
    // per-frame uniforms (lights, instance records) go through a ring of pinned staging slabs: no host<->device rendezvous per frame
    struct StageSlot { void* host = nullptr; size_t cap = 0; cudaEvent_t done = nullptr; bool pending = false; };
    StageSlot stage_ring[8];
    unsigned stage_seq = 0;
    cudaEvent_t ev_copy_ring[4] = {nullptr, nullptr, nullptr, nullptr};   // completion of the last 4 queued copies, by sequence number
    unsigned long long copy_seq = 0;

    bool profiling = false;
    bool count_trav = false;
    bool accumulate_counters = false;     // ray counters run on across frames until vkp_get_counters reads (and clears) them
    cudaEvent_t ev[12]{};
    int ev_stage[12]{};
    int n_ev = 0;
    cudaEvent_t timer0 = nullptr, timer1 = nullptr;
    uint32_t last_launches = 0;
    int build_launches = 0;
    // acceleration-structure figures (vkp_get_build_stats)
    vkp_build_stats bstats{};
    BvhBuildResult* build_results = nullptr;    // pinned: one record per BLAS of the batch being built
    uint32_t build_results_cap = 0;
    uint32_t bstats_rebuilt = 0;                // BLAS whose queued PLOC rounds did not suffice (rebuilt with waits)
    cudaEvent_t ev_build0 = nullptr, ev_build1 = nullptr, ev_refit0 = nullptr, ev_refit1 = nullptr;
    bool refit_timed = false;
    // every BLAS's nodes + leaf triangles, compacted into ONE allocation after the builds: the L2 persistence window covers it
    DevBuf bvh_arena;
    int l2_persist = -1;          // -1 built-in default, 0 off, 1 on (vkp_set_l2_persistence / VKP_L2_PERSIST)
    size_t l2_set_aside = 0;
    int blas_leaf_max = 4; // triangles per BLAS leaf (the traversal's leaf loop holds 4; VKP_LEAF_MAX=1..4 is a measurement knob)
    int builder = 1;      // BLAS hierarchy: 0 = Karras LBVH, 1 = PLOC (default; VKP_BLAS_BUILDER=lbvh|ploc or vkp_set_builder)
};

namespace {

int fail(vkp_ctx* c, int code, const std::string& msg)
{
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}
// H2D of a small pageable block, stream-ordered, without blocking on the frames in flight (a pageable cudaMemcpyAsync would
// rendezvous with the stream).  The slab is reused 8 uploads later; its event guards that.
cudaError_t upload_staged(vkp_ctx* ctx, cudaStream_t stream, void* dst, const void* src, size_t bytes)
{
    if (!bytes) return cudaSuccess;
    if (bytes % 4 == 0 && bytes <= VKP_STAGE_INLINE_MAX) return launch_stage_inline(stream, dst, src, bytes);     // rides in the kernel parameters
    vkp_ctx::StageSlot& sl = ctx->stage_ring[ctx->stage_seq++ & 7u];
    cudaError_t e;
    if (!sl.done && (e = cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming)) != cudaSuccess) return e;
    if (sl.pending) { if ((e = cudaEventSynchronize(sl.done)) != cudaSuccess) return e; sl.pending = false; }
    if (sl.cap < bytes) {
        if (sl.host) cudaFreeHost(sl.host);
        sl.host = nullptr; sl.cap = 0;
        size_t cap = 4096; while (cap < bytes) cap <<= 1;
        if ((e = cudaMallocHost(&sl.host, cap)) != cudaSuccess) return e;
        sl.cap = cap;
    }
    memcpy(sl.host, src, bytes);
    if (bytes % 4 == 0) e = launch_stage_words(stream, dst, sl.host, bytes);       // pinned memory is device-visible under UVA; no copy engine involved
    else e = cudaMemcpyAsync(dst, sl.host, bytes, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    if ((e = cudaEventRecord(sl.done, stream)) != cudaSuccess) return e;
    sl.pending = true;
    return cudaSuccess;
}

int cuda_fail(vkp_ctx* c, cudaError_t e, const char* what)
{
    return fail(c, e == cudaErrorMemoryAllocation ? VKP_ERR_OOM : VKP_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call)                                                   \
    do {                                                           \
        cudaError_t _e = (call);                                   \
        if (_e != cudaSuccess) return cuda_fail(ctx, _e, #call);   \
    } while (0)

size_t plane_elem_bytes(int plane)
{
    switch (plane) {
    case VKP_PLANE_POSITION: case VKP_PLANE_NORMAL: case VKP_PLANE_HIT_IDS: return 8;
    case VKP_PLANE_ALBEDO: case VKP_PLANE_MOTION: case VKP_PLANE_MATERIAL: case VKP_PLANE_EMISSIVE: case VKP_PLANE_FINAL_SRGB8: return 4;
    case VKP_PLANE_FINAL_F32: return 16;
    default: break;
    }
    if (plane >= VKP_PLANE_SHADOW0 && plane < VKP_PLANE_SHADOW0 + VKP_MAX_LIGHTS) return 1;
    if (plane >= VKP_PLANE_DENOISED0 && plane < VKP_PLANE_DENOISED0 + VKP_MAX_LIGHTS) return 1;
    return 0;
}

void* plane_ptr(vkp_ctx* c, int plane)
{
    const size_t px = (size_t)c->plane_w * c->plane_h;
    switch (plane) {
    case VKP_PLANE_POSITION: return c->position.p;
    case VKP_PLANE_NORMAL: return c->normal.p;
    case VKP_PLANE_ALBEDO: return c->albedo.p;
    case VKP_PLANE_MOTION: return c->motion.p;
    case VKP_PLANE_MATERIAL: return c->material.p;
    case VKP_PLANE_EMISSIVE: return c->emissive.p;
    case VKP_PLANE_HIT_IDS: return c->hit_ids.p;
    case VKP_PLANE_FINAL_F32: return c->final_f32.p;
    case VKP_PLANE_FINAL_SRGB8: return c->out_sel ? c->final_srgb8_alt.p : c->final_srgb8.p;
    default: break;
    }
    if (plane >= VKP_PLANE_SHADOW0 && plane < VKP_PLANE_SHADOW0 + c->plane_lights) return c->shadow.as<uint8_t>() + px * (plane - VKP_PLANE_SHADOW0);
    if (plane >= VKP_PLANE_DENOISED0 && plane < VKP_PLANE_DENOISED0 + c->plane_lights)
        return c->denoised[c->denoised_cur].as<uint8_t>() + px * (plane - VKP_PLANE_DENOISED0);
    return nullptr;
}

int ensure_planes(vkp_ctx* ctx)
{
    if (!ctx->have_params) return fail(ctx, VKP_ERR_INVALID, "vkp_set_params has not been called");
    const int W = ctx->params.width, H = ctx->params.height, L = (int)ctx->lights.size() > 0 ? (int)ctx->lights.size() : 1;
    if (W == ctx->plane_w && H == ctx->plane_h && L == ctx->plane_lights) return VKP_OK;
    const size_t px = (size_t)W * H;
    CK(ctx->position.ensure(px * 8)); CK(ctx->normal.ensure(px * 8)); CK(ctx->albedo.ensure(px * 4)); CK(ctx->motion.ensure(px * 4));
    CK(ctx->material.ensure(px * 4)); CK(ctx->emissive.ensure(px * 4)); CK(ctx->hit_ids.ensure(px * 8)); CK(ctx->seed.ensure(px * 4));
    CK(ctx->final_f32.ensure(px * 16)); CK(ctx->final_srgb8.ensure(px * 4)); CK(ctx->final_srgb8_alt.ensure(px * 4));
    CK(ctx->shadow.ensure(px * L)); CK(ctx->denoised[0].ensure(px * L)); CK(ctx->denoised[1].ensure(px * L));
    CK(ctx->counters.ensure(64));
    CK(ctx->path_q[0].ensure(px * sizeof(PathItem))); CK(ctx->path_q[1].ensure(px * sizeof(PathItem)));
    CK(ctx->path_count.ensure(sizeof(unsigned int) * (VKP_MAX_BOUNCES + 2)));
    CK(ctx->chain_buf.ensure(px * sizeof(float4) * VKP_CHAIN_MAX)); CK(ctx->chain_len.ensure(px * sizeof(int))); CK(ctx->chain_slot.ensure(px * sizeof(int))); CK(ctx->chain_pix.ensure(px * sizeof(int)));
    CK(ctx->chain_long.ensure(px)); CK(cudaMemsetAsync(ctx->chain_long.p, 0, px, ctx->stream));
    CK(ctx->chain_vl.ensure(px * sizeof(int))); CK(ctx->chain_kind.ensure(px)); CK(cudaMemsetAsync(ctx->chain_kind.p, 0, px, ctx->stream));
    ctx->plane_w = W; ctx->plane_h = H; ctx->plane_lights = L;
    // defined initial state: history and accumulation start at zero (blur.comp:73 reads uninitialised memory upstream)
    CK(cudaMemsetAsync(ctx->position.p, 0, px * 8, ctx->stream)); CK(cudaMemsetAsync(ctx->normal.p, 0, px * 8, ctx->stream));
    CK(cudaMemsetAsync(ctx->albedo.p, 0, px * 4, ctx->stream)); CK(cudaMemsetAsync(ctx->motion.p, 0, px * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->material.p, 0, px * 4, ctx->stream)); CK(cudaMemsetAsync(ctx->emissive.p, 0, px * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->hit_ids.p, 0xff, px * 8, ctx->stream));
    CK(cudaMemsetAsync(ctx->final_f32.p, 0, px * 16, ctx->stream)); CK(cudaMemsetAsync(ctx->final_srgb8.p, 0, px * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->final_srgb8_alt.p, 0, px * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->shadow.p, 0, px * L, ctx->stream));
    CK(cudaMemsetAsync(ctx->denoised[0].p, 0, px * L, ctx->stream)); CK(cudaMemsetAsync(ctx->denoised[1].p, 0, px * L, ctx->stream));
    CK(cudaMemsetAsync(ctx->counters.p, 0, 64, ctx->stream));
    ctx->denoised_cur = 0;
    return VKP_OK;
}

int upload_textures(vkp_ctx* ctx)
{
    if (!ctx->textures_dirty) return VKP_OK;
    std::vector<DTexture> t(ctx->textures.size() ? ctx->textures.size() : 1);
    for (size_t i = 0; i < ctx->textures.size(); i++) { t[i].texels = ctx->textures[i]->texels.as<uchar4>(); t[i].w = ctx->textures[i]->w; t[i].h = ctx->textures[i]->h; }
    CK(ctx->d_textures.ensure(sizeof(DTexture) * t.size()));
    CK(cudaMemcpyAsync(ctx->d_textures.p, t.data(), sizeof(DTexture) * t.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->textures_dirty = false;
    return VKP_OK;
}

// Brings the uniform set the next frame will read up to date with the host state, on the update stream.  If the current set has
// been read by a frame since it was written, the other set is taken (after the last frame that read *it*), so the frame in flight
// is never disturbed and never waited for on the host.
int flush_uniforms(vkp_ctx* ctx, bool lights_only = false)
{
    vkp_ctx::FrameUniforms* U = &ctx->uni[ctx->ucur];
    const bool want_tlas = ctx->tlas_built;
    if (U->lights_ver == ctx->lights_ver && (!want_tlas || U->inst_ver == ctx->inst_ver)) return VKP_OK;
    if (!ctx->update_stream) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        CK(cudaStreamCreateWithPriority(&ctx->update_stream, cudaStreamNonBlocking, hi));      // its few small CTAs go first
        CK(cudaEventCreateWithFlags(&ctx->ev_updated, cudaEventDisableTiming));
        for (auto& u : ctx->uni) CK(cudaEventCreateWithFlags(&u.last_use, cudaEventDisableTiming));
    }
    cudaStream_t us = ctx->update_stream;
    if (U->consumed) { ctx->ucur ^= 1; U = &ctx->uni[ctx->ucur]; }
    if (U->consumed) { CK(cudaStreamWaitEvent(us, U->last_use, 0)); U->consumed = false; }
    if (U->lights_ver != ctx->lights_ver) {
        const size_t n = ctx->lights.size();
        CK(U->lights.ensure(sizeof(vkp_light) * (n ? n : 1)));
        if (n) CK(upload_staged(ctx, us, U->lights.p, ctx->lights.data(), sizeof(vkp_light) * n));
        U->lights_ver = ctx->lights_ver;
    }
    if (want_tlas && !lights_only && U->inst_ver != ctx->inst_ver) {     // (lights_only: vkp_build_tlas or the next frame finishes the set)
        const uint32_t n = (uint32_t)ctx->instances.size();
        CK(U->inst_trav.ensure(sizeof(DInstTrav) * ctx->h_trav.size()));
        CK(U->instances.ensure(sizeof(DInstance) * ctx->h_inst.size()));
        CK(U->tlas_nodes.ensure(ctx->tlas.nodes.bytes ? ctx->tlas.nodes.bytes : sizeof(vkp_bvh_node)));
        CK(upload_staged(ctx, us, U->inst_trav.p, ctx->h_trav.data(), sizeof(DInstTrav) * ctx->h_trav.size()));
        CK(upload_staged(ctx, us, U->instances.p, ctx->h_inst.data(), sizeof(DInstance) * ctx->h_inst.size()));
        CK(ctx->tlas_scratch.boxes.ensure(sizeof(float) * 6 * (size_t)(n ? n : 1)));
        if (!ctx->ev_refit0) { CK(cudaEventCreate(&ctx->ev_refit0)); CK(cudaEventCreate(&ctx->ev_refit1)); }
        CK(cudaEventRecord(ctx->ev_refit0, us));
        CK(bvh_instance_boxes(us, U->instances.as<DInstance>(), n, ctx->tlas_scratch.boxes.as<float>()));
        // refit on the topology of the last full build, boxes written into this set's node arrays
        std::swap(ctx->tlas.nodes, U->tlas_nodes);
        int launches = 1;      // k_instance_boxes
        cudaError_t e = bvh_refit(us, ctx->tlas_scratch, ctx->tlas, &launches);
        std::swap(ctx->tlas.nodes, U->tlas_nodes);
        CK(e);
        CK(cudaEventRecord(ctx->ev_refit1, us));
        ctx->refit_timed = true; ctx->bstats.refit_launches = (uint32_t)launches;
        U->inst_ver = ctx->inst_ver;
    }
    CK(cudaEventRecord(ctx->ev_updated, us));
    ctx->update_pending = true;
    return VKP_OK;
}

// geometry + light pointers of the current uniform set; the caller's stream work is ordered behind the pending update
int scene_uniforms(vkp_ctx* ctx, DScene& S)
{
    int r = flush_uniforms(ctx);
    if (r != VKP_OK) return r;
    if (ctx->update_pending) { CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_updated, 0)); ctx->update_pending = false; }
    vkp_ctx::FrameUniforms& U = ctx->uni[ctx->ucur];
    S.tlas_nodes2 = U.tlas_nodes.as<float4>(); S.tlas_root_ref2 = ctx->tlas.root_ref;
    S.instances = U.instances.as<DInstance>();
    S.inst_trav = U.inst_trav.as<DInstTrav>();
    S.lights = U.lights.as<vkp_light>();
    S.n_instances = (int)ctx->instances.size();
    S.n_lights = (int)ctx->lights.size();
    U.consumed = true;
    return VKP_OK;
}

// after the last launch that reads the scene: the set may be rewritten once this point of the stream has passed
int mark_scene_use(vkp_ctx* ctx)
{
    vkp_ctx::FrameUniforms& U = ctx->uni[ctx->ucur];
    if (U.last_use) CK(cudaEventRecord(U.last_use, ctx->stream));
    return VKP_OK;
}

int make_scene(vkp_ctx* ctx, DScene& S)
{
    if (!ctx->tlas_built) return fail(ctx, VKP_ERR_INVALID, "acceleration structures are not built (vkp_build_blas / vkp_build_tlas)");
    if (ctx->sky < 0 || ctx->irr < 0) return fail(ctx, VKP_ERR_INVALID, "vkp_set_env has not been called");
    if (ctx->materials.empty()) return fail(ctx, VKP_ERR_INVALID, "vkp_set_materials has not been called");
    int r = upload_textures(ctx);
    if (r != VKP_OK) return r;
    if ((r = scene_uniforms(ctx, S)) != VKP_OK) return r;
    S.lut_pow22 = ctx->luts.as<float>();
    S.materials = ctx->d_materials.as<vkp_material>();
    S.textures = ctx->d_textures.as<DTexture>();
    S.n_instances = (int)ctx->instances.size(); S.n_materials = (int)ctx->materials.size();
    S.n_lights = (int)ctx->lights.size(); S.n_textures = (int)ctx->textures.size();
    S.sky_tex = ctx->sky; S.irr_tex = ctx->irr;
    return VKP_OK;
}

int make_launch(vkp_ctx* ctx, FrameLaunch& L)
{
    int r = ensure_planes(ctx);
    if (r != VKP_OK) return r;
    if ((r = make_scene(ctx, L.scene)) != VKP_OK) return r;
    if (!ctx->have_cam) return fail(ctx, VKP_ERR_INVALID, "vkp_set_camera has not been called");
    DFrame& F = L.frame;
    F.cam = ctx->cam; F.p = ctx->params;
    F.tile_w = ctx->tile_w; F.tile_h = ctx->tile_h;
    F.tiles_x = (F.p.width + F.tile_w - 1) / F.tile_w; F.tiles_y = (F.p.height + F.tile_h - 1) / F.tile_h;
    F.rank = ctx->rank; F.world = ctx->world;
    F.fused = 0; F.fuse_accumulate = 0;
    F.count_trav = ctx->count_trav ? 1 : 0;
    F.static_camera = memcmp(F.cam.view, F.cam.prev_view, 64) == 0 && memcmp(F.cam.proj, F.cam.prev_proj, 64) == 0;
    F.position = ctx->position.as<uint2>(); F.normal = ctx->normal.as<uint2>(); F.albedo = ctx->albedo.as<uchar4>();
    F.motion = ctx->motion.as<uint32_t>(); F.material = ctx->material.as<uchar4>(); F.emissive = ctx->emissive.as<uchar4>();
    F.hit_ids = ctx->hit_ids.as<int2>(); F.final_f32 = ctx->final_f32.as<float4>();
    F.sky_shadow_w = nullptr; F.sky_shadow_r = nullptr;
    F.seed_w = nullptr; F.seed_r = nullptr;       // vkp_render (Hybrid) switches the seed plane on for the frame it renders
    F.final_srgb8 = ctx->peer_srgb8[ctx->out_sel] ? reinterpret_cast<uchar4*>(ctx->peer_srgb8[ctx->out_sel])
                                                  : (ctx->out_sel ? ctx->final_srgb8_alt.as<uchar4>() : ctx->final_srgb8.as<uchar4>());
    F.shadow = ctx->shadow.as<uint8_t>();
    F.denoised = ctx->denoised[ctx->denoised_cur].as<uint8_t>();
    F.denoised_next = ctx->denoised[ctx->denoised_cur ^ 1].as<uint8_t>();
    F.counters = ctx->counters.as<unsigned long long>();
    F.path_q[0] = ctx->path_q[0].as<PathItem>(); F.path_q[1] = ctx->path_q[1].as<PathItem>();
    F.path_count = ctx->path_count.as<unsigned int>();
    F.lut_pow22 = ctx->luts.as<float>(); F.lut_post = ctx->luts.as<uint8_t>() + 1024;
    F.chain_buf = ctx->chain_buf.as<float4>(); F.chain_len = ctx->chain_len.as<int>(); F.chain_slot = ctx->chain_slot.as<int>(); F.chain_pix = ctx->chain_pix.as<int>();
    F.chain_base = F.p.wavefront == 6 ? F.path_q[0] : F.path_q[1];       // path_q[1] is schedule 6's continuation queue, path_q[0] holds the staged kernels' records
    F.chain_long = ctx->chain_long.as<uint8_t>();
    F.chain_split = F.p.wavefront >= 7 ? 1 : 0;
    {
        static int vl = -1, quad_all = -1;
        // -1 = by default only on a tile-sharded frame: there the chain's end is the frame's floor and the quads are worth 5 % (8 GPUs);
        // on one GPU they end the chain earlier and lengthen the shadow stage by as much
        if (quad_all < 0) { const char* e = getenv("VKP_CHAIN_VERY_LONG"); vl = e ? (atoi(e) != 0) : -1; const char* q = getenv("VKP_CHAIN_LONG_QUAD"); quad_all = q ? (atoi(q) != 0) : 0; }
        const bool use_vl = vl >= 0 ? vl != 0 : ctx->world > 1;
        F.chain_vl_list = (F.chain_split && use_vl && !quad_all) ? ctx->chain_vl.as<int>() : nullptr;
        F.chain_kind = ctx->chain_kind.as<uint8_t>();
        F.chain_quad_all = quad_all;
    }
    F.chain_in_gbuffer = 0; F.first_blocks = 0u; F.mid_at = 0u; F.chain_tag = 1u;      // schedule 8 switches them on inside vkp_render
    F.chain_cap = (unsigned int)((size_t)F.p.width * F.p.height);
    const int total = F.tiles_x * F.tiles_y;
    L.n_own_tiles = total > F.rank ? (total - F.rank + F.world - 1) / F.world : 0;
    F.n_cta = L.n_own_tiles * (F.tile_w >> 4) * (F.tile_h >> 3);
    {
        // magic numbers of fast_div: q = (n * (floor(2^40 / d) + 1)) >> 40 equals n / d whenever n * d < 2^40
        auto magic = [](unsigned long long d) { return ((1ull << 40) / (d ? d : 1ull)) + 1ull; };
        const unsigned long long bpt_x = (unsigned long long)(F.tile_w >> 4), bpt = bpt_x * (unsigned long long)(F.tile_h >> 3);
        const unsigned long long n_lights = ctx->lights.empty() ? 1ull : (unsigned long long)ctx->lights.size();
        const unsigned long long max_block = (unsigned long long)(F.n_cta > 0 ? F.n_cta : 1) * n_lights, max_tile = (unsigned long long)total + (unsigned long long)F.world;
        F.magic_bpt = magic(bpt); F.magic_tiles_x = magic((unsigned long long)F.tiles_x); F.magic_bpt_x = magic(bpt_x);
        F.magic_n_cta = magic((unsigned long long)F.n_cta);
        // exact while n * d < 2^40; the 64-bit product n * magic must not wrap either: n < 2^24 * d
        auto ok = [](unsigned long long n, unsigned long long d) { return n * d < (1ull << 40) && n < (d << 24); };
        F.div_exact = (ok(max_block, bpt) && ok(max_tile, (unsigned long long)F.tiles_x) && ok(bpt, bpt_x) &&
                       ok(max_block, (unsigned long long)(F.n_cta > 0 ? F.n_cta : 1))) ? 1 : 0;
    }
    for (int k = 0; k < 2; k++) {
        const uint32_t n = (uint32_t)F.n_cta * (k ? (uint32_t)L.scene.n_lights : 1u);
        F.cta_cost[k] = nullptr; F.cta_order[k] = nullptr;
        if (!ctx->schedule || !ctx->order_stream || n == 0) continue;
        if (n != ctx->order_n[k]) {       // another grid: start over with the hardware order for one frame
            if (ctx->order_pending[k]) { CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_order_done[k], 0)); ctx->order_pending[k] = false; }
            CK(ctx->cta_cost[k].ensure(sizeof(uint32_t) * n)); CK(ctx->cta_order[k].ensure(sizeof(uint32_t) * n));
            CK(launch_stage_words(ctx->stream, ctx->cta_cost[k].p, nullptr, sizeof(uint32_t) * n));
            ctx->order_n[k] = n; ctx->order_ready[k] = false;
        }
        F.cta_cost[k] = ctx->cta_cost[k].as<uint32_t>();
        if (ctx->order_ready[k]) F.cta_order[k] = ctx->cta_order[k].as<uint32_t>();
    }
    L.stream = ctx->stream;
    L.frame_dev = nullptr;
    return VKP_OK;
}

#ifndef VKP_L2_PERSIST_DEFAULT
#define VKP_L2_PERSIST_DEFAULT 1
#endif

// (re)applies the access-policy window over the BLAS arena to the streams that trace rays
int apply_l2_window(vkp_ctx* ctx)
{
    const bool on = (ctx->l2_persist < 0 ? VKP_L2_PERSIST_DEFAULT : ctx->l2_persist) != 0 && ctx->bvh_arena.p && ctx->bvh_arena.bytes;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, ctx->device));
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    ctx->bstats.l2_persist_bytes = 0; ctx->bstats.l2_hit_ratio = 0.0f;
    if (on && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
        size_t window = ctx->bvh_arena.bytes < (size_t)prop.accessPolicyMaxWindowSize ? ctx->bvh_arena.bytes : (size_t)prop.accessPolicyMaxWindowSize;
        size_t aside = window < (size_t)prop.persistingL2CacheMaxSize ? window : (size_t)prop.persistingL2CacheMaxSize;
        if (aside != ctx->l2_set_aside) { CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, aside)); ctx->l2_set_aside = aside; }
        attr.accessPolicyWindow.base_ptr = ctx->bvh_arena.p;
        attr.accessPolicyWindow.num_bytes = window;
        attr.accessPolicyWindow.hitRatio = (float)((double)aside / (double)window);
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        ctx->bstats.l2_persist_bytes = aside; ctx->bstats.l2_hit_ratio = attr.accessPolicyWindow.hitRatio;
    } else {
        attr.accessPolicyWindow.num_bytes = 0;      // disables the window
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal; attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    }
    cudaStream_t streams[] = {ctx->stream, ctx->own_stream, ctx->chain_stream, ctx->chain_stream2, ctx->chain_stream3};
    for (cudaStream_t st : streams) if (st) CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr));
    if (!on && ctx->l2_set_aside) { cudaCtxResetPersistingL2Cache(); }
    return VKP_OK;
}

// nodes + leaf triangles of every BLAS into one allocation (256-byte aligned pieces); the build's own arrays are released
int compact_blas_arena(vkp_ctx* ctx, float* copy_ms = nullptr)
{
    auto al = [](size_t n) { return (n + 255u) & ~(size_t)255u; };
    size_t total = 0;
    for (auto& b : ctx->blas) total += al(sizeof(vkp_bvh_node) * (size_t)b->bvh.n_nodes) + al(sizeof(float4) * 3 * (size_t)(b->bvh.n_prims ? b->bvh.n_prims : 1));
    DevBuf arena;
    CK(arena.ensure(total ? total : 256));
    size_t off = 0;
    std::vector<void*> placed;
    cudaEvent_t ev_c0 = nullptr, ev_c1 = nullptr;
    if (copy_ms) { CK(cudaEventCreate(&ev_c0)); CK(cudaEventCreate(&ev_c1)); CK(cudaEventRecord(ev_c0, ctx->stream)); }
    for (auto& b : ctx->blas) {
        const size_t nb = sizeof(vkp_bvh_node) * (size_t)b->bvh.n_nodes, tb = sizeof(float4) * 3 * (size_t)(b->bvh.n_prims ? b->bvh.n_prims : 1);
        char* base = static_cast<char*>(arena.p);
        CK(cudaMemcpyAsync(base + off, b->bvh.nodes.p, nb, cudaMemcpyDeviceToDevice, ctx->stream));
        placed.push_back(base + off); off += al(nb);
        CK(cudaMemcpyAsync(base + off, b->bvh.tris.p, tb, cudaMemcpyDeviceToDevice, ctx->stream));
        placed.push_back(base + off); off += al(tb);
    }
    if (copy_ms) CK(cudaEventRecord(ev_c1, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));           // one wait for all copies: the sources are freed (or are views of the old arena, freed below)
    if (copy_ms) { CK(cudaEventElapsedTime(copy_ms, ev_c0, ev_c1)); cudaEventDestroy(ev_c0); cudaEventDestroy(ev_c1); }
    for (size_t k = 0; k < ctx->blas.size(); k++) {
        Bvh& v = ctx->blas[k]->bvh;
        v.nodes.view(placed[2 * k], sizeof(vkp_bvh_node) * (size_t)v.n_nodes);
        v.tris.view(placed[2 * k + 1], sizeof(float4) * 3 * (size_t)(v.n_prims ? v.n_prims : 1));
    }
    ctx->bvh_arena.release();
    ctx->bvh_arena = arena;
    ctx->bstats.arena_bytes = total;
    return apply_l2_window(ctx);
}

int stage_event(vkp_ctx* ctx, int stage)
{
    if (!ctx->profiling || ctx->n_ev >= 12) return VKP_OK;
    if (!ctx->ev[ctx->n_ev]) CK(cudaEventCreate(&ctx->ev[ctx->n_ev]));
    CK(cudaEventRecord(ctx->ev[ctx->n_ev], ctx->stream));
    ctx->ev_stage[ctx->n_ev] = stage;
    ctx->n_ev++;
    return VKP_OK;
}

// continuation items the path queue of chain schedule 6 can hold (path_q[1]: one PathItem per pixel)
unsigned int chain_cont_cap(const vkp_ctx* ctx) { return (unsigned int)((size_t)ctx->plane_w * ctx->plane_h); }

// kernels launch_chain() queues: counter clear + k_chain_scan + k_chain, or the four kernels of the staged chain (mode 4)
uint32_t chain_launches(const FrameLaunch& L) { return L.frame.p.wavefront == 4 ? (L.scene.n_lights > 0 ? 6u : 5u) : 3u; }

int run_one(vkp_ctx* ctx, FrameLaunch& L, int stage)
{
    int r = stage_event(ctx, stage);
    if (r != VKP_OK) return r;
    if (L.n_own_tiles == 0) return VKP_OK;
    switch (stage) {
    case VKP_STAGE_GBUFFER:
    case VKP_STAGE_SHADOW: {
        const int k = stage == VKP_STAGE_SHADOW ? 1 : 0;
        if (ctx->order_pending[k]) { CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_order_done[k], 0)); ctx->order_pending[k] = false; }
        if (k) {
            // (A/B knob) VKP_SHADOW_WAIT_SCAN=1: k_shadow becomes eligible together with the chain's kernels instead of ~15 us before them,
            // so that the high-priority chain CTAs are placed first even when k_shadow is a persistent grid that would fill every SM
            if (ctx->shadow_wait_scan && ctx->chain_in_flight) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_scan, 0));
            {
                // many-sample frames: the per-light lists of (pixel, light) pairs that trace rays (k_shadow_classify / k_shadow_items); 4 bytes per
                // pixel and light, only allocated when such a frame is rendered, and only up to 2 GB (beyond that: the uncompacted kernel)
                uint32_t* sq = nullptr; unsigned int* sqs = nullptr;
                const size_t need = (size_t)ctx->plane_w * ctx->plane_h * (size_t)L.scene.n_lights * sizeof(uint32_t);
                if (L.frame.p.shadow_samples >= 4 && need <= ((size_t)2 << 30) && L.scene.n_lights <= 32) {
                    CK(ctx->shadow_queue.ensure(need)); CK(ctx->shadow_qstate.ensure(64 * sizeof(unsigned int)));
                    sq = ctx->shadow_queue.as<uint32_t>(); sqs = ctx->shadow_qstate.as<unsigned int>();
                }
                CK(launch_shadow(L, ctx->path_count.as<unsigned int>() + 8, sq, sqs));
            }      // [8]: the work counter of the persistent variant (VKP_SHADOW_PERSIST)
        } else CK(launch_gbuffer(L));
        if (L.frame.cta_cost[k]) {       // next frame's order from this frame's costs, beside the rest of the frame
            CK(cudaEventRecord(ctx->ev_stage_done[k], ctx->stream));
            CK(cudaStreamWaitEvent(ctx->order_stream, ctx->ev_stage_done[k], 0));
            CK(launch_build_order(ctx->order_stream, ctx->cta_cost[k].as<uint32_t>(), ctx->cta_order[k].as<uint32_t>(), ctx->order_n[k]));
            CK(cudaEventRecord(ctx->ev_order_done[k], ctx->order_stream));
            ctx->order_pending[k] = true; ctx->order_ready[k] = true;
            ctx->last_launches++;
        }
        break;
    }
    case VKP_STAGE_DENOISE:
        if (L.frame.fuse_accumulate) {
            // k_shadow has already written the new history (store_shadow): only the ping-pong flips
            ctx->last_launches--;          // (run_one counts one launch per stage below)
        } else
        if (L.frame.p.denoiser == 1) {
            // the a-trous taps reach 8 pixels across tile borders: one GPU only (a tile-sharded frame would need a halo exchange every frame)
            if (L.frame.world > 1) return fail(ctx, VKP_ERR_INVALID, "vkp_params.denoiser = 1 (a-trous) is not available on a tile-sharded frame (world > 1)");
            const size_t bytes = (size_t)ctx->plane_w * ctx->plane_h * ctx->plane_lights;
            CK(ctx->atrous[0].ensure(bytes)); CK(ctx->atrous[1].ensure(bytes));
            CK(launch_denoise(L, ctx->atrous[0].as<uint8_t>(), ctx->atrous[1].as<uint8_t>()));
            ctx->last_launches += 3;
        } else
        CK(launch_denoise(L));
        ctx->denoised_cur ^= 1;
        L.frame.denoised = ctx->denoised[ctx->denoised_cur].as<uint8_t>();
        L.frame.denoised_next = ctx->denoised[ctx->denoised_cur ^ 1].as<uint8_t>();
        break;
    case VKP_STAGE_RESOLVE: {
        int n = 0;
        const bool chain2 = L.frame.p.wavefront >= 2 && L.frame.p.max_bounces <= VKP_CHAIN_MAX;
        const bool late = chain2 && L.frame.p.wavefront >= 5;      // the resolve does not wait for the chain: k_resolve_paths merges it afterwards
        if (chain2 && !ctx->chain_in_flight) {
            CK(launch_chain(L, ctx->stream, nullptr, ++ctx->chain_tag, chain_cont_cap(ctx))); ctx->last_launches += chain_launches(L);
            if (L.frame.chain_vl_list) { CK(launch_chain_very_long(L, ctx->stream)); ctx->last_launches++; }
            if (L.frame.p.wavefront >= 7) { CK(launch_chain_staged(L, ctx->stream)); ctx->last_launches += L.scene.n_lights > 0 ? 4u : 3u; }
        }      // stand-alone stage: serial (clear + scan + chain)
        if (chain2 && ctx->chain_in_flight && !late) { CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_chain, 0)); ctx->chain_in_flight = false; }
        if (late && ctx->chain_in_flight) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_scan, 0));
        CK(launch_resolve(L, &n));
        if (late) {
            if (ctx->chain_in_flight) { CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_chain, 0)); ctx->chain_in_flight = false; }
            CK(launch_resolve_paths(L));
            n++;
        }
        if (L.frame.fused) {
            ctx->denoised_cur ^= 1;             // the fused filter wrote the other ping-pong buffer
            L.frame.denoised = ctx->denoised[ctx->denoised_cur].as<uint8_t>();
            L.frame.denoised_next = ctx->denoised[ctx->denoised_cur ^ 1].as<uint8_t>();
        } ctx->last_launches += (uint32_t)n - 1u;
        break;
    }
    case VKP_STAGE_POST: CK(launch_post(L)); break;
    case VKP_STAGE_RAYTRACE: CK(launch_raytrace(L)); break;
    case VKP_STAGE_DEFERRED: CK(launch_deferred(L)); break;
    default: return fail(ctx, VKP_ERR_INVALID, "unknown stage");
    }
    ctx->last_launches++;
    return VKP_OK;
}

} // namespace

extern "C" {

int vkp_abi_version(void) { return 1; }

const char* vkp_last_error(const vkp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int vkp_create(vkp_ctx** out, int device_ordinal)
{
    vkp_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, VKP_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, VKP_ERR_NO_DEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                                    " (libvkp has no CPU fallback)");
    if (device_ordinal < 0 || device_ordinal >= n) return fail(nullptr, VKP_ERR_INVALID, "device ordinal out of range");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device_ordinal)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major != 10) {
        char b[160]; snprintf(b, sizeof b, "device %d is sm_%d%d; libvkp is built for sm_100a only", device_ordinal, prop.major, prop.minor);
        return fail(nullptr, VKP_ERR_NO_DEVICE, b);
    }
    if ((e = cudaSetDevice(device_ordinal)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    ctx = new vkp_ctx();
    ctx->device = device_ordinal;
    {
        // VKP_CHAIN_PRIO=low: the frame's own stream takes the greatest priority and the chain stream the least (see below)
        const char* cp = getenv("VKP_CHAIN_PRIO");
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        e = (cp && strcmp(cp, "low") == 0) ? cudaStreamCreateWithPriority(&ctx->own_stream, cudaStreamNonBlocking, hi) : cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete ctx; return cuda_fail(nullptr, e, "cudaStreamCreate"); }
    }
    ctx->stream = ctx->own_stream;
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);               // hi = greatest priority (numerically lowest)
        // the reflection chain's stream: greatest priority by default (its CTAs start the moment slots free up);
        // VKP_CHAIN_PRIO=low lets it fill only what the main stream leaves idle (A/B knob)
        const char* cp = getenv("VKP_CHAIN_PRIO");
        cudaStreamCreateWithPriority(&ctx->chain_stream, cudaStreamNonBlocking, (cp && strcmp(cp, "low") == 0) ? lo : hi);
        cudaStreamCreateWithPriority(&ctx->order_stream, cudaStreamNonBlocking, hi);
        {
            const char* c2 = getenv("VKP_CHAIN2_PRIO");          // A/B knob: "low" = the staged kernels of schedule 7 at the main stream's priority
            cudaStreamCreateWithPriority(&ctx->chain_stream2, cudaStreamNonBlocking, (c2 && strcmp(c2, "low") == 0) ? lo : hi);
        }
        cudaEventCreateWithFlags(&ctx->ev_chain2, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_shadow, cudaEventDisableTiming);
        cudaStreamCreateWithPriority(&ctx->chain_stream3, cudaStreamNonBlocking, hi);
        cudaEventCreateWithFlags(&ctx->ev_chain3, cudaEventDisableTiming);
        for (int k = 0; k < 2; k++) { cudaEventCreateWithFlags(&ctx->ev_stage_done[k], cudaEventDisableTiming); cudaEventCreateWithFlags(&ctx->ev_order_done[k], cudaEventDisableTiming); }
        if (const char* e = getenv("VKP_SCHEDULE")) ctx->schedule = atoi(e) != 0;
        if (const char* e = getenv("VKP_SEED_PLANE")) ctx->use_seed_plane = atoi(e) != 0;
        if (const char* e = getenv("VKP_PREFETCH")) ctx->prefetch = atoi(e);
        if (const char* e = getenv("VKP_SKY_CODES")) ctx->use_sky_codes = atoi(e) != 0;
        if (const char* e = getenv("VKP_FUSE_ACCUMULATE")) ctx->fuse_accumulate = atoi(e) != 0;
        if (const char* e = getenv("VKP_SHADOW_WAIT_SCAN")) ctx->shadow_wait_scan = atoi(e) != 0;
        if (const char* e = getenv("VKP_DEBUG_CHAIN_TIMING")) ctx->debug_chain_timing = atoi(e) != 0;
        if (const char* e = getenv("VKP_SCAN_IN_GBUFFER")) ctx->scan_in_gbuffer = atoi(e) != 0;
        cudaEventCreateWithFlags(&ctx->ev_gbuffer, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_chain, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_scan, cudaEventDisableTiming);
    }
    {
        // one shared-memory / L1 split for all frame kernels (launch_set_carveout): 100 KB of shared memory covers 8 k_shadow CTAs (64 KB of
        // traversal stacks) plus the chain's beside them; VKP_SMEM_CARVEOUT=<percent> or -1 (driver default per kernel) for A/B runs
        int pct = 44;
        if (const char* c = getenv("VKP_SMEM_CARVEOUT")) pct = atoi(c);
        if ((e = launch_set_carveout(pct)) != cudaSuccess) { vkp_destroy(ctx); return cuda_fail(nullptr, e, "cudaFuncSetAttribute(carveout)"); }
    }
    if (const char* b = getenv("VKP_BLAS_BUILDER")) ctx->builder = (strcmp(b, "lbvh") == 0) ? 0 : 1;
    if (const char* e = getenv("VKP_LEAF_MAX")) { const int v = atoi(e); if (v >= 1 && v <= 4) ctx->blas_leaf_max = v; }
    if (const char* l = getenv("VKP_L2_PERSIST")) ctx->l2_persist = atoi(l) != 0;
    cudaEventCreate(&ctx->timer0); cudaEventCreate(&ctx->timer1);
    // decode / gamma tables, computed on the device with the contract's own vkp_powf so they are bit-identical to the call
    if ((e = ctx->luts.ensure(2048)) != cudaSuccess || (e = launch_build_luts(ctx->stream, ctx->luts.as<float>(), ctx->luts.as<uint8_t>() + 1024)) != cudaSuccess ||
        (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) {
        std::string msg = std::string("LUT kernel launch failed (is this an sm_100 device?): ") + cudaGetErrorString(e);
        vkp_destroy(ctx);
        return fail(nullptr, VKP_ERR_CUDA, msg);
    }
    *out = ctx;
    return VKP_OK;
}

void vkp_destroy(vkp_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->update_stream) cudaStreamSynchronize(ctx->update_stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->order_stream) cudaStreamSynchronize(ctx->order_stream);
    for (void* p : ctx->ipc_opened) cudaIpcCloseMemHandle(p);
    for (auto& m : ctx->meshes) { m->verts.release(); m->idx.release(); }
    for (auto& t : ctx->textures) t->texels.release();
    for (auto& b : ctx->blas) b->bvh.release();
    ctx->tlas.release(); ctx->scratch.release(); ctx->tlas_scratch.release();
    ctx->atrous[0].release(); ctx->atrous[1].release();
    if (ctx->l2_set_aside) cudaCtxResetPersistingL2Cache();
    ctx->bvh_arena.release();
    for (cudaEvent_t ev : {ctx->ev_build0, ctx->ev_build1, ctx->ev_refit0, ctx->ev_refit1}) if (ev) cudaEventDestroy(ev);
    for (auto& U : ctx->uni) {
        U.lights.release(); U.inst_trav.release(); U.instances.release(); U.tlas_nodes.release();
        if (U.last_use) cudaEventDestroy(U.last_use);
    }
    if (ctx->update_stream) cudaStreamDestroy(ctx->update_stream);
    if (ctx->ev_updated) cudaEventDestroy(ctx->ev_updated);
    DevBuf* bufs[] = {&ctx->d_materials, &ctx->d_textures, &ctx->position, &ctx->seed, &ctx->normal, &ctx->albedo, &ctx->motion,
                      &ctx->material, &ctx->emissive, &ctx->hit_ids, &ctx->final_f32, &ctx->final_srgb8, &ctx->shadow, &ctx->denoised[0], &ctx->denoised[1],
                      &ctx->sync_buf, &ctx->counters, &ctx->io_tmp, &ctx->io_tmp2, &ctx->flush, &ctx->path_q[0], &ctx->path_q[1], &ctx->path_count, &ctx->luts, &ctx->chain_buf, &ctx->chain_len, &ctx->chain_slot, &ctx->chain_pix, &ctx->sky_shadow, &ctx->chain_long, &ctx->chain_vl, &ctx->chain_kind, &ctx->shadow_queue, &ctx->shadow_qstate, &ctx->final_srgb8_alt};
    for (DevBuf* b : bufs) b->release();
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->timer0) cudaEventDestroy(ctx->timer0);
    if (ctx->timer1) cudaEventDestroy(ctx->timer1);
    if (ctx->chain_stream) cudaStreamDestroy(ctx->chain_stream);
    if (ctx->order_stream) cudaStreamDestroy(ctx->order_stream);
    for (int k = 0; k < 2; k++) {
        ctx->cta_cost[k].release(); ctx->cta_order[k].release();
        if (ctx->ev_stage_done[k]) cudaEventDestroy(ctx->ev_stage_done[k]);
        if (ctx->ev_order_done[k]) cudaEventDestroy(ctx->ev_order_done[k]);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->ev_frame_done) cudaEventDestroy(ctx->ev_frame_done);
    for (auto& ev : ctx->ev_copy_done) if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->ev_copy_ring) if (ev) cudaEventDestroy(ev);
    for (auto& sl : ctx->stage_ring) { if (sl.host) cudaFreeHost(sl.host); if (sl.done) cudaEventDestroy(sl.done); }
    if (ctx->build_results) cudaFreeHost(ctx->build_results);
    if (ctx->ev_gbuffer) cudaEventDestroy(ctx->ev_gbuffer);
    if (ctx->ev_chain) cudaEventDestroy(ctx->ev_chain);
    if (ctx->ev_scan) cudaEventDestroy(ctx->ev_scan);
    if (ctx->ev_prefetch) cudaEventDestroy(ctx->ev_prefetch);
    if (ctx->ev_chain2) cudaEventDestroy(ctx->ev_chain2);
    if (ctx->ev_shadow) cudaEventDestroy(ctx->ev_shadow);
    if (ctx->ev_chain_long) cudaEventDestroy(ctx->ev_chain_long);
    if (ctx->ev_chain3) cudaEventDestroy(ctx->ev_chain3);
    if (ctx->chain_stream3) cudaStreamDestroy(ctx->chain_stream3);
    if (ctx->chain_stream2) cudaStreamDestroy(ctx->chain_stream2);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int vkp_set_stream(vkp_ctx* ctx, void* cuda_stream)
{
    if (!ctx) return VKP_ERR_INVALID;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return apply_l2_window(ctx);
}

int vkp_upload_mesh(vkp_ctx* ctx, const vkp_vertex* vertices, uint32_t nv, const uint32_t* indices, uint32_t ni, int* mesh_id)
{
    if (!ctx || !vertices || !indices || !mesh_id) return fail(ctx, VKP_ERR_INVALID, "vkp_upload_mesh: NULL argument");
    if (ni % 3u) return fail(ctx, VKP_ERR_INVALID, "vkp_upload_mesh: index count is not a multiple of 3");
    for (uint32_t i = 0; i < ni; i++) if (indices[i] >= nv) return fail(ctx, VKP_ERR_INVALID, "vkp_upload_mesh: index out of range");
    CK(cudaSetDevice(ctx->device));
    std::unique_ptr<MeshRec> m(new MeshRec());
    m->nv = nv; m->ni = ni;
    CK(m->verts.ensure(sizeof(vkp_vertex) * (size_t)(nv ? nv : 1)));
    CK(m->idx.ensure(sizeof(uint32_t) * (size_t)(ni ? ni : 1)));
    CK(cudaMemcpyAsync(m->verts.p, vertices, sizeof(vkp_vertex) * (size_t)nv, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(m->idx.p, indices, sizeof(uint32_t) * (size_t)ni, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->meshes.push_back(std::move(m));
    *mesh_id = (int)ctx->meshes.size() - 1;
    return VKP_OK;
}

int vkp_upload_texture_rgba8(vkp_ctx* ctx, const uint8_t* rgba, int w, int h, int* tex_id)
{
    if (!ctx || !rgba || !tex_id || w <= 0 || h <= 0) return fail(ctx, VKP_ERR_INVALID, "vkp_upload_texture_rgba8: bad argument");
    CK(cudaSetDevice(ctx->device));
    std::unique_ptr<TexRec> t(new TexRec());
    t->w = w; t->h = h;
    CK(t->texels.ensure((size_t)w * h * 4));
    CK(cudaMemcpyAsync(t->texels.p, rgba, (size_t)w * h * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->textures.push_back(std::move(t));
    ctx->textures_dirty = true;
    *tex_id = (int)ctx->textures.size() - 1;
    return VKP_OK;
}

int vkp_set_env(vkp_ctx* ctx, int sky, int irr)
{
    if (!ctx) return VKP_ERR_INVALID;
    if (sky < 0 || irr < 0 || sky >= (int)ctx->textures.size() || irr >= (int)ctx->textures.size()) return fail(ctx, VKP_ERR_INVALID, "vkp_set_env: texture id out of range");
    ctx->sky = sky; ctx->irr = irr;
    return VKP_OK;
}

int vkp_set_materials(vkp_ctx* ctx, const vkp_material* m, int n)
{
    if (!ctx || !m || n <= 0) return fail(ctx, VKP_ERR_INVALID, "vkp_set_materials: bad argument");
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 4; k++)
            if (m[i].textures[k] > -1 && (int)m[i].textures[k] >= (int)ctx->textures.size()) return fail(ctx, VKP_ERR_INVALID, "vkp_set_materials: texture id out of range");
    CK(cudaSetDevice(ctx->device));
    ctx->materials.assign(m, m + n);
    CK(ctx->d_materials.ensure(sizeof(vkp_material) * n));
    CK(cudaMemcpyAsync(ctx->d_materials.p, ctx->materials.data(), sizeof(vkp_material) * n, cudaMemcpyHostToDevice, ctx->stream));
    return VKP_OK;
}

int vkp_set_lights(vkp_ctx* ctx, const vkp_light* l, int n)
{
    if (!ctx || (n > 0 && !l) || n < 0 || n > VKP_MAX_LIGHTS) return fail(ctx, VKP_ERR_INVALID, "vkp_set_lights: bad argument (0..32 lights)");
    CK(cudaSetDevice(ctx->device));
    ctx->lights.assign(l, l + n);
    ctx->lights_ver++;
    return flush_uniforms(ctx, true);          // on the update stream, beside the frame in flight
}

int vkp_set_instances(vkp_ctx* ctx, const vkp_instance* inst, int n)
{
    if (!ctx || (n > 0 && !inst) || n < 0) return fail(ctx, VKP_ERR_INVALID, "vkp_set_instances: bad argument");
    for (int i = 0; i < n; i++) {
        if (inst[i].mesh_id < 0 || inst[i].mesh_id >= (int)ctx->meshes.size()) return fail(ctx, VKP_ERR_INVALID, "vkp_set_instances: mesh id out of range");
        const MeshRec& m = *ctx->meshes[inst[i].mesh_id];
        if (inst[i].index_count % 3u || (uint64_t)inst[i].first_index + inst[i].index_count > m.ni) return fail(ctx, VKP_ERR_INVALID, "vkp_set_instances: index range out of the mesh");
    }
    ctx->instances.assign(inst, inst + n);
    ctx->tlas_built = false;
    return VKP_OK;
}

int vkp_set_builder(vkp_ctx* ctx, int builder)
{
    if (!ctx || builder < 0 || builder > 1) return fail(ctx, VKP_ERR_INVALID, "vkp_set_builder: 0 = LBVH, 1 = PLOC");
    ctx->builder = builder;
    return VKP_OK;
}

int vkp_build_blas(vkp_ctx* ctx)
{
    if (!ctx) return VKP_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    const auto wall0 = std::chrono::steady_clock::now();
    ctx->build_launches = 0;
    if (!ctx->ev_build0) { CK(cudaEventCreate(&ctx->ev_build0)); CK(cudaEventCreate(&ctx->ev_build1)); }
    // the geometries that have no BLAS yet (Renderer::create_bottom_acceleration_structure builds one per Node geometry, renderer.cpp:1696-1753)
    std::vector<std::unique_ptr<BlasRec>> fresh;
    uint32_t n_max = 0;
    for (const vkp_instance& I : ctx->instances) {
        bool found = false;
        for (auto& b : ctx->blas) if (b->mesh == I.mesh_id && b->first == I.first_index && b->count == I.index_count) { found = true; break; }
        for (auto& b : fresh) if (b->mesh == I.mesh_id && b->first == I.first_index && b->count == I.index_count) { found = true; break; }
        if (found) continue;
        std::unique_ptr<BlasRec> b(new BlasRec());
        b->mesh = I.mesh_id; b->first = I.first_index; b->count = I.index_count;
        n_max = I.index_count / 3u > n_max ? I.index_count / 3u : n_max;
        fresh.push_back(std::move(b));
    }
    int host_syncs = 0; const uint32_t built = (uint32_t)fresh.size(); uint64_t built_tris = 0;
    ctx->bstats_rebuilt = 0;
    if (built) {
        // Every build of the batch is queued before anything is waited for: scratch sized once for the largest geometry (no
        // cudaFree between builds), PLOC rounds driven from device memory, one pinned result record per BLAS, ONE synchronisation.
        CK(bvh_reserve(ctx->scratch, n_max, ctx->builder, true));
        if (ctx->build_results_cap < built) {
            if (ctx->build_results) cudaFreeHost(ctx->build_results);
            ctx->build_results = nullptr; ctx->build_results_cap = 0;
            CK(cudaMallocHost(&ctx->build_results, sizeof(BvhBuildResult) * built));
            ctx->build_results_cap = built;
        }
        memset(ctx->build_results, 0, sizeof(BvhBuildResult) * built);
        // the per-BLAS outputs are allocated up front as well (upper bounds: 2n - 1 nodes): no cudaMalloc between the queued builds
        for (uint32_t k = 0; k < built; k++) {
            Bvh& v = fresh[k]->bvh;
            const size_t ntri = fresh[k]->count / 3u, ni = ntri > 1 ? ntri - 1 : 0;
            CK(v.nodes.ensure(sizeof(vkp_bvh_node) * (2 * ni + 1))); CK(v.items.ensure(sizeof(uint32_t) * (ntri ? ntri : 1)));
            CK(v.tris.ensure(sizeof(float4) * 3 * (ntri ? ntri : 1)));
        }
        CK(cudaEventRecord(ctx->ev_build0, ctx->stream));
        for (uint32_t k = 0; k < built; k++) {
            BlasRec& b = *fresh[k];
            const MeshRec& m = *ctx->meshes[b.mesh];
            const uint32_t ntri = b.count / 3u;
            const uint32_t* idx = m.idx.as<uint32_t>() + b.first;
            CK(bvh_triangle_boxes(ctx->stream, m.verts.as<vkp_vertex>(), idx, ntri, ctx->scratch.boxes.as<float>()));
            CK(bvh_build(ctx->stream, ctx->scratch, ntri, ctx->blas_leaf_max, b.bvh, &ctx->build_launches, ctx->builder, &host_syncs, ctx->build_results + k, true));
            CK(bvh_emit_triangles(ctx->stream, m.verts.as<vkp_vertex>(), idx, b.bvh));
            ctx->build_launches += 2;      // k_triangle_boxes, k_emit_triangles
            built_tris += ntri;
        }
        CK(cudaEventRecord(ctx->ev_build1, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream)); host_syncs++;
        for (uint32_t k = 0; k < built; k++) {
            BlasRec& b = *fresh[k];
            if (!bvh_build_finish(b.bvh, ctx->build_results[k])) {
                // the queued PLOC rounds did not finish this hierarchy (never seen on triangle meshes): build it again, waiting per batch of rounds
                const MeshRec& m = *ctx->meshes[b.mesh];
                const uint32_t ntri = b.count / 3u;
                const uint32_t* idx = m.idx.as<uint32_t>() + b.first;
                CK(bvh_triangle_boxes(ctx->stream, m.verts.as<vkp_vertex>(), idx, ntri, ctx->scratch.boxes.as<float>()));
                CK(bvh_build(ctx->stream, ctx->scratch, ntri, ctx->blas_leaf_max, b.bvh, &ctx->build_launches, ctx->builder, &host_syncs, nullptr, true));
                CK(bvh_emit_triangles(ctx->stream, m.verts.as<vkp_vertex>(), idx, b.bvh));
                ctx->bstats_rebuilt++;
            }
            // the topology arrays were views of the scratch: a BLAS is never refitted
            b.bvh.child.release(); b.bvh.range.release(); b.bvh.parent.release(); b.bvh.ibox.release(); b.bvh.flags.release(); b.bvh.rank.release(); b.bvh.keep.release();
            if (b.bvh.height + 16 > VKP_TRAV_STACK) {     // one push per level + the TLAS part (checked exactly in vkp_build_tlas)
                char msg[160]; snprintf(msg, sizeof msg, "BLAS height %d exceeds the traversal stack budget", b.bvh.height);
                return fail(ctx, VKP_ERR_INVALID, msg);
            }
        }
        for (auto& b : fresh) ctx->blas.push_back(std::move(b));
        float ms_build = 0.0f, ms_compact = 0.0f;
        int r = compact_blas_arena(ctx, &ms_compact);      // one allocation for every BLAS + the L2 persistence window over it
        if (r != VKP_OK) return r;
        CK(cudaEventElapsedTime(&ms_build, ctx->ev_build0, ctx->ev_build1));
        ctx->tlas_built = false;              // the instance records name the old arrays: vkp_build_tlas has to follow
        ctx->bstats.blas_ms = ms_build; ctx->bstats.blas_compact_ms = ms_compact;
        ctx->bstats.blas_wall_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
        ctx->bstats.blas_built = built; ctx->bstats.blas_launches = (uint32_t)ctx->build_launches; ctx->bstats.blas_host_syncs = (uint32_t)host_syncs;
        ctx->bstats.blas_triangles = built_tris;
        int rmax = 0;
        for (auto& b : ctx->blas) rmax = b->bvh.ploc_rounds > rmax ? b->bvh.ploc_rounds : rmax;
        ctx->bstats.blas_ploc_rounds_max = (uint16_t)rmax; ctx->bstats.blas_rebuilt = (uint16_t)ctx->bstats_rebuilt;
    }
    uint32_t hmax = 0;
    for (auto& b : ctx->blas) hmax = (uint32_t)b->bvh.height > hmax ? (uint32_t)b->bvh.height : hmax;
    ctx->bstats.blas_height_max = hmax;
    ctx->blas_built = true;
    return VKP_OK;
}

int vkp_build_tlas(vkp_ctx* ctx, int update)
{
    if (!ctx) return VKP_ERR_INVALID;
    if (!ctx->blas_built) return fail(ctx, VKP_ERR_INVALID, "vkp_build_tlas: call vkp_build_blas first");
    CK(cudaSetDevice(ctx->device));
    const int n = (int)ctx->instances.size();
    std::vector<DInstance>& d = ctx->h_inst; std::vector<DInstTrav>& tv = ctx->h_trav;
    d.assign(n ? n : 1, DInstance{}); tv.assign(n ? n : 1, DInstTrav{});
    for (int i = 0; i < n; i++) {
        const vkp_instance& I = ctx->instances[i];
        BlasRec* rec = nullptr;
        for (auto& b : ctx->blas) if (b->mesh == I.mesh_id && b->first == I.first_index && b->count == I.index_count) { rec = b.get(); break; }
        if (!rec) return fail(ctx, VKP_ERR_INVALID, "vkp_build_tlas: an instance references geometry without a BLAS; call vkp_build_blas");
        const MeshRec& m = *ctx->meshes[I.mesh_id];
        memcpy(d[i].model, I.model, 64); memcpy(d[i].inv_model, I.inv_model, 64);
        d[i].nodes = rec->bvh.nodes.as<vkp_bvh_node>(); d[i].tris = rec->bvh.tris.as<float4>();
        d[i].vertices = m.verts.as<vkp_vertex>(); d[i].indices = m.idx.as<uint32_t>() + I.first_index;
        d[i].material_id = I.material_id;
        const float* im = I.inv_model;
        tv[i].r0 = make_float4(im[0], im[4], im[8], im[12]); tv[i].r1 = make_float4(im[1], im[5], im[9], im[13]); tv[i].r2 = make_float4(im[2], im[6], im[10], im[14]);
        tv[i].nodes = rec->bvh.nodes.as<float4>(); tv[i].tris = rec->bvh.tris.as<float4>(); tv[i].root_ref = rec->bvh.root_ref;
    }
    const bool refit = update && ctx->tlas_topology && ctx->tlas.n_prims == (uint32_t)n && ctx->tlas.n_nodes > 0;
    if (!refit) {
        // full build (topology): rare and synchronous.  Nothing may be in flight on either stream while the shared arrays are rebuilt.
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->update_stream) CK(cudaStreamSynchronize(ctx->update_stream));
        vkp_ctx::FrameUniforms& U = ctx->uni[ctx->ucur];
        CK(U.instances.ensure(sizeof(DInstance) * d.size()));
        CK(cudaMemcpy(U.instances.p, d.data(), sizeof(DInstance) * d.size(), cudaMemcpyHostToDevice));
        CK(ctx->tlas_scratch.boxes.ensure(sizeof(float) * 6 * (size_t)(n ? n : 1)));
        CK(bvh_instance_boxes(ctx->stream, U.instances.as<DInstance>(), (uint32_t)n, ctx->tlas_scratch.boxes.as<float>()));
        int launches = 0;
        if (!ctx->ev_build0) { CK(cudaEventCreate(&ctx->ev_build0)); CK(cudaEventCreate(&ctx->ev_build1)); }
        CK(cudaEventRecord(ctx->ev_build0, ctx->stream));
        CK(bvh_build(ctx->stream, ctx->tlas_scratch, (uint32_t)n, 1, ctx->tlas, &launches, 0));
        CK(cudaEventRecord(ctx->ev_build1, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaEventElapsedTime(&ctx->bstats.tlas_ms, ctx->ev_build0, ctx->ev_build1));
        ctx->bstats.tlas_height = (uint32_t)ctx->tlas.height;
        {
            // the per-thread traversal stack holds the sentinel, the TLAS part, the return marker and the deepest BLAS part
            int blas_h = 0;
            for (auto& b : ctx->blas) blas_h = b->bvh.height > blas_h ? b->bvh.height : blas_h;
            const int need = ctx->tlas.height + blas_h + 3;
            if (need > VKP_TRAV_STACK) {
                char msg[200]; snprintf(msg, sizeof msg, "TLAS height %d + BLAS height %d exceed the traversal stack budget (%d entries)", ctx->tlas.height, blas_h, VKP_TRAV_STACK);
                ctx->tlas_topology = false;
                return fail(ctx, VKP_ERR_INVALID, msg);
            }
        }
        ctx->tlas_topology = true;
        for (auto& u : ctx->uni) { u.consumed = false; u.inst_ver = 0; }
    }
    // both sets are now stale; the one the next frame reads is brought up to date here, overlapping whatever frame is in flight
    ctx->inst_ver++;
    ctx->tlas_built = true;
    return flush_uniforms(ctx);
}

int vkp_get_build_stats(vkp_ctx* ctx, vkp_build_stats* out)
{
    if (!ctx || !out) return VKP_ERR_INVALID;
    if (ctx->refit_timed) {
        CK(cudaEventSynchronize(ctx->ev_refit1));
        CK(cudaEventElapsedTime(&ctx->bstats.refit_ms, ctx->ev_refit0, ctx->ev_refit1));
    }
    *out = ctx->bstats;
    return VKP_OK;
}

int vkp_set_l2_persistence(vkp_ctx* ctx, int on)
{
    if (!ctx || on < -1 || on > 1) return fail(ctx, VKP_ERR_INVALID, "vkp_set_l2_persistence: -1 default, 0 off, 1 on");
    CK(cudaSetDevice(ctx->device));
    ctx->l2_persist = on;
    return apply_l2_window(ctx);
}

int vkp_set_camera(vkp_ctx* ctx, const vkp_camera* cam)
{
    if (!ctx || !cam) return fail(ctx, VKP_ERR_INVALID, "vkp_set_camera: NULL");
    ctx->cam = *cam; ctx->have_cam = true;
    return VKP_OK;
}

int vkp_set_params(vkp_ctx* ctx, const vkp_params* p)
{
    if (!ctx || !p) return fail(ctx, VKP_ERR_INVALID, "vkp_set_params: NULL");
    if (p->width <= 0 || p->height <= 0 || p->width > 16384 || p->height > 16384) return fail(ctx, VKP_ERR_INVALID, "vkp_set_params: bad image size");
    if (p->mode < 0 || p->mode > 2) return fail(ctx, VKP_ERR_INVALID, "vkp_set_params: bad mode");
    if (p->max_bounces > VKP_MAX_BOUNCES) return fail(ctx, VKP_ERR_INVALID, "vkp_set_params: max_bounces > 64");
    if (p->spp < 1 || p->shadow_samples < 1 || p->max_bounces < 0) return fail(ctx, VKP_ERR_INVALID, "vkp_set_params: spp / shadow_samples must be >= 1");
    if (p->denoiser < 0 || p->denoiser > 1) return fail(ctx, VKP_ERR_INVALID, "vkp_set_params: denoiser is 0 (blur.comp) or 1 (a-trous)");
    if (p->wavefront < 0 || p->wavefront > 8) return fail(ctx, VKP_ERR_INVALID, "vkp_set_params: wavefront (chain schedule) is 0 .. 8");
    ctx->params = *p; ctx->have_params = true;
    return VKP_OK;
}

int vkp_set_tiling(vkp_ctx* ctx, int rank, int world, int tile_w, int tile_h)
{
    if (!ctx || world < 1 || rank < 0 || rank >= world) return fail(ctx, VKP_ERR_INVALID, "vkp_set_tiling: bad rank / world");
    if (tile_w < 16 || tile_h < 8 || tile_w % 16 || tile_h % 8) return fail(ctx, VKP_ERR_INVALID, "vkp_set_tiling: tiles must be multiples of 16 x 8 pixels");
    ctx->rank = rank; ctx->world = world; ctx->tile_w = tile_w; ctx->tile_h = tile_h;
    return VKP_OK;
}

int vkp_render(vkp_ctx* ctx)
{
    if (!ctx) return VKP_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    FrameLaunch L;
    int r = make_launch(ctx, L);
    if (r != VKP_OK) return r;
    ctx->n_ev = 0; ctx->last_launches = 0;
    // a D2H copy still reading a buffer this frame overwrites (vkp_readback_async) is waited for on the device, not the host
    for (int b = 0; b < 3; b++) {
        if (!ctx->copy_pending[b] || (b < 2 && b != ctx->out_sel)) continue;
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy_done[b], 0));
        ctx->copy_pending[b] = false;
    }
    if (!ctx->accumulate_counters) CK(launch_stage_words(ctx->stream, ctx->counters.p, nullptr, 64));
    if (ctx->prefetch > 0) {
        // the BVH arena (nodes + leaf triangles of every BLAS) into L2 ahead of the traversals; level 2 adds the vertex and index
        // buffers the hit shaders read.  On a side stream: the first primary rays do not wait for it.
        if (!ctx->ev_prefetch) { CK(cudaEventCreateWithFlags(&ctx->ev_prefetch, cudaEventDisableTiming)); }
        cudaStream_t ps = ctx->order_stream ? ctx->order_stream : ctx->stream;
        CK(cudaEventRecord(ctx->ev_prefetch, ctx->stream));
        CK(cudaStreamWaitEvent(ps, ctx->ev_prefetch, 0));
        CK(launch_prefetch_l2(ps, ctx->bvh_arena.p, ctx->bvh_arena.bytes));
        ctx->last_launches++;
        if (ctx->prefetch > 1)
            for (auto& m : ctx->meshes) { CK(launch_prefetch_l2(ps, m->verts.p, m->verts.bytes)); CK(launch_prefetch_l2(ps, m->idx.p, m->idx.bytes)); ctx->last_launches += 2; }
    }
    static const int hybrid[] = {VKP_STAGE_GBUFFER, VKP_STAGE_SHADOW, VKP_STAGE_DENOISE, VKP_STAGE_RESOLVE};
    static const int hybrid_fused[] = {VKP_STAGE_GBUFFER, VKP_STAGE_SHADOW, VKP_STAGE_RESOLVE};
    static const int rt[] = {VKP_STAGE_RAYTRACE};
    static const int deferred[] = {VKP_STAGE_GBUFFER, VKP_STAGE_DEFERRED};
    const int* seq = hybrid; int n = 4;
    // the wavefront variant re-reads the resolve inputs per bounce; it keeps the stand-alone filter pass
    const bool fused = ctx->params.mode == VKP_MODE_HYBRID && ctx->params.fuse_denoise != 0 && ctx->params.wavefront != 1 && ctx->params.denoiser == 0;
    if (fused) { seq = hybrid_fused; n = 3; L.frame.fused = 1; }
    if (ctx->params.mode == VKP_MODE_RAYTRACING) { seq = rt; n = 1; }
    else if (ctx->params.mode == VKP_MODE_DEFERRED) { seq = deferred; n = 2; }
    // accumulation frames (denoise_frame > 0): blur.comp is a per-pixel blend, applied by the shadow threads themselves (VKP_FUSE_ACCUMULATE=0: own pass)
    if (ctx->params.mode == VKP_MODE_HYBRID && !fused && ctx->params.denoiser == 0 && ctx->params.denoise_frame > 0 && ctx->params.shadow_samples < 4 &&
        ctx->fuse_accumulate && !ctx->lights.empty())
        L.frame.fuse_accumulate = 1;
    const bool seeds = ctx->params.mode == VKP_MODE_HYBRID && ctx->use_seed_plane;
    if (seeds) L.frame.seed_w = ctx->seed.as<uint32_t>();
    const bool sky_codes = ctx->params.mode == VKP_MODE_HYBRID && ctx->use_sky_codes;
    if (sky_codes) { CK(ctx->sky_shadow.ensure(64 * sizeof(short))); L.frame.sky_shadow_w = ctx->sky_shadow.as<short>(); }
    const bool chained = ctx->params.mode == VKP_MODE_HYBRID && ctx->params.wavefront >= 2 && ctx->params.max_bounces <= VKP_CHAIN_MAX && L.n_own_tiles > 0;
    // schedule 8: k_gbuffer queues the path pixels itself, the first reflection rays ride in k_shadow's grid, and what depends on them
    // (per-light any-hit rays, shading, stragglers) follows the shadow pass on the third stream, beside the filter and the resolve
    const bool folded = chained && ctx->params.wavefront == 8;
    // schedules 7 and 8: k_gbuffer queues the path pixels itself (no k_chain_scan pass over the frame on the chain's critical path;
    // VKP_SCAN_IN_GBUFFER=0 keeps the separate pass for schedule 7)
    const bool scan_in_gbuffer = folded || (chained && ctx->params.wavefront == 7 && ctx->scan_in_gbuffer);
    if (scan_in_gbuffer) {
        CK(launch_stage_words(ctx->stream, ctx->path_count.p, nullptr, 7 * sizeof(unsigned int)));
        L.frame.chain_in_gbuffer = 1;
        ctx->last_launches++;
    }
    for (int i = 0; i < n; i++) {
        if (folded && seq[i] == VKP_STAGE_SHADOW) {
            if (shadow_is_persistent(L)) { CK(launch_chain_first_only(L, ctx->stream)); ctx->last_launches++; }      // (the persistent grid has no leading blocks)
            else {
                L.frame.first_blocks = chain_first_blocks(L);
                // two thirds down the shadow blocks: late enough for the leading blocks to be done, early enough for the last third to cover the any-hit tail
                L.frame.mid_at = (unsigned int)(((unsigned long long)L.frame.n_cta * (unsigned long long)L.scene.n_lights * 2ull) / 3ull);
                L.frame.chain_tag = ++ctx->chain_tag;
                if (L.frame.chain_tag == 0u) L.frame.chain_tag = ++ctx->chain_tag;
            }
        }
        if ((r = run_one(ctx, L, seq[i])) != VKP_OK) return r;
        if (seq[i] == VKP_STAGE_GBUFFER && seeds) { L.frame.seed_w = nullptr; L.frame.seed_r = ctx->seed.as<uint32_t>(); }
        if (seq[i] == VKP_STAGE_GBUFFER && sky_codes) { L.frame.sky_shadow_w = nullptr; L.frame.sky_shadow_r = ctx->sky_shadow.as<short>(); }
        if (seq[i] == VKP_STAGE_GBUFFER && chained) {
            // fork: the reflection chain needs only the G-buffer; it overlaps the shadow + denoise passes on its own stream
            CK(cudaEventRecord(ctx->ev_gbuffer, ctx->stream));
            CK(cudaStreamWaitEvent(ctx->chain_stream, ctx->ev_gbuffer, 0));
            if (ctx->profiling) {
                if (!ctx->ev_chain_t0) { CK(cudaEventCreate(&ctx->ev_chain_t0)); CK(cudaEventCreate(&ctx->ev_chain_t1)); }
                CK(cudaEventRecord(ctx->ev_chain_t0, ctx->chain_stream));
            }
            CK(launch_chain(L, ctx->chain_stream, ctx->ev_scan, ++ctx->chain_tag, chain_cont_cap(ctx)));
            ctx->last_launches += scan_in_gbuffer ? 1u : chain_launches(L);
            if (L.frame.chain_vl_list) {
                // the very long paths, a quad of lanes each, beside k_chain (which skips them)
                CK(cudaStreamWaitEvent(ctx->chain_stream3, ctx->ev_scan, 0));
                CK(launch_chain_very_long(L, ctx->chain_stream3));
                CK(cudaEventRecord(ctx->ev_chain3, ctx->chain_stream3));
                CK(cudaStreamWaitEvent(ctx->chain_stream, ctx->ev_chain3, 0));
                ctx->last_launches++;
            }
            if (ctx->profiling && ctx->debug_chain_timing) {
                if (!ctx->ev_chain_long) CK(cudaEventCreate(&ctx->ev_chain_long));
                CK(cudaEventRecord(ctx->ev_chain_long, ctx->chain_stream));          // end of the long-path kernel alone (before it waits for the staged half)
            }
            if (ctx->params.wavefront == 7) {
                // the bulk of the paths (predicted to end after one bounce) through the lean staged kernels, beside k_chain's long ones
                CK(cudaStreamWaitEvent(ctx->chain_stream2, ctx->ev_scan, 0));
                CK(launch_chain_staged(L, ctx->chain_stream2));
                CK(cudaEventRecord(ctx->ev_chain2, ctx->chain_stream2));
                CK(cudaStreamWaitEvent(ctx->chain_stream, ctx->ev_chain2, 0));       // ev_chain below then means: both halves are done
                ctx->last_launches += L.scene.n_lights > 0 ? 4u : 3u;
            }
            if (!folded) {
                if (ctx->profiling) { CK(cudaEventRecord(ctx->ev_chain_t1, ctx->chain_stream)); ctx->chain_timed = true; }
                CK(cudaEventRecord(ctx->ev_chain, ctx->chain_stream));
            }
            ctx->chain_in_flight = true;
        }
        if (seq[i] == VKP_STAGE_SHADOW && folded) {
            CK(cudaEventRecord(ctx->ev_shadow, ctx->stream));
            CK(cudaStreamWaitEvent(ctx->chain_stream2, ctx->ev_shadow, 0));
            CK(launch_chain_after_first(L, ctx->chain_stream2));       // (skips k_chain_shadow0 when its work rode in k_shadow's grid)
            const bool rode = L.frame.first_blocks != 0u;
            L.frame.first_blocks = 0u;
            CK(cudaEventRecord(ctx->ev_chain2, ctx->chain_stream2));
            CK(cudaStreamWaitEvent(ctx->chain_stream, ctx->ev_chain2, 0));
            if (ctx->profiling) { CK(cudaEventRecord(ctx->ev_chain_t1, ctx->chain_stream)); ctx->chain_timed = true; }
            CK(cudaEventRecord(ctx->ev_chain, ctx->chain_stream));                   // both halves are done
            ctx->last_launches += (L.scene.n_lights > 0 && !rode) ? 3u : 2u;
        }
    }
    if ((r = mark_scene_use(ctx)) != VKP_OK) return r;
    return stage_event(ctx, -1);
}

int vkp_run_stage(vkp_ctx* ctx, int stage)
{
    if (!ctx) return VKP_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    FrameLaunch L;
    int r = make_launch(ctx, L);
    if (r != VKP_OK) return r;
    ctx->n_ev = 0; ctx->last_launches = 0;
    if (!ctx->accumulate_counters) CK(launch_stage_words(ctx->stream, ctx->counters.p, nullptr, 64));
    if ((r = run_one(ctx, L, stage)) != VKP_OK) return r;
    if ((r = mark_scene_use(ctx)) != VKP_OK) return r;
    return stage_event(ctx, -1);
}

int vkp_sync(vkp_ctx* ctx)
{
    if (!ctx) return VKP_ERR_INVALID;
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->update_stream) CK(cudaStreamSynchronize(ctx->update_stream));
    return VKP_OK;
}

int vkp_reset_history(vkp_ctx* ctx)
{
    if (!ctx) return VKP_ERR_INVALID;
    const size_t px = (size_t)ctx->plane_w * ctx->plane_h;
    if (!px) return VKP_OK;
    CK(cudaMemsetAsync(ctx->denoised[0].p, 0, px * ctx->plane_lights, ctx->stream));
    CK(cudaMemsetAsync(ctx->denoised[1].p, 0, px * ctx->plane_lights, ctx->stream));
    CK(cudaMemsetAsync(ctx->final_f32.p, 0, px * 16, ctx->stream));
    return VKP_OK;
}

int vkp_plane_bytes(const vkp_ctx* ctx, int plane, size_t* bytes)
{
    if (!ctx || !bytes || !ctx->have_params) return VKP_ERR_INVALID;
    size_t e = plane_elem_bytes(plane);
    if (!e) return VKP_ERR_INVALID;
    *bytes = e * (size_t)ctx->params.width * ctx->params.height;
    return VKP_OK;
}

int vkp_readback(vkp_ctx* ctx, int plane, void* dst, size_t bytes)
{
    if (!ctx || !dst) return fail(ctx, VKP_ERR_INVALID, "vkp_readback: NULL");
    int r = ensure_planes(ctx);
    if (r != VKP_OK) return r;
    void* p = plane_ptr(ctx, plane);
    size_t need = plane_elem_bytes(plane) * (size_t)ctx->plane_w * ctx->plane_h;
    if (!p || !need) return fail(ctx, VKP_ERR_INVALID, "vkp_readback: unknown plane");
    if (bytes < need) return fail(ctx, VKP_ERR_INVALID, "vkp_readback: destination too small");
    CK(cudaMemcpyAsync(dst, p, need, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VKP_OK;
}

int vkp_readback_async(vkp_ctx* ctx, int plane, void* dst, size_t bytes)
{
    if (!ctx || !dst) return fail(ctx, VKP_ERR_INVALID, "vkp_readback_async: NULL");
    int r = ensure_planes(ctx);
    if (r != VKP_OK) return r;
    void* p = plane_ptr(ctx, plane);
    size_t need = plane_elem_bytes(plane) * (size_t)ctx->plane_w * ctx->plane_h;
    if (!p || !need) return fail(ctx, VKP_ERR_INVALID, "vkp_readback_async: unknown plane");
    if (bytes < need) return fail(ctx, VKP_ERR_INVALID, "vkp_readback_async: destination too small");
    if (!ctx->copy_stream) {
        CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->ev_frame_done, cudaEventDisableTiming));
        for (auto& ev : ctx->ev_copy_done) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        for (auto& ev : ctx->ev_copy_ring) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    }
    const int slot = plane == VKP_PLANE_FINAL_SRGB8 ? ctx->out_sel : 2;
    CK(cudaEventRecord(ctx->ev_frame_done, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_frame_done, 0));
    CK(cudaMemcpyAsync(dst, p, need, cudaMemcpyDeviceToHost, ctx->copy_stream));
    CK(cudaEventRecord(ctx->ev_copy_done[slot], ctx->copy_stream));
    ctx->copy_pending[slot] = true;
    CK(cudaEventRecord(ctx->ev_copy_ring[ctx->copy_seq & 3ull], ctx->copy_stream));
    ctx->copy_seq++;
    return VKP_OK;
}

int vkp_readback_wait(vkp_ctx* ctx, int keep_in_flight)
{
    if (!ctx || keep_in_flight < 0) return VKP_ERR_INVALID;
    if (!ctx->copy_stream || ctx->copy_seq <= (unsigned long long)keep_in_flight) return VKP_OK;
    if (keep_in_flight >= 4) return VKP_OK;      // the ring remembers 4 copies: nothing older can still be pending
    if (keep_in_flight == 0) { CK(cudaStreamSynchronize(ctx->copy_stream)); return VKP_OK; }
    CK(cudaEventSynchronize(ctx->ev_copy_ring[(ctx->copy_seq - 1ull - (unsigned long long)keep_in_flight) & 3ull]));
    return VKP_OK;
}

int vkp_upload_plane(vkp_ctx* ctx, int plane, const void* src, size_t bytes)
{
    if (!ctx || !src) return fail(ctx, VKP_ERR_INVALID, "vkp_upload_plane: NULL");
    int r = ensure_planes(ctx);
    if (r != VKP_OK) return r;
    void* p = plane_ptr(ctx, plane);
    size_t need = plane_elem_bytes(plane) * (size_t)ctx->plane_w * ctx->plane_h;
    if (!p || !need) return fail(ctx, VKP_ERR_INVALID, "vkp_upload_plane: unknown plane");
    if (bytes < need) return fail(ctx, VKP_ERR_INVALID, "vkp_upload_plane: source too small");
    CK(cudaMemcpyAsync(p, src, need, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VKP_OK;
}

int vkp_plane_device_ptr(vkp_ctx* ctx, int plane, void** device_ptr, size_t* bytes)
{
    if (!ctx || !device_ptr) return VKP_ERR_INVALID;
    int r = ensure_planes(ctx);
    if (r != VKP_OK) return r;
    void* p = plane_ptr(ctx, plane);
    if (!p) return fail(ctx, VKP_ERR_INVALID, "vkp_plane_device_ptr: unknown plane");
    *device_ptr = p;
    if (bytes) *bytes = plane_elem_bytes(plane) * (size_t)ctx->plane_w * ctx->plane_h;
    return VKP_OK;
}

int vkp_ipc_export_plane(vkp_ctx* ctx, int plane, void* handle64)
{
    if (!ctx || !handle64) return VKP_ERR_INVALID;
    int r = ensure_planes(ctx);
    if (r != VKP_OK) return r;
    void* p = plane_ptr(ctx, plane);
    if (!p) return fail(ctx, VKP_ERR_INVALID, "vkp_ipc_export_plane: unknown plane");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CK(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), p));
    return VKP_OK;
}

int vkp_ipc_open(vkp_ctx* ctx, const void* handle64, void** device_ptr)
{
    if (!ctx || !handle64 || !device_ptr) return VKP_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    CK(cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->ipc_opened.push_back(*device_ptr);
    return VKP_OK;
}

static int ensure_sync_buf(vkp_ctx* ctx)
{
    if (ctx->sync_buf.p) return VKP_OK;
    CK(ctx->sync_buf.ensure(1024));
    CK(cudaMemset(ctx->sync_buf.p, 0, 1024));
    return VKP_OK;
}

int vkp_sync_export(vkp_ctx* ctx, void* handle64)
{
    if (!ctx || !handle64) return VKP_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    int r = ensure_sync_buf(ctx);
    if (r != VKP_OK) return r;
    ctx->sync_flags = ctx->sync_buf.as<uint32_t>();
    CK(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), ctx->sync_buf.p));
    return VKP_OK;
}

int vkp_sync_attach(vkp_ctx* ctx, const void* handle64)
{
    if (!ctx || !handle64) return VKP_ERR_INVALID;
    int r = ensure_sync_buf(ctx);
    if (r != VKP_OK) return r;
    void* p = nullptr;
    if ((r = vkp_ipc_open(ctx, handle64, &p)) != VKP_OK) return r;
    ctx->sync_flags = reinterpret_cast<uint32_t*>(p);
    return VKP_OK;
}

int vkp_sync_arrive(vkp_ctx* ctx, uint32_t frame)
{
    if (!ctx || !ctx->sync_flags || ctx->rank <= 0 || ctx->rank >= 64) return fail(ctx, VKP_ERR_INVALID, "vkp_sync_arrive: not an attached peer rank");
    CK(launch_sync_signal(ctx->stream, ctx->sync_flags + ctx->rank, frame));
    return VKP_OK;
}

int vkp_sync_wait_arrivals(vkp_ctx* ctx, uint32_t frame)
{
    if (!ctx || !ctx->sync_flags || ctx->rank != 0 || ctx->world > 64) return fail(ctx, VKP_ERR_INVALID, "vkp_sync_wait_arrivals: rank 0 of <= 64 only");
    CK(launch_sync_wait(ctx->stream, ctx->sync_flags + 1, ctx->world - 1, frame, ctx->sync_buf.as<uint32_t>() + 65));
    return VKP_OK;
}

int vkp_sync_release(vkp_ctx* ctx, uint32_t frame, int after_present)
{
    if (!ctx || !ctx->sync_flags || ctx->rank != 0) return fail(ctx, VKP_ERR_INVALID, "vkp_sync_release: rank 0 only");
    cudaStream_t s = (after_present && ctx->copy_stream) ? ctx->copy_stream : ctx->stream;      // behind the queued present copies
    CK(launch_sync_signal(s, ctx->sync_flags + 64, frame));
    return VKP_OK;
}

int vkp_sync_wait_release(vkp_ctx* ctx, uint32_t frame)
{
    if (!ctx || !ctx->sync_flags) return fail(ctx, VKP_ERR_INVALID, "vkp_sync_wait_release: not attached");
    CK(launch_sync_wait(ctx->stream, ctx->sync_flags + 64, 1, frame, ctx->sync_buf.as<uint32_t>() + 65));
    return VKP_OK;
}

int vkp_sync_timed_out(vkp_ctx* ctx, int* timed_out)
{
    if (!ctx || !timed_out) return VKP_ERR_INVALID;
    *timed_out = 0;
    if (!ctx->sync_buf.p) return VKP_OK;
    uint32_t v = 0;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(&v, ctx->sync_buf.as<uint32_t>() + 65, 4, cudaMemcpyDeviceToHost));
    *timed_out = (int)v;
    return VKP_OK;
}

int vkp_set_output_redirect(vkp_ctx* ctx, void* peer_buffer0, void* peer_buffer1)
{
    if (!ctx) return VKP_ERR_INVALID;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->peer_srgb8[0] = peer_buffer0; ctx->peer_srgb8[1] = peer_buffer1;
    return VKP_OK;
}

int vkp_set_output_buffer(vkp_ctx* ctx, int which)
{
    if (!ctx || which < 0 || which > 1) return VKP_ERR_INVALID;
    ctx->out_sel = which;
    return VKP_OK;
}

int vkp_get_counters(vkp_ctx* ctx, vkp_counters* out)
{
    if (!ctx || !out) return VKP_ERR_INVALID;
    memset(out, 0, sizeof(*out));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->counters.p) {
        unsigned long long c[8];
        CK(cudaMemcpy(c, ctx->counters.p, sizeof(c), cudaMemcpyDeviceToHost));
        out->rays_primary = c[0]; out->rays_shadow = c[1]; out->rays_reflect = c[2]; out->rays_secondary_shadow = c[3];
        out->node_fetches = c[4]; out->triangle_tests = c[5]; out->instance_entries = c[6];
        out->worst_thread_nodes = (uint32_t)(c[7] >> 32); out->worst_thread_tag = (uint32_t)c[7];
        if (ctx->accumulate_counters) CK(launch_stage_words(ctx->stream, ctx->counters.p, nullptr, 64));
    }
    if (ctx->profiling && ctx->n_ev >= 2) {
        for (int i = 0; i + 1 < ctx->n_ev; i++) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]));
            if (ctx->ev_stage[i] >= 0 && ctx->ev_stage[i] < 8) out->ms_stage[ctx->ev_stage[i]] += ms;
        }
        CK(cudaEventElapsedTime(&out->ms_frame, ctx->ev[0], ctx->ev[ctx->n_ev - 1]));
        if (ctx->chain_timed) {
            // the reflection chain on its own stream: [7] = its duration (queue clear + k_chain_scan + chain kernels), ms_tlas = when it ended,
            // measured from the start of the frame
            CK(cudaEventSynchronize(ctx->ev_chain_t1));
            CK(cudaEventElapsedTime(&out->ms_stage[7], ctx->ev_chain_t0, ctx->ev_chain_t1));
            CK(cudaEventElapsedTime(&out->ms_tlas, ctx->ev[0], ctx->ev_chain_t1));
            if (ctx->debug_chain_timing && ctx->ev_chain_long) {
                float a = 0.f;
                if (cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev_chain_long) == cudaSuccess) fprintf(stderr, "vkp: long-path kernel ends at %.3f ms, chain stream at %.3f ms\n", a, out->ms_tlas);
            }
            ctx->chain_timed = false;
        }
    }
    out->kernel_launches = ctx->last_launches;
    return VKP_OK;
}

int vkp_set_counter_accumulation(vkp_ctx* ctx, int on)
{
    if (!ctx) return VKP_ERR_INVALID;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->accumulate_counters = on != 0;
    if (ctx->counters.p) CK(launch_stage_words(ctx->stream, ctx->counters.p, nullptr, 64));
    return VKP_OK;
}

int vkp_set_profiling(vkp_ctx* ctx, int on)
{
    if (!ctx) return VKP_ERR_INVALID;
    ctx->profiling = on != 0;
    ctx->count_trav = on == 2;
    return VKP_OK;
}

int vkp_trace_rays(vkp_ctx* ctx, const vkp_ray* rays, uint32_t n, int any_hit, vkp_hit* hits)
{
    if (!ctx || (n && (!rays || !hits))) return fail(ctx, VKP_ERR_INVALID, "vkp_trace_rays: NULL");
    CK(cudaSetDevice(ctx->device));
    DScene S;
    // tracing needs only geometry: tolerate a missing environment / materials
    if (!ctx->tlas_built) return fail(ctx, VKP_ERR_INVALID, "vkp_trace_rays: acceleration structures are not built");
    memset(&S, 0, sizeof(S));
    { int r = scene_uniforms(ctx, S); if (r != VKP_OK) return r; }
    if (n == 0) return VKP_OK;
    CK(ctx->io_tmp.ensure(sizeof(vkp_ray) * (size_t)n));
    CK(ctx->io_tmp2.ensure(sizeof(vkp_hit) * (size_t)n));
    CK(cudaMemcpyAsync(ctx->io_tmp.p, rays, sizeof(vkp_ray) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(launch_trace_rays(ctx->stream, S, ctx->io_tmp.as<vkp_ray>(), n, any_hit, ctx->io_tmp2.as<vkp_hit>()));
    { int r = mark_scene_use(ctx); if (r != VKP_OK) return r; }
    CK(cudaMemcpyAsync(hits, ctx->io_tmp2.p, sizeof(vkp_hit) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VKP_OK;
}

int vkp_trace_stats(vkp_ctx* ctx, const vkp_ray* rays, uint32_t n, int any_hit, uint32_t* steps_and_tris)
{
    if (!ctx || (n && (!rays || !steps_and_tris))) return fail(ctx, VKP_ERR_INVALID, "vkp_trace_stats: NULL");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->tlas_built) return fail(ctx, VKP_ERR_INVALID, "vkp_trace_stats: acceleration structures are not built");
    DScene S; memset(&S, 0, sizeof(S));
    { int r = scene_uniforms(ctx, S); if (r != VKP_OK) return r; }
    if (n == 0) return VKP_OK;
    CK(ctx->io_tmp.ensure(sizeof(vkp_ray) * (size_t)n));
    CK(ctx->io_tmp2.ensure(sizeof(uint32_t) * 3 * (size_t)n));
    CK(cudaMemcpyAsync(ctx->io_tmp.p, rays, sizeof(vkp_ray) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(launch_trace_stats(ctx->stream, S, ctx->io_tmp.as<vkp_ray>(), n, any_hit, ctx->io_tmp2.as<uint32_t>()));
    { int r = mark_scene_use(ctx); if (r != VKP_OK) return r; }
    CK(cudaMemcpyAsync(steps_and_tris, ctx->io_tmp2.p, sizeof(uint32_t) * 3 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VKP_OK;
}

int vkp_selftest_sort(vkp_ctx* ctx, const uint32_t* keys, uint32_t n, int end_bit, uint32_t* sorted_keys, uint32_t* sorted_vals, uint64_t* scan)
{
    if (!ctx || (n && (!keys || !sorted_keys || !sorted_vals || !scan))) return fail(ctx, VKP_ERR_INVALID, "vkp_selftest_sort: NULL");
    CK(cudaSetDevice(ctx->device));
    CK(bvh_selftest_sort(ctx->stream, keys, n, end_bit, sorted_keys, sorted_vals, reinterpret_cast<unsigned long long*>(scan)));
    return VKP_OK;
}

int vkp_eval_math(vkp_ctx* ctx, int func, const float* x, const float* y, uint32_t n, float* out)
{
    if (!ctx || (n && (!x || !y || !out))) return fail(ctx, VKP_ERR_INVALID, "vkp_eval_math: NULL");
    if (n == 0) return VKP_OK;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->io_tmp.ensure(sizeof(float) * 2 * (size_t)n));
    CK(ctx->io_tmp2.ensure(sizeof(float) * (size_t)n));
    float* dx = ctx->io_tmp.as<float>(); float* dy = dx + n;
    CK(cudaMemcpyAsync(dx, x, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dy, y, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(launch_eval_math(ctx->stream, func, dx, dy, n, ctx->io_tmp2.as<float>()));
    CK(cudaMemcpyAsync(out, ctx->io_tmp2.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VKP_OK;
}

static Bvh* pick_bvh(vkp_ctx* ctx, int blas_index)
{
    if (blas_index < 0) return ctx->tlas_built ? &ctx->tlas : nullptr;
    if (blas_index >= (int)ctx->blas.size()) return nullptr;
    return &ctx->blas[blas_index]->bvh;
}

int vkp_bvh_info(vkp_ctx* ctx, int blas_index, uint32_t* n_nodes, uint32_t* n_prims)
{
    if (!ctx) return VKP_ERR_INVALID;
    Bvh* b = pick_bvh(ctx, blas_index);
    if (!b) return fail(ctx, VKP_ERR_INVALID, "vkp_bvh_info: no such BVH");
    if (n_nodes) *n_nodes = b->n_nodes;
    if (n_prims) *n_prims = b->n_prims;
    return VKP_OK;
}

int vkp_bvh_download(vkp_ctx* ctx, int blas_index, vkp_bvh_node* nodes, uint32_t* prim_ids)
{
    if (!ctx) return VKP_ERR_INVALID;
    Bvh* b = pick_bvh(ctx, blas_index);
    if (!b) return fail(ctx, VKP_ERR_INVALID, "vkp_bvh_download: no such BVH");
    const void* nodes_src = b->nodes.p;
    if (blas_index < 0) {      // the TLAS boxes live in the current uniform set (refitted per frame), the topology in ctx->tlas
        int r = flush_uniforms(ctx);
        if (r != VKP_OK) return r;
        if (ctx->update_stream) CK(cudaStreamSynchronize(ctx->update_stream));
        nodes_src = ctx->uni[ctx->ucur].tlas_nodes.p;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    if (nodes) CK(cudaMemcpy(nodes, nodes_src, sizeof(vkp_bvh_node) * (size_t)b->n_nodes, cudaMemcpyDeviceToHost));
    if (prim_ids && b->n_prims) CK(cudaMemcpy(prim_ids, b->items.p, sizeof(uint32_t) * (size_t)b->n_prims, cudaMemcpyDeviceToHost));
    return VKP_OK;
}

int vkp_timer_begin(vkp_ctx* ctx)
{
    if (!ctx) return VKP_ERR_INVALID;
    CK(cudaEventRecord(ctx->timer0, ctx->stream));
    return VKP_OK;
}

int vkp_timer_end(vkp_ctx* ctx, float* ms)
{
    if (!ctx || !ms) return VKP_ERR_INVALID;
    CK(cudaEventRecord(ctx->timer1, ctx->stream));
    CK(cudaEventSynchronize(ctx->timer1));
    CK(cudaEventElapsedTime(ms, ctx->timer0, ctx->timer1));
    return VKP_OK;
}

int vkp_flush_l2(vkp_ctx* ctx)
{
    if (!ctx) return VKP_ERR_INVALID;
    const size_t bytes = (size_t)256 << 20;   // 2x the 126 MB L2
    CK(ctx->flush.ensure(bytes));
    CK(launch_l2_flush(ctx->stream, ctx->flush.p, bytes));
    return VKP_OK;
}

int vkp_tiles_packed_bytes(const vkp_ctx* ctx, int plane, int rank, int world, size_t* bytes)
{
    if (!ctx || !bytes || !ctx->have_params || world < 1) return VKP_ERR_INVALID;
    size_t e = plane_elem_bytes(plane);
    if (!e) return VKP_ERR_INVALID;
    int tiles_x = (ctx->params.width + ctx->tile_w - 1) / ctx->tile_w, tiles_y = (ctx->params.height + ctx->tile_h - 1) / ctx->tile_h;
    int total = tiles_x * tiles_y;
    int own = total > rank ? (total - rank + world - 1) / world : 0;
    *bytes = (size_t)own * ctx->tile_w * ctx->tile_h * e;
    return VKP_OK;
}

int vkp_tiles_pack(vkp_ctx* ctx, int plane, void* device_dst, size_t bytes)
{
    if (!ctx || !device_dst) return VKP_ERR_INVALID;
    size_t need = 0;
    if (vkp_tiles_packed_bytes(ctx, plane, ctx->rank, ctx->world, &need) != VKP_OK || bytes < need) return fail(ctx, VKP_ERR_INVALID, "vkp_tiles_pack: bad plane / size");
    void* p = plane_ptr(ctx, plane);
    if (!p) return fail(ctx, VKP_ERR_INVALID, "vkp_tiles_pack: unknown plane");
    CK(launch_tiles_pack(ctx->stream, p, device_dst, (int)plane_elem_bytes(plane), ctx->plane_w, ctx->plane_h, ctx->tile_w, ctx->tile_h, ctx->rank, ctx->world));
    ctx->last_launches++;
    return VKP_OK;
}

int vkp_tiles_unpack(vkp_ctx* ctx, int plane, int rank, int world, const void* device_src, size_t bytes)
{
    if (!ctx || !device_src) return VKP_ERR_INVALID;
    size_t need = 0;
    if (vkp_tiles_packed_bytes(ctx, plane, rank, world, &need) != VKP_OK || bytes < need) return fail(ctx, VKP_ERR_INVALID, "vkp_tiles_unpack: bad plane / size");
    int r = ensure_planes(ctx);
    if (r != VKP_OK) return r;
    void* p = plane_ptr(ctx, plane);
    if (!p) return fail(ctx, VKP_ERR_INVALID, "vkp_tiles_unpack: unknown plane");
    CK(launch_tiles_unpack(ctx->stream, p, device_src, (int)plane_elem_bytes(plane), ctx->plane_w, ctx->plane_h, ctx->tile_w, ctx->tile_h, rank, world));
    ctx->last_launches++;
    return VKP_OK;
}

} // extern "C"

// ---- internal accessors for vkp_group.cu ---------------------------------------------------------------------------------------
cudaStream_t vkp_internal_stream(vkp_ctx* ctx) { return ctx->stream; }
// the queued present copy (vkp_readback_async) still reading sRGB8 output buffer `buf`, or NULL
cudaEvent_t vkp_internal_pending_copy(vkp_ctx* ctx, int buf) { return (buf >= 0 && buf < 2 && ctx->copy_pending[buf]) ? ctx->ev_copy_done[buf] : nullptr; }
int vkp_internal_output_buffers(vkp_ctx* ctx, void** b0, void** b1)
{
    int r = ensure_planes(ctx);
    if (r != VKP_OK) return r;
    *b0 = ctx->final_srgb8.p; *b1 = ctx->final_srgb8_alt.p;
    return VKP_OK;
}
