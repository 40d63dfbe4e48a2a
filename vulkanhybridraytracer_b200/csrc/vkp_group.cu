// vkp_group.cu -- several B200s of one box behind ONE handle, in one process (include/vkp.h "multi-GPU group").
//
// The reference renders on a single device (one VkQueue, vk_engine.cpp:296-302); its seam for a frame is
// Renderer::rasterize_hybrid / raytrace / render (renderer.cpp:403, :465, :530).  A group keeps one vkp_ctx per GPU with the
// scene and the BVH replicated, deals the screen tiles t -> rank t % n (vkp_set_tiling) and composites without a gather:
// peer access to rank 0 is enabled once (cudaDeviceEnablePeerAccess -- one process, no IPC handles) and the resolve kernels of
// the other ranks store their finished sRGB8 pixels straight into rank 0's double-buffered image over NVLink
// (vkp_set_output_redirect).  What is left of the composite is ordering, and inside one process that is a CUDA event per rank
// which rank 0's stream waits for -- no collective call, no host round trip, no flag kernels.
//
// Reset frames of the Hybrid pipeline (denoise_frame <= 0, blur.comp:49-80) read the 3x3 neighbourhood of the shadow planes
// and the reprojected history across tile borders: there the ranks exchange their tiles of those planes (pack, peer copy,
// unpack) between the shadow pass and the filter, exactly like the one-process-per-GPU driver (multigpu.py).
#include "vkp_internal.h"

#include <string>
#include <vector>

struct vkp_group {
    std::vector<vkp_ctx*> ctx;
    std::vector<int> dev;
    std::vector<cudaEvent_t> ev_done;       // rank r finished its part of the current frame
    std::vector<cudaEvent_t> ev_packed;     // rank r packed its tiles for an exchange
    cudaEvent_t ev_free[2] = {nullptr, nullptr};     // rank 0 passed the end of the frame that last used output buffer b
    bool have_free[2] = {false, false};
    std::vector<DevBuf> send;               // [rank]
    std::vector<std::vector<DevBuf>> recv;  // [to][from]
    std::vector<char> peer_ok;              // rank r can store into rank 0's memory
    vkp_params params{};
    bool have_params = false;
    int n_lights = 0;
    int tile_w = 32, tile_h = 16;
    unsigned frame_no = 0;
    std::string err;
};

namespace {

int gfail(vkp_group* g, int code, const std::string& msg) { if (g) g->err = msg; return code; }
int gfrom(vkp_group* g, int rank, int code)
{
    if (code != VKP_OK) g->err = std::string("rank ") + std::to_string(rank) + ": " + vkp_last_error(g->ctx[rank]);
    return code;
}
int gcuda(vkp_group* g, cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return VKP_OK;
    g->err = std::string(what) + ": " + cudaGetErrorString(e);
    return VKP_ERR_CUDA;
}
#define GK(call) do { int _r = gcuda(g, (call), #call); if (_r != VKP_OK) return _r; } while (0)
#define GC(rank, call) do { cudaSetDevice(g->dev[rank]); int _r = gfrom(g, rank, (call)); if (_r != VKP_OK) return _r; } while (0)

// every rank's tiles of `planes` to every other rank (all-gather): pack on the owner, one peer copy per (from, to) pair, unpack
int exchange_planes(vkp_group* g, const std::vector<int>& planes)
{
    const int n = (int)g->ctx.size();
    std::vector<size_t> sz(n, 0);
    for (int r = 0; r < n; r++) {
        size_t b = 0;
        for (int p : planes) { size_t one = 0; GC(r, vkp_tiles_packed_bytes(g->ctx[r], p, r, n, &one)); b = one > b ? one : b; }
        sz[r] = (b + 255u) & ~(size_t)255u;             // per-plane stride in the rank's buffer
    }
    for (int r = 0; r < n; r++) {
        GK(cudaSetDevice(g->dev[r]));
        GK(g->send[r].ensure(sz[r] * planes.size() + 256));
        for (size_t k = 0; k < planes.size(); k++)
            GC(r, vkp_tiles_pack(g->ctx[r], planes[k], static_cast<char*>(g->send[r].p) + k * sz[r], sz[r]));
        GK(cudaEventRecord(g->ev_packed[r], vkp_internal_stream(g->ctx[r])));
    }
    for (int q = 0; q < n; q++) {
        GK(cudaSetDevice(g->dev[q]));
        cudaStream_t sq = vkp_internal_stream(g->ctx[q]);
        for (int r = 0; r < n; r++) {
            if (r == q) continue;
            GK(g->recv[q][r].ensure(sz[r] * planes.size() + 256));
            GK(cudaStreamWaitEvent(sq, g->ev_packed[r], 0));
            GK(cudaMemcpyPeerAsync(g->recv[q][r].p, g->dev[q], g->send[r].p, g->dev[r], sz[r] * planes.size(), sq));
            for (size_t k = 0; k < planes.size(); k++)
                GC(q, vkp_tiles_unpack(g->ctx[q], planes[k], r, n, static_cast<char*>(g->recv[q][r].p) + k * sz[r], sz[r]));
        }
    }
    // the send buffers are reused by the next exchange: exchanges are rare (reset frames, read-backs of gathered planes), wait here
    for (int r = 0; r < n; r++) { GK(cudaSetDevice(g->dev[r])); GK(cudaStreamSynchronize(vkp_internal_stream(g->ctx[r]))); }
    return VKP_OK;
}

// the other ranks' tiles of `plane` into rank 0's copy of it
int gather_plane_to_root(vkp_group* g, int plane)
{
    const int n = (int)g->ctx.size();
    cudaStream_t s0 = vkp_internal_stream(g->ctx[0]);
    for (int r = 1; r < n; r++) {
        size_t b = 0;
        GC(r, vkp_tiles_packed_bytes(g->ctx[r], plane, r, n, &b));
        GK(cudaSetDevice(g->dev[r]));
        GK(g->send[r].ensure(b + 256));
        GC(r, vkp_tiles_pack(g->ctx[r], plane, g->send[r].p, b));
        GK(cudaEventRecord(g->ev_packed[r], vkp_internal_stream(g->ctx[r])));
        GK(cudaSetDevice(g->dev[0]));
        GK(g->recv[0][r].ensure(b + 256));
        GK(cudaStreamWaitEvent(s0, g->ev_packed[r], 0));
        GK(cudaMemcpyPeerAsync(g->recv[0][r].p, g->dev[0], g->send[r].p, g->dev[r], b, s0));
        GC(0, vkp_tiles_unpack(g->ctx[0], plane, r, n, g->recv[0][r].p, b));
    }
    GK(cudaSetDevice(g->dev[0]));
    GK(cudaStreamSynchronize(s0));
    return VKP_OK;
}

} // namespace

extern "C" {

int vkp_group_create(vkp_group** out, const int* device_ordinals, int n_devices, int tile_w, int tile_h)
{
    if (!out || !device_ordinals || n_devices < 1 || n_devices > 64) return VKP_ERR_INVALID;
    *out = nullptr;
    vkp_group* g = new vkp_group();
    g->tile_w = tile_w > 0 ? tile_w : 32; g->tile_h = tile_h > 0 ? tile_h : 16;
    for (int r = 0; r < n_devices; r++) {
        vkp_ctx* c = nullptr;
        int e = vkp_create(&c, device_ordinals[r]);
        if (e != VKP_OK) { vkp_group_destroy(g); return e; }       // vkp_last_error(NULL) has the text
        g->ctx.push_back(c); g->dev.push_back(device_ordinals[r]);
        if ((e = vkp_set_tiling(c, r, n_devices, g->tile_w, g->tile_h)) != VKP_OK) { vkp_group_destroy(g); return e; }
    }
    g->ev_done.assign(n_devices, nullptr); g->ev_packed.assign(n_devices, nullptr);
    g->send.resize(n_devices); g->recv.resize(n_devices);
    for (auto& v : g->recv) v.resize(n_devices);
    g->peer_ok.assign(n_devices, 0);
    for (int r = 0; r < n_devices; r++) {
        cudaSetDevice(g->dev[r]);
        cudaEventCreateWithFlags(&g->ev_done[r], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&g->ev_packed[r], cudaEventDisableTiming);
        if (r == 0) { cudaEventCreateWithFlags(&g->ev_free[0], cudaEventDisableTiming); cudaEventCreateWithFlags(&g->ev_free[1], cudaEventDisableTiming); continue; }
        int can = 0;
        if (g->dev[r] != g->dev[0] && cudaDeviceCanAccessPeer(&can, g->dev[r], g->dev[0]) == cudaSuccess && can) {
            cudaError_t e = cudaDeviceEnablePeerAccess(g->dev[0], 0);
            if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) g->peer_ok[r] = 1;
            cudaGetLastError();
        } else if (g->dev[r] == g->dev[0]) g->peer_ok[r] = 1;      // several ranks on one GPU (tests): same address space
    }
    *out = g;
    return VKP_OK;
}

void vkp_group_destroy(vkp_group* g)
{
    if (!g) return;
    for (size_t r = 0; r < g->ctx.size(); r++) {
        cudaSetDevice(g->dev[r]);
        if (g->ctx[r]) { vkp_sync(g->ctx[r]); }
    }
    for (size_t r = 0; r < g->ctx.size(); r++) {
        cudaSetDevice(g->dev[r]);
        if (r < g->send.size()) g->send[r].release();
        if (r < g->recv.size()) for (auto& b : g->recv[r]) b.release();
        if (r < g->ev_done.size() && g->ev_done[r]) cudaEventDestroy(g->ev_done[r]);
        if (r < g->ev_packed.size() && g->ev_packed[r]) cudaEventDestroy(g->ev_packed[r]);
        if (g->ctx[r]) vkp_destroy(g->ctx[r]);
    }
    for (auto& e : g->ev_free) if (e) cudaEventDestroy(e);
    delete g;
}

int vkp_group_size(const vkp_group* g) { return g ? (int)g->ctx.size() : 0; }
vkp_ctx* vkp_group_member(vkp_group* g, int rank) { return (g && rank >= 0 && rank < (int)g->ctx.size()) ? g->ctx[rank] : nullptr; }
const char* vkp_group_last_error(const vkp_group* g) { return g ? g->err.c_str() : vkp_last_error(nullptr); }

#define FOR_ALL(call) do { if (!g) return VKP_ERR_INVALID; for (int r = 0; r < (int)g->ctx.size(); r++) { vkp_ctx* c = g->ctx[r]; GC(r, call); } return VKP_OK; } while (0)

int vkp_group_upload_mesh(vkp_group* g, const vkp_vertex* v, uint32_t nv, const uint32_t* idx, uint32_t ni, int* mesh_id) { FOR_ALL(vkp_upload_mesh(c, v, nv, idx, ni, mesh_id)); }
int vkp_group_upload_texture_rgba8(vkp_group* g, const uint8_t* rgba, int w, int h, int* tex_id) { FOR_ALL(vkp_upload_texture_rgba8(c, rgba, w, h, tex_id)); }
int vkp_group_set_env(vkp_group* g, int sky, int irr) { FOR_ALL(vkp_set_env(c, sky, irr)); }
int vkp_group_set_materials(vkp_group* g, const vkp_material* m, int n) { FOR_ALL(vkp_set_materials(c, m, n)); }
int vkp_group_set_lights(vkp_group* g, const vkp_light* l, int n) { if (g) g->n_lights = n; FOR_ALL(vkp_set_lights(c, l, n)); }
int vkp_group_set_instances(vkp_group* g, const vkp_instance* inst, int n) { FOR_ALL(vkp_set_instances(c, inst, n)); }
int vkp_group_build_blas(vkp_group* g) { FOR_ALL(vkp_build_blas(c)); }
int vkp_group_build_tlas(vkp_group* g, int update) { FOR_ALL(vkp_build_tlas(c, update)); }
int vkp_group_set_camera(vkp_group* g, const vkp_camera* cam) { FOR_ALL(vkp_set_camera(c, cam)); }
int vkp_group_set_params(vkp_group* g, const vkp_params* p) { if (g && p) { g->params = *p; g->have_params = true; } FOR_ALL(vkp_set_params(c, p)); }
int vkp_group_reset_history(vkp_group* g) { FOR_ALL(vkp_reset_history(c)); }
int vkp_group_sync(vkp_group* g) { FOR_ALL(vkp_sync(c)); }

int vkp_group_render(vkp_group* g)
{
    if (!g) return VKP_ERR_INVALID;
    if (!g->have_params) return gfail(g, VKP_ERR_INVALID, "vkp_group_set_params has not been called");
    const int n = (int)g->ctx.size();
    if (n == 1) { GC(0, vkp_render(g->ctx[0])); return VKP_OK; }
    g->frame_no++;
    const int buf = (int)(g->frame_no & 1u);
    cudaStream_t s0 = vkp_internal_stream(g->ctx[0]);
    // rank 0's two output buffers; the peers store their pixels into the one of this frame
    void *b0 = nullptr, *b1 = nullptr;
    GC(0, vkp_internal_output_buffers(g->ctx[0], &b0, &b1));
    bool all_peer = true;
    for (int r = 0; r < n; r++) {
        GC(r, vkp_set_output_buffer(g->ctx[r], buf));
        if (r == 0) continue;
        if (g->peer_ok[r]) GC(r, vkp_set_output_redirect(g->ctx[r], b0, b1)); else all_peer = false;
        // the buffer is free again once rank 0 has passed the end of the frame that used it last (two frames ago)
        if (g->have_free[buf]) { GK(cudaSetDevice(g->dev[r])); GK(cudaStreamWaitEvent(vkp_internal_stream(g->ctx[r]), g->ev_free[buf], 0)); }
        // ... and a present copy of that frame queued on rank 0 (vkp_readback_async) has read it
        if (cudaEvent_t pc = vkp_internal_pending_copy(g->ctx[0], buf)) { GK(cudaSetDevice(g->dev[r])); GK(cudaStreamWaitEvent(vkp_internal_stream(g->ctx[r]), pc, 0)); }
    }
    const bool reset = g->params.mode == VKP_MODE_HYBRID && g->params.denoise_frame <= 0;
    if (!reset) {
        for (int r = 0; r < n; r++) GC(r, vkp_render(g->ctx[r]));
    } else {
        for (int r = 0; r < n; r++) { GC(r, vkp_run_stage(g->ctx[r], VKP_STAGE_GBUFFER)); GC(r, vkp_run_stage(g->ctx[r], VKP_STAGE_SHADOW)); }
        std::vector<int> planes;
        for (int l = 0; l < g->n_lights; l++) { planes.push_back(VKP_PLANE_SHADOW0 + l); planes.push_back(VKP_PLANE_DENOISED0 + l); }
        if (!planes.empty()) { int e = exchange_planes(g, planes); if (e != VKP_OK) return e; }
        for (int r = 0; r < n; r++) { GC(r, vkp_run_stage(g->ctx[r], VKP_STAGE_DENOISE)); GC(r, vkp_run_stage(g->ctx[r], VKP_STAGE_RESOLVE)); }
    }
    // composite: the pixels are already in rank 0's image; rank 0's stream passes this point when every rank's frame is complete
    for (int r = 1; r < n; r++) {
        GK(cudaSetDevice(g->dev[r]));
        GK(cudaEventRecord(g->ev_done[r], vkp_internal_stream(g->ctx[r])));
        GK(cudaSetDevice(g->dev[0]));
        GK(cudaStreamWaitEvent(s0, g->ev_done[r], 0));
    }
    if (!all_peer) {       // no peer access to rank 0 from some GPU: gather that image the classic way
        int e = gather_plane_to_root(g, VKP_PLANE_FINAL_SRGB8);
        if (e != VKP_OK) return e;
    }
    GK(cudaSetDevice(g->dev[0]));
    GK(cudaEventRecord(g->ev_free[buf], s0));
    g->have_free[buf] = true;
    return VKP_OK;
}

int vkp_group_readback(vkp_group* g, int plane, void* dst, size_t bytes)
{
    if (!g) return VKP_ERR_INVALID;
    const int n = (int)g->ctx.size();
    if (n > 1 && plane != VKP_PLANE_FINAL_SRGB8) {
        int e = gather_plane_to_root(g, plane);
        if (e != VKP_OK) return e;
    }
    GC(0, vkp_readback(g->ctx[0], plane, dst, bytes));
    return VKP_OK;
}

int vkp_group_get_counters(vkp_group* g, vkp_counters* out)
{
    if (!g || !out) return VKP_ERR_INVALID;
    vkp_counters sum{};
    for (int r = 0; r < (int)g->ctx.size(); r++) {
        vkp_counters c{};
        GC(r, vkp_get_counters(g->ctx[r], &c));
        if (r == 0) sum = c;
        else {
            sum.rays_primary += c.rays_primary; sum.rays_shadow += c.rays_shadow; sum.rays_reflect += c.rays_reflect;
            sum.rays_secondary_shadow += c.rays_secondary_shadow; sum.kernel_launches += c.kernel_launches;
            sum.node_fetches += c.node_fetches; sum.triangle_tests += c.triangle_tests; sum.instance_entries += c.instance_entries;
        }
    }
    *out = sum;
    return VKP_OK;
}

} // extern "C"
