/*
 * vkp.h -- C-ABI of the B200 hybrid ray-tracing hot path (libvkp.so).
 *
 * The reference has no FFI; its de-facto boundary is class Renderer and the buffers it binds
 * (SURVEY.md section 8b).  Each entry point below names the reference interface it replaces.
 * Conventions: every call returns 0 on success and a negative vkp_status on failure (the reference
 * aborts through VK_CHECK, vk_engine.h:15-24; a library must not), vkp_last_error() gives the text,
 * one context == one GPU == one host thread at a time (the reference is single-threaded, single
 * queue).  Host pointers are borrowed for the duration of the call only; device memory is owned
 * by the context.  There is no CPU fallback: without a usable CUDA device vkp_create fails.
 */
#ifndef VKP_H
#define VKP_H

#include "vkp_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define VKP_API
#else
#define VKP_API __attribute__((visibility("default")))
#endif

enum vkp_status {
    VKP_OK = 0,
    VKP_ERR_INVALID = -1,   /* bad argument / call order            */
    VKP_ERR_CUDA = -2,      /* CUDA runtime error (text in last_error) */
    VKP_ERR_NO_DEVICE = -3, /* no usable sm_100 device              */
    VKP_ERR_OOM = -4
};

typedef struct vkp_ctx vkp_ctx;

/* library ------------------------------------------------------------------------------------ */
VKP_API int         vkp_abi_version(void);
VKP_API const char* vkp_last_error(const vkp_ctx* ctx);          /* ctx may be NULL: last create error */

/* VulkanEngine::init_vulkan + new Renderer (vk_engine.cpp:236, :50): one context per GPU */
VKP_API int  vkp_create(vkp_ctx** out, int device_ordinal);
VKP_API void vkp_destroy(vkp_ctx* ctx);                           /* VulkanEngine::cleanup */

/* geometry & resources --------------------------------------------------------------------- */
/* Mesh::upload (vk_mesh.cpp:265-359): vertex + index buffers; returns the mesh handle */
VKP_API int vkp_upload_mesh(vkp_ctx* ctx, const vkp_vertex* vertices, uint32_t n_vertices,
                            const uint32_t* indices, uint32_t n_indices, int* mesh_id);
/* vkutil::load_image_from_file (vk_textures.cpp:11-103): R8G8B8A8_UNORM, no mips, NEAREST/REPEAT */
VKP_API int vkp_upload_texture_rgba8(vkp_ctx* ctx, const uint8_t* rgba, int width, int height, int* tex_id);
/* hybrid descriptor binding 10, environmentTexture[0] = sky, [1] = irradiance (renderer.cpp:3073-3074) */
VKP_API int vkp_set_env(vkp_ctx* ctx, int sky_tex_id, int irradiance_tex_id);
/* binding 9 _matBuffer (renderer.cpp:1488-1499) */
VKP_API int vkp_set_materials(vkp_ctx* ctx, const vkp_material* materials, int n);
/* binding 4 _lightBuffer, rewritten every frame by VulkanEngine::update (vk_engine.cpp:146-164) */
VKP_API int vkp_set_lights(vkp_ctx* ctx, const vkp_light* lights, int n);
/* Renderer::_tlas + id buffer + matrices buffer (vk_engine.cpp:172-185, vk_mesh.cpp:506-533) */
VKP_API int vkp_set_instances(vkp_ctx* ctx, const vkp_instance* instances, int n);

/* acceleration structures ------------------------------------------------------------------- */
/* Renderer::create_bottom_acceleration_structure + buildBlas (renderer.cpp:1643-1753):
 * one BLAS per distinct (mesh, first_index, index_count) referenced by the instances */
VKP_API int vkp_build_blas(vkp_ctx* ctx);
/* BLAS hierarchy builder for subsequent vkp_build_blas calls: 0 = Morton LBVH (Karras radix tree), 1 = PLOC (default) */
VKP_API int vkp_set_builder(vkp_ctx* ctx, int builder);
/* Renderer::buildTlas(instances, flags, update) (renderer.cpp:1757-1865); update=1 == MODE_UPDATE refit */
VKP_API int vkp_build_tlas(vkp_ctx* ctx, int update);

/* build / refit times, tree heights, the BLAS arena and its L2 persistence window (syncs) */
VKP_API int vkp_get_build_stats(vkp_ctx* ctx, vkp_build_stats* out);
/* L2 persistence window over the BLAS arena (cudaStreamAttributeAccessPolicyWindow, hitProp persisting) on the context's streams:
 * what north_star calls "node and triangle fetches staged through ... the L2 persistence window".  -1 = the built-in default
 * (VKP_L2_PERSIST in the environment overrides it), 0 = off, 1 = on.  Takes effect at once; changes no result. */
VKP_API int vkp_set_l2_persistence(vkp_ctx* ctx, int on);

/* per-frame state --------------------------------------------------------------------------- */
/* VulkanEngine::updateCameraMatrices (vk_engine.cpp:688-741) */
VKP_API int vkp_set_camera(vkp_ctx* ctx, const vkp_camera* camera);
VKP_API int vkp_set_params(vkp_ctx* ctx, const vkp_params* params);
/* multi-GPU: this context renders only tiles t with (t mod world) == rank; world=1 = all */
VKP_API int vkp_set_tiling(vkp_ctx* ctx, int rank, int world, int tile_w, int tile_h);

/* frame --------------------------------------------------------------------------------------- */
/* Renderer::render / raytrace / rasterize_hybrid by params.mode (renderer.cpp:403, :465, :530).
 * Asynchronous on the context stream. */
VKP_API int vkp_render(vkp_ctx* ctx);
/* a single pass of the frame, for stage-isolated parity tests */
VKP_API int vkp_run_stage(vkp_ctx* ctx, int stage);
VKP_API int vkp_sync(vkp_ctx* ctx);                                /* vkQueueWaitIdle */
/* VulkanEngine::resetFrame (vk_engine.cpp:683): zero the denoise / accumulation history */
VKP_API int vkp_reset_history(vkp_ctx* ctx);

/* results ------------------------------------------------------------------------------------- */
VKP_API int vkp_plane_bytes(const vkp_ctx* ctx, int plane, size_t* bytes);
VKP_API int vkp_readback(vkp_ctx* ctx, int plane, void* dst, size_t bytes);       /* syncs */
/* Present with frames in flight -- the FRAME_OVERLAP = 2 of renderer.h:34 / get_current_frame() renderer.cpp:363: the copy of
 * `plane` to `dst` (pinned memory) is queued behind the frames submitted so far on a copy stream and the call returns at once;
 * vkp_readback_wait blocks until at most `keep_in_flight` of the queued copies are still outstanding (0 = all landed).  For FINAL_SRGB8 alternate vkp_set_output_buffer(ctx, k & 1) per
 * frame: frame k+1 then renders into the other buffer while frame k is still on the bus.  A later vkp_render that would
 * overwrite a buffer still being copied waits for that copy on the device. */
VKP_API int vkp_readback_async(vkp_ctx* ctx, int plane, void* dst, size_t bytes);
VKP_API int vkp_readback_wait(vkp_ctx* ctx, int keep_in_flight);
VKP_API int vkp_upload_plane(vkp_ctx* ctx, int plane, const void* src, size_t bytes);
VKP_API int vkp_plane_device_ptr(vkp_ctx* ctx, int plane, void** device_ptr, size_t* bytes);
VKP_API int vkp_get_counters(vkp_ctx* ctx, vkp_counters* out);                    /* syncs */
/* on: the ray counters are not cleared per frame but run on until vkp_get_counters reads them (which also clears them): a frame
 * loop can then be timed with events only, without a host round trip per frame */
VKP_API int vkp_set_counter_accumulation(vkp_ctx* ctx, int on);
VKP_API int vkp_set_profiling(vkp_ctx* ctx, int on);            /* 0 off, 1 per-stage CUDA events, 2 also count traversal fetches (slower) */

/* traceRayEXT made explicit: trace n rays against the current TLAS (host buffers). any_hit != 0
 * == gl_RayFlagsTerminateOnFirstHitEXT|SkipClosestHitShaderEXT: only hit.t >= 0 / < 0 is defined */
VKP_API int vkp_trace_rays(vkp_ctx* ctx, const vkp_ray* rays, uint32_t n, int any_hit, vkp_hit* hits);

/* instrumented traversal: per ray {node pairs fetched, triangles tested, SM cycles} (3 x uint32 per ray), for build-quality
 * and latency studies */
VKP_API int vkp_trace_stats(vkp_ctx* ctx, const vkp_ray* rays, uint32_t n, int any_hit, uint32_t* steps_and_tris);

/* BVH introspection (tests): copy out the nodes and the leaf-ordered primitive ids of one BLAS
 * (index = order of first use in vkp_set_instances) or of the TLAS (blas_index = -1) */
VKP_API int vkp_bvh_info(vkp_ctx* ctx, int blas_index, uint32_t* n_nodes, uint32_t* n_prims);
VKP_API int vkp_bvh_download(vkp_ctx* ctx, int blas_index, vkp_bvh_node* nodes, uint32_t* prim_ids);

/* device-side timing on the context stream (CUDA events) */
VKP_API int vkp_timer_begin(vkp_ctx* ctx);
VKP_API int vkp_timer_end(vkp_ctx* ctx, float* elapsed_ms);                       /* syncs */
/* write > L2-size bytes so the next timed frame starts with a cold L2 */
VKP_API int vkp_flush_l2(vkp_ctx* ctx);

/* multi-GPU composite: gather this rank's tiles of a plane into a dense buffer
 * (pack) and scatter packed tiles of rank r of `world` into the full-frame plane (unpack).
 * Buffers are DEVICE pointers so they can be handed to NCCL / peer copies without staging. */
VKP_API int vkp_tiles_packed_bytes(const vkp_ctx* ctx, int plane, int rank, int world, size_t* bytes);
VKP_API int vkp_tiles_pack(vkp_ctx* ctx, int plane, void* device_dst, size_t bytes);
VKP_API int vkp_tiles_unpack(vkp_ctx* ctx, int plane, int rank, int world, const void* device_src, size_t bytes);
/* multi-GPU composite without a gather: rank 0 exports its FINAL_SRGB8 plane (64-byte CUDA IPC handle), the other ranks
 * open it and redirect their own FINAL_SRGB8 stores to it, so the resolve kernels write finished pixels straight into
 * rank 0's image over NVLink; the only per-frame collective left is a barrier.  NULL restores the local planes. */
VKP_API int vkp_ipc_export_plane(vkp_ctx* ctx, int plane, void* handle64);
VKP_API int vkp_ipc_open(vkp_ctx* ctx, const void* handle64, void** device_ptr);
VKP_API int vkp_set_output_redirect(vkp_ctx* ctx, void* peer_buffer0, void* peer_buffer1);
/* FINAL_SRGB8 is double-buffered: select which buffer (0 / 1) the next frame writes and vkp_readback / export read.  With peer
 * stores, alternating per frame lets rank 0 consume frame k while the peers already write frame k+1 (one barrier per frame). */
VKP_API int vkp_set_output_buffer(vkp_ctx* ctx, int which);
/* Frame barrier of the peer composite, on the device over NVLink (no collective call, no host round trip per frame).  Rank 0
 * exports a small flag array; the others attach to it.  All calls are stream-ordered kernels on the context's stream:
 *   peer rank:  [vkp_sync_wait_release(f - 2)]  render frame f (pixels land in rank 0's buffer f & 1)  vkp_sync_arrive(f)
 *   rank 0:     render frame f   vkp_sync_wait_arrivals(f)   <frame complete: present it>   vkp_sync_release(f, after_present)
 * after_present != 0 queues the release behind the present copies (vkp_readback_async).  A wait gives up after ~10 s and sets a
 * flag vkp_sync_timed_out reports (a dead peer must not hang the GPU). */
VKP_API int vkp_sync_export(vkp_ctx* ctx, void* handle64);
VKP_API int vkp_sync_attach(vkp_ctx* ctx, const void* handle64);
VKP_API int vkp_sync_arrive(vkp_ctx* ctx, uint32_t frame);
VKP_API int vkp_sync_wait_arrivals(vkp_ctx* ctx, uint32_t frame);
VKP_API int vkp_sync_release(vkp_ctx* ctx, uint32_t frame, int after_present);
VKP_API int vkp_sync_wait_release(vkp_ctx* ctx, uint32_t frame);
VKP_API int vkp_sync_timed_out(vkp_ctx* ctx, int* timed_out);
/* build-primitive probe (tests): stable radix sort of (keys[i], i) on bits [0, end_bit) and an exclusive 64-bit prefix sum of
 * ((keys[i] & 1) << 32 | (keys[i] & 3)) with the BVH builder's own kernels (bvh_sort.cuh) */
VKP_API int vkp_selftest_sort(vkp_ctx* ctx, const uint32_t* keys, uint32_t n, int end_bit, uint32_t* sorted_keys, uint32_t* sorted_vals, uint64_t* scan);
/* numerics-contract probe (tests): evaluate one vkp_detmath.h function on the DEVICE over host arrays.
 * func: 0 sin, 1 cos, 2 atan2(x,y), 3 asin, 4 pow(x,y), 5 log2, 6 exp2, 7 f32->f16->f32, 8 unorm8 */
VKP_API int vkp_eval_math(vkp_ctx* ctx, int func, const float* x, const float* y, uint32_t n, float* out);
/* run subsequent work on a caller-owned cudaStream_t (e.g. torch's current stream); NULL = own stream */
VKP_API int vkp_set_stream(vkp_ctx* ctx, void* cuda_stream);

/* multi-GPU group ------------------------------------------------------------------------------
 * Several GPUs of one box behind one handle, in ONE process: what the C++ host (Renderer, vkpinut --gpus N) uses where the
 * reference has its single device and queue (vk_engine.cpp:296-302, frame seam Renderer::rasterize_hybrid renderer.cpp:530-600).
 * One vkp_ctx per device, scene and BVH replicated, screen tiles dealt t -> rank t % n; the resolve kernels of ranks 1..n-1
 * store their sRGB8 pixels straight into rank 0's double-buffered image over NVLink peer memory (cudaDeviceEnablePeerAccess),
 * ordered by one CUDA event per rank; Hybrid reset frames exchange their shadow / history tiles between the shadow pass and the
 * filter.  Every vkp_group_<call> below means vkp_<call> on every member.  Images are bit-identical to one GPU. */
typedef struct vkp_group vkp_group;
VKP_API int      vkp_group_create(vkp_group** out, const int* device_ordinals, int n_devices, int tile_w, int tile_h);
VKP_API void     vkp_group_destroy(vkp_group* g);
VKP_API int      vkp_group_size(const vkp_group* g);
VKP_API vkp_ctx* vkp_group_member(vkp_group* g, int rank);           /* rank 0 owns the composited image */
VKP_API const char* vkp_group_last_error(const vkp_group* g);
VKP_API int vkp_group_upload_mesh(vkp_group* g, const vkp_vertex* vertices, uint32_t n_vertices, const uint32_t* indices, uint32_t n_indices, int* mesh_id);
VKP_API int vkp_group_upload_texture_rgba8(vkp_group* g, const uint8_t* rgba, int width, int height, int* tex_id);
VKP_API int vkp_group_set_env(vkp_group* g, int sky_tex_id, int irradiance_tex_id);
VKP_API int vkp_group_set_materials(vkp_group* g, const vkp_material* materials, int n);
VKP_API int vkp_group_set_lights(vkp_group* g, const vkp_light* lights, int n);
VKP_API int vkp_group_set_instances(vkp_group* g, const vkp_instance* instances, int n);
VKP_API int vkp_group_build_blas(vkp_group* g);
VKP_API int vkp_group_build_tlas(vkp_group* g, int update);
VKP_API int vkp_group_set_camera(vkp_group* g, const vkp_camera* camera);
VKP_API int vkp_group_set_params(vkp_group* g, const vkp_params* params);
VKP_API int vkp_group_reset_history(vkp_group* g);
/* one frame of params.mode across the group; asynchronous.  Afterwards rank 0's stream is ordered behind every rank's part. */
VKP_API int vkp_group_render(vkp_group* g);
VKP_API int vkp_group_sync(vkp_group* g);
/* FINAL_SRGB8 is read from rank 0 (already composited); any other plane is gathered to rank 0 first (pack, peer copy, unpack) */
VKP_API int vkp_group_readback(vkp_group* g, int plane, void* dst, size_t bytes);
VKP_API int vkp_group_get_counters(vkp_group* g, vkp_counters* out);            /* ray / launch counters summed over the ranks */

#ifdef __cplusplus
}
#endif
#endif /* VKP_H */
