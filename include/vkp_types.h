/*
 * vkp_types.h -- plain-old-data layouts crossing the C-ABI.
 *
 * Every struct mirrors a buffer the reference binds in its hybrid / shadow
 * descriptor sets (renderer.cpp:2932-2958, :1932-1949); byte layouts follow the
 * reference structs so a maintainer can memcpy the same host data.
 * Matrices are column-major float[16] (GLM convention, m[col*4+row]).
 */
#ifndef VKP_TYPES_H
#define VKP_TYPES_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/vk_mesh.h:21-27  Vertex {vec3 position, normal, color; vec2 uv} = 44 B */
typedef struct vkp_vertex {
    float position[3];
    float normal[3];
    float color[3];
    float uv[2];
} vkp_vertex;

/* src/material.h:5-10 GPUMaterial, packed by Material::materialToShader (material.cpp:19-26):
 * diffuse = (r,g,b, ior); textures = (diffuse, normal, emissive, metallicRoughness; -1 = none);
 * shadingMetallicRoughness = (shadingModel, metallic, roughness, materialIndex). 48 B */
typedef struct vkp_material {
    float diffuse[4];
    float textures[4];
    float shadingMetallicRoughness[4];
} vkp_material;

/* src/vk_engine.h:76-80 uboLight (std140 Light, helpers.glsl:18-22). 48 B
 * position.w = maxDistance (point) or -1 (directional); color.w = intensity. */
typedef struct vkp_light {
    float position[4];
    float color[4];
    float radius;
    float _pad[3];
} vkp_light;

/* One TLAS instance == one primitive (vk_mesh.cpp:472-492, renderer.cpp:1895-1916) plus the
 * id record (instanceID, materialID, transformID, firstIndex) of vk_mesh.cpp:521-533 and the
 * model / inverse pair the reference pushes as constants (vk_mesh.h:55-58 ModelMatrices). */
typedef struct vkp_instance {
    int32_t  mesh_id;       /* handle returned by vkp_upload_mesh                      */
    uint32_t first_index;   /* Primitive::firstIndex                                   */
    uint32_t index_count;   /* Primitive::indexCount (3 * triangles)                   */
    int32_t  material_id;   /* Primitive::materialID                                   */
    float    model[16];     /* obj->m_matrix * node global matrix                      */
    float    inv_model[16]; /* inverse(model), computed by the host (glm::inverse)     */
} vkp_instance;

/* GPUCameraData (vk_engine.h:48-54, 256 B) + RTCameraData (vk_engine.h:56-62) */
typedef struct vkp_camera {
    float view[16];
    float proj[16];        /* already with proj[1][1] *= -1 (vk_engine.cpp:695) */
    float prev_view[16];
    float prev_proj[16];
    float inv_view[16];
    float inv_proj[16];
    float position[4];     /* _cameraPositionBuffer (vk_engine.cpp:699-702) */
} vkp_camera;

/* src/vk_engine.h:42-46 */
enum vkp_render_mode { VKP_MODE_DEFERRED = 0, VKP_MODE_RAYTRACING = 1, VKP_MODE_HYBRID = 2 };

typedef struct vkp_params {
    int32_t width, height;
    int32_t mode;                    /* vkp_render_mode                                            */
    int32_t spp;                     /* NSAMPLES (helpers.glsl:6) made runtime; raytraced mode     */
    int32_t shadow_samples;          /* VulkanEngine::_samples, 1..64 (renderer.cpp:656)           */
    int32_t max_bounces;             /* MAX_RECURSION (helpers.glsl:8) = 8                         */
    int32_t frame;                   /* RTCameraData.frame = _denoise ? _denoise_frame : 0         */
    int32_t denoise_frame;           /* blur.comp frameBuffer.frame = _denoise_frame (always)      */
    int32_t faithful_rng;            /* 1: sampleDisk gets the seed by value (helpers.glsl:82)     */
    int32_t denoise_cover_remainder; /* 0: dispatch W/16 x H/16 groups exactly (renderer.cpp:2672) */
    int32_t debug_target;            /* deferred.frag debug.target 0..6                            */
    int32_t wavefront;               /* reflection chain (hybridRaygen.rgen:153-170): 0 = inline in the resolve kernel,
                                        1 = wavefront path queues with ray compaction, one launch per bounce,
                                        2 = mirror / glass pixels compacted into a path queue, one thread per path, on a second stream
                                            overlapping shadow + denoise (default of the host engine),
                                        3 = as 2 with one warp per path and warp-cooperative traversal (measured slower, kept for A/B),
                                        4 = as 2 with the first bounce staged as a wavefront: closest hit, per-light any-hit rays and hit
                                            shading in three lean kernels, the paths that scatter compacted for the remaining bounces,
                                        5 = as 2, and the resolve does not wait for the chain either: mirror / glass pixels are finished by
                                            k_resolve_paths once the chain is there (the chain overlaps shadow + denoise + resolve),
                                        6 = as 5 with the chain as a persistent wavefront: one bounce per work item, a queue-fed persistent
                                            grid, scattered paths appended to the queue (ballot / prefix compaction) instead of looping,
                                        7 = as 5 with the paths split by what they did in the previous frame: those that went beyond their first
                                            bounce start first in k_chain, the others go through the lean staged kernels of 4 on a third stream,
                                        8 = as 7 with the first bounce folded into the frame's own kernels: k_gbuffer queues the path pixels, the first
                                            reflection rays are the leading blocks of k_shadow's grid, the rest follows the shadow pass */
    int32_t fuse_denoise;            /* vkp_render (Hybrid): run blur.comp inside the resolve kernel instead of as its own pass */
    int32_t reproject_fix;           /* blur.comp:52-53 on a reset frame.  0 (default): the reference's history fetch, pixel + motion.x * (W, H)
                                        (twice the right distance, x component used for both axes).
                                        1: the previous frame's pixel, floor(pixel + 0.5 + 0.5 * motion * (W, H)); the stored motion is
                                        previous - current NDC position (basic.vert / geometry_shader.frag cross the two varyings)      */
    int32_t denoiser;                /* shadow filter: 0 = the reference's blur.comp (3x3 clamp + temporal blend), 1 = edge-aware a-trous
                                        filter over the R8 planes guided by the G-buffer */
    int32_t reserved[1];
} vkp_params;

/* Planes readable with vkp_readback / writable with vkp_upload_plane.  Layouts are row-major,
 * pixel (x,y) at index y*W+x, formats follow renderer.cpp:255-261 and :1535-1538. */
enum vkp_plane {
    VKP_PLANE_POSITION   = 0,  /* 4 x fp16 : world pos xyz, materialIdx              (8 B/px) */
    VKP_PLANE_NORMAL     = 1,  /* 4 x fp16 : N*0.5+0.5, 1   (background: 0,0,0,0)    (8 B/px) */
    VKP_PLANE_ALBEDO     = 2,  /* 4 x u8   : albedo rgb, 255                         (4 B/px) */
    VKP_PLANE_MOTION     = 3,  /* 2 x fp16 : motion*0.5+0.5                          (4 B/px) */
    VKP_PLANE_MATERIAL   = 4,  /* 4 x u8   : (0, roughness, metallic, 255)           (4 B/px) */
    VKP_PLANE_EMISSIVE   = 5,  /* 4 x u8   : emissive rgb, 255                       (4 B/px) */
    VKP_PLANE_HIT_IDS    = 6,  /* 2 x i32  : primary hit (instance, primitive), -1 = miss (debug, 8 B/px) */
    VKP_PLANE_FINAL_F32  = 7,  /* 4 x f32  : linear radiance before the 8-bit store (16 B/px) */
    VKP_PLANE_FINAL_SRGB8= 8,  /* 4 x u8   : RGBA after unorm8 store + postFragment.frag gamma */
    VKP_PLANE_SHADOW0    = 16, /* + light index: 1 x u8 R8_UNORM shadowImage[i]       (1 B/px) */
    VKP_PLANE_DENOISED0  = 48  /* + light index: 1 x u8 R8_UNORM _denoisedImages[i]   (1 B/px) */
};
#define VKP_MAX_LIGHTS 32

/* Pipeline stages, for stage-isolated parity tests (vkp_run_stage) */
enum vkp_stage {
    VKP_STAGE_GBUFFER  = 0,  /* a3: basic.vert + geometry_shader.frag + skybox.frag        */
    VKP_STAGE_SHADOW   = 1,  /* a4: shadowRaygen.rgen + shadow.rmiss                        */
    VKP_STAGE_DENOISE  = 2,  /* a5: blur.comp                                               */
    VKP_STAGE_RESOLVE  = 3,  /* a6+a7+a8: hybridRaygen.rgen + hybridHit.rchit + miss.rmiss  */
    VKP_STAGE_POST     = 4,  /* a10: postFragment.frag                                      */
    VKP_STAGE_RAYTRACE = 5,  /* a11: raygen.rgen + closesthit.rchit                         */
    VKP_STAGE_DEFERRED = 6   /* a12: deferred.frag                                          */
};

/* Ray / hit records for vkp_trace_rays (the traceRayEXT contract made explicit) */
typedef struct vkp_ray {
    float origin[3]; float tmin;
    float dir[3];    float tmax;
} vkp_ray;

typedef struct vkp_hit {
    float   t;          /* gl_HitTEXT, or -1 on miss                         */
    float   u, v;       /* hitAttributeEXT attribs.xy                        */
    int32_t instance;   /* gl_InstanceCustomIndexEXT, -1 on miss             */
    int32_t primitive;  /* gl_PrimitiveID, -1 on miss                        */
} vkp_hit;

typedef struct vkp_counters {
    uint64_t rays_primary;          /* our G-buffer visibility rays (reported separately)  */
    uint64_t rays_shadow;           /* shadowRaygen.rgen:78                                */
    uint64_t rays_reflect;          /* hybridRaygen.rgen:158 / raygen.rgen:49              */
    uint64_t rays_secondary_shadow; /* hybridHit.rchit:123 / closesthit.rchit:111          */
    float    ms_stage[8];           /* per vkp_stage, valid when profiling is on           */
    float    ms_tlas;
    float    ms_frame;
    uint32_t kernel_launches;       /* kernels launched by the last vkp_render             */
    uint32_t _pad;
    uint64_t node_fetches;          /* traversal: child-pair fetches (64 B each)            */
    uint64_t triangle_tests;        /* traversal: leaf triangles fetched + tested (48 B)    */
    uint64_t instance_entries;      /* traversal: TLAS leaves entered (80 B record each)    */
    uint32_t worst_thread_nodes;    /* most child-pair fetches any one thread made (the kernel-tail length) ...            */
    uint32_t worst_thread_tag;      /* ... and who: bits[31:28] stage, [27:23] light, [22:0] pixel index y*W+x            */
} vkp_counters;

/* acceleration-structure build / refit figures (vkp_get_build_stats): what vkCmdBuildAccelerationStructuresKHR cost the
 * reference's driver (renderer.cpp:1747 buildBlas, :1853 buildTlas, per-frame MODE_UPDATE refit vk_engine.cpp:172-187) */
typedef struct vkp_build_stats {
    float    blas_ms;               /* CUDA-event time of the last vkp_build_blas: every BLAS it built, queued back to back (memory allocated before) */
    float    tlas_ms;               /* last full TLAS build */
    float    refit_ms;              /* last TLAS refit (instance boxes + bottom-up fit + node emit), on the update stream */
    uint32_t blas_built;            /* BLAS built by the last vkp_build_blas */
    uint32_t blas_launches;         /* kernels it launched */
    uint32_t blas_host_syncs;       /* stream synchronisations the host needed inside the build (0 = fully device-driven) */
    uint32_t blas_height_max;       /* tallest BLAS of the scene */
    uint32_t tlas_height;
    uint32_t refit_launches;
    uint64_t blas_triangles;        /* triangles of the BLAS built by the last vkp_build_blas (not instanced) */
    uint64_t arena_bytes;           /* nodes + leaf triangles of every BLAS, one contiguous allocation (the L2 persistence window) */
    uint64_t l2_persist_bytes;      /* bytes of the arena set aside as persisting L2 lines (0 = window off) */
    float    l2_hit_ratio;          /* hitRatio of the access-policy window */
    uint16_t blas_ploc_rounds_max;  /* most PLOC rounds any BLAS of the last vkp_build_blas needed */
    uint16_t blas_rebuilt;          /* BLAS whose queued rounds did not suffice and which were built again with host waits (0 = none) */
    float    blas_wall_ms;          /* host wall clock of the whole vkp_build_blas call: allocations + queueing + the one wait + compaction + frees */
    float    blas_compact_ms;       /* CUDA-event time of the copies into the arena (part of blas_wall_ms, not of blas_ms) */
} vkp_build_stats;

/* 32-byte BVH node (both BLAS and TLAS).  ref: bit31 = leaf; leaf: bits[30:3] = first item,
 * bits[2:0] = count-1 ... see DESIGN.md "BVH layout". */
typedef struct vkp_bvh_node {
    float    bmin[3];
    uint32_t ref;
    float    bmax[3];
    uint32_t aux;
} vkp_bvh_node;

#ifdef __cplusplus
}
#endif
#endif /* VKP_TYPES_H */
