/*
 * oracle.h -- TEST INFRASTRUCTURE.  C interface of the CPU oracle (liboracle.so).
 *
 * A scalar restatement of the reference's Hybrid / Raytraced / Deferred shader math
 * (data/shaders/{x}.rgen|rchit|rmiss|comp|frag, see oracle.cpp for file:line citations) over a
 * brute-force or simple binned-SAH CPU BVH.
 *
 * PINNING.  The reference ships no tests, golden images or known-answer vectors
 * and cannot be built or run here; it does ship its compiled shaders.  Every SHADING stage of this oracle is
 * pinned against outputs of those binaries executed by oracle/spirv_interp.py (tests/golden/spirv_*.npz,
 * tests/test_oracle_spirv.py: integer results identical, float results within 2e-5 relative, stored bytes
 * identical).  The INTERSECTION side is unpinned by nature -- the reference's BVH build / traversal /
 * ray-triangle test live in the Vulkan driver and RT hardware; its pins are brute-force == BVH agreement
 * and the mathematical definition of nearest-hit with the tie rule below.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may load this.
 *
 * Intersection contract (shared with the CUDA path, stated in DESIGN.md):
 *   Moeller-Trumbore in fp32 in OBJECT space (ray transformed by inv_model, t preserved), no
 *   culling, accept tmin < t < tmax, u >= 0, v >= 0, u+v <= 1; closest = min t, ties -> lower
 *   instance id, then lower primitive id.  Any-hit results are booleans.
 */
#ifndef ORACLE_H
#define ORACLE_H

#include "../include/vkp_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_mesh { const vkp_vertex* vertices; uint32_t n_vertices; const uint32_t* indices; uint32_t n_indices; } orc_mesh;
typedef struct orc_texture { const uint8_t* rgba; int32_t width, height; } orc_texture;

typedef struct orc_scene_desc {
    const orc_mesh*     meshes;     int32_t n_meshes;
    const orc_texture*  textures;   int32_t n_textures;
    int32_t sky_tex, irradiance_tex;
    const vkp_material* materials;  int32_t n_materials;
    const vkp_light*    lights;     int32_t n_lights;
    const vkp_instance* instances;  int32_t n_instances;
} orc_scene_desc;

typedef struct orc_scene orc_scene;

/* G-buffer planes, same binary layout as enum vkp_plane */
typedef struct orc_gbuffer {
    uint16_t* position;  /* W*H*4 fp16 */
    uint16_t* normal;    /* W*H*4 fp16 */
    uint8_t*  albedo;    /* W*H*4 */
    uint16_t* motion;    /* W*H*2 fp16 */
    uint8_t*  material;  /* W*H*4 */
    uint8_t*  emissive;  /* W*H*4 */
    int32_t*  hit_ids;   /* W*H*2 (may be NULL) */
} orc_gbuffer;

/* rays[0]=primary(G-buffer) [1]=shadow [2]=reflect [3]=secondary shadow */
typedef struct orc_counters { uint64_t rays[4]; } orc_counters;

int  orc_abi_version(void);
int  orc_num_threads(void);
void orc_set_num_threads(int n);
int  orc_uses_libm(void);
/* bench timing aid: passes only process rows y % stride == offset (results of skipped rows are left untouched) */
void orc_set_row_sampling(int stride, int offset);

orc_scene* orc_scene_create(const orc_scene_desc* desc);   /* deep-copies everything, builds BVHs */
void orc_scene_destroy(orc_scene* s);
void orc_scene_set_lights(orc_scene* s, const vkp_light* lights, int n);
void orc_scene_set_instances(orc_scene* s, const vkp_instance* instances, int n);
uint64_t orc_scene_triangles(const orc_scene* s);           /* instanced triangle count */

/* traceRayEXT: brute_force != 0 tests every triangle of every instance */
void orc_trace_rays(const orc_scene* s, const vkp_ray* rays, uint32_t n, int any_hit, int brute_force, vkp_hit* hits);

/* stages; every stage reads/writes caller-owned planes so they can be isolated */
void orc_gbuffer_pass(const orc_scene* s, const vkp_camera* cam, const vkp_params* p, orc_gbuffer* out, orc_counters* c);
void orc_shadow_pass(const orc_scene* s, const vkp_camera* cam, const vkp_params* p, const orc_gbuffer* g,
                     uint8_t* shadow /* n_lights planes of W*H */, orc_counters* c);
void orc_denoise_pass(const vkp_params* p, int n_lights, const orc_gbuffer* g, const uint8_t* shadow,
                      uint8_t* denoised /* in: history, out: new history */);
void orc_resolve_pass(const orc_scene* s, const vkp_camera* cam, const vkp_params* p, const orc_gbuffer* g,
                      const uint8_t* denoised, float* final_rgba32f, orc_counters* c);
void orc_post_pass(const vkp_params* p, const float* final_rgba32f, uint8_t* srgb8);
void orc_raytrace_pass(const orc_scene* s, const vkp_camera* cam, const vkp_params* p,
                       float* accum_rgba32f /* in: previous frame, out */, orc_counters* c);
void orc_deferred_pass(const orc_scene* s, const vkp_camera* cam, const vkp_params* p, const orc_gbuffer* g,
                       float* final_rgba32f);

/* scalar known-answer helpers (tests) */
uint32_t orc_tea(uint32_t v0, uint32_t v1);
uint32_t orc_lcg(uint32_t* state);
float    orc_rnd(uint32_t* state);
void     orc_sample_disk(const vkp_light* light, const float pos[3], const float L[3], uint32_t seed, float out_dir[3]);
float    orc_distribution_ggx(const float N[3], const float H[3], float roughness);
float    orc_geometry_smith(const float N[3], const float V[3], const float L[3], float roughness);
void     orc_fresnel_schlick(float cos_theta, const float F0[3], float out[3]);
void     orc_env_uv(const float dir[3], int use_miss_pi, float out_uv[2]);
uint16_t orc_to_half(float f);
float    orc_from_half(uint16_t h);
uint8_t  orc_to_unorm8(float f);

#ifdef __cplusplus
}
#endif
#endif
