"""ctypes / numpy mirrors of include/vkp_types.h (byte-exact; checked by tests/test_abi.py)."""
import ctypes as C

import numpy as np


class Vertex(C.Structure):          # vkp_vertex, 44 B
    _fields_ = [("position", C.c_float * 3), ("normal", C.c_float * 3), ("color", C.c_float * 3), ("uv", C.c_float * 2)]


class Material(C.Structure):        # vkp_material, 48 B
    _fields_ = [("diffuse", C.c_float * 4), ("textures", C.c_float * 4), ("shadingMetallicRoughness", C.c_float * 4)]


class Light(C.Structure):           # vkp_light, 48 B
    _fields_ = [("position", C.c_float * 4), ("color", C.c_float * 4), ("radius", C.c_float), ("_pad", C.c_float * 3)]


class Instance(C.Structure):        # vkp_instance, 144 B
    _fields_ = [("mesh_id", C.c_int32), ("first_index", C.c_uint32), ("index_count", C.c_uint32), ("material_id", C.c_int32),
                ("model", C.c_float * 16), ("inv_model", C.c_float * 16)]


class Camera(C.Structure):          # vkp_camera, 400 B
    _fields_ = [("view", C.c_float * 16), ("proj", C.c_float * 16), ("prev_view", C.c_float * 16), ("prev_proj", C.c_float * 16),
                ("inv_view", C.c_float * 16), ("inv_proj", C.c_float * 16), ("position", C.c_float * 4)]


class Params(C.Structure):          # vkp_params, 64 B
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("mode", C.c_int32), ("spp", C.c_int32), ("shadow_samples", C.c_int32),
                ("max_bounces", C.c_int32), ("frame", C.c_int32), ("denoise_frame", C.c_int32), ("faithful_rng", C.c_int32),
                ("denoise_cover_remainder", C.c_int32), ("debug_target", C.c_int32), ("wavefront", C.c_int32), ("fuse_denoise", C.c_int32), ("reproject_fix", C.c_int32), ("denoiser", C.c_int32), ("reserved", C.c_int32 * 1)]


class Ray(C.Structure):             # vkp_ray, 32 B
    _fields_ = [("origin", C.c_float * 3), ("tmin", C.c_float), ("dir", C.c_float * 3), ("tmax", C.c_float)]


class Hit(C.Structure):             # vkp_hit, 20 B
    _fields_ = [("t", C.c_float), ("u", C.c_float), ("v", C.c_float), ("instance", C.c_int32), ("primitive", C.c_int32)]


class Counters(C.Structure):        # vkp_counters
    _fields_ = [("rays_primary", C.c_uint64), ("rays_shadow", C.c_uint64), ("rays_reflect", C.c_uint64), ("rays_secondary_shadow", C.c_uint64),
                ("ms_stage", C.c_float * 8), ("ms_tlas", C.c_float), ("ms_frame", C.c_float), ("kernel_launches", C.c_uint32), ("_pad", C.c_uint32),
                ("node_fetches", C.c_uint64), ("triangle_tests", C.c_uint64), ("instance_entries", C.c_uint64),
                ("worst_thread_nodes", C.c_uint32), ("worst_thread_tag", C.c_uint32)]


class BuildStats(C.Structure):      # vkp_build_stats
    _fields_ = [("blas_ms", C.c_float), ("tlas_ms", C.c_float), ("refit_ms", C.c_float), ("blas_built", C.c_uint32), ("blas_launches", C.c_uint32),
                ("blas_host_syncs", C.c_uint32), ("blas_height_max", C.c_uint32), ("tlas_height", C.c_uint32), ("refit_launches", C.c_uint32),
                ("blas_triangles", C.c_uint64), ("arena_bytes", C.c_uint64), ("l2_persist_bytes", C.c_uint64), ("l2_hit_ratio", C.c_float), ("blas_ploc_rounds_max", C.c_uint16), ("blas_rebuilt", C.c_uint16),
                ("blas_wall_ms", C.c_float), ("blas_compact_ms", C.c_float)]


class BvhNode(C.Structure):         # vkp_bvh_node, 32 B
    _fields_ = [("bmin", C.c_float * 3), ("ref", C.c_uint32), ("bmax", C.c_float * 3), ("aux", C.c_uint32)]


SIZES = {Vertex: 44, Material: 48, Light: 48, Instance: 144, Camera: 400, Params: 64, Ray: 32, Hit: 20, BvhNode: 32, BuildStats: 80}

vertex_dtype = np.dtype([("position", "<f4", 3), ("normal", "<f4", 3), ("color", "<f4", 3), ("uv", "<f4", 2)])
material_dtype = np.dtype([("diffuse", "<f4", 4), ("textures", "<f4", 4), ("shadingMetallicRoughness", "<f4", 4)])
light_dtype = np.dtype([("position", "<f4", 4), ("color", "<f4", 4), ("radius", "<f4"), ("_pad", "<f4", 3)])
instance_dtype = np.dtype([("mesh_id", "<i4"), ("first_index", "<u4"), ("index_count", "<u4"), ("material_id", "<i4"),
                           ("model", "<f4", 16), ("inv_model", "<f4", 16)])
ray_dtype = np.dtype([("origin", "<f4", 3), ("tmin", "<f4"), ("dir", "<f4", 3), ("tmax", "<f4")])
hit_dtype = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("instance", "<i4"), ("primitive", "<i4")])
bvh_node_dtype = np.dtype([("bmin", "<f4", 3), ("ref", "<u4"), ("bmax", "<f4", 3), ("aux", "<u4")])

assert vertex_dtype.itemsize == 44 and material_dtype.itemsize == 48 and light_dtype.itemsize == 48
assert instance_dtype.itemsize == 144 and ray_dtype.itemsize == 32 and hit_dtype.itemsize == 20 and bvh_node_dtype.itemsize == 32
