"""ctypes binding of the C-ABI (include/vkp.h) and of the host shim (host/host_capi.cpp).

Nothing here computes: every call forwards to libvkp.so / libvkp_host.so.  A missing library raises
``VkpError`` at load time -- there is no Python fallback.
"""
import ctypes as C
import os

import numpy as np

from . import types as T

MODE_DEFERRED, MODE_RAYTRACING, MODE_HYBRID = 0, 1, 2


class PLANE:
    POSITION, NORMAL, ALBEDO, MOTION, MATERIAL, EMISSIVE, HIT_IDS, FINAL_F32, FINAL_SRGB8 = range(9)
    SHADOW0, DENOISED0 = 16, 48


class STAGE:
    GBUFFER, SHADOW, DENOISE, RESOLVE, POST, RAYTRACE, DEFERRED = range(7)


# plane -> (numpy dtype, channels)
_PLANE_FMT = {
    PLANE.POSITION: (np.float16, 4), PLANE.NORMAL: (np.float16, 4), PLANE.ALBEDO: (np.uint8, 4), PLANE.MOTION: (np.float16, 2),
    PLANE.MATERIAL: (np.uint8, 4), PLANE.EMISSIVE: (np.uint8, 4), PLANE.HIT_IDS: (np.int32, 2), PLANE.FINAL_F32: (np.float32, 4),
    PLANE.FINAL_SRGB8: (np.uint8, 4),
}


def plane_format(plane):
    if plane in _PLANE_FMT:
        return _PLANE_FMT[plane]
    if PLANE.SHADOW0 <= plane < PLANE.SHADOW0 + 32 or PLANE.DENOISED0 <= plane < PLANE.DENOISED0 + 32:
        return (np.uint8, 1)
    raise ValueError(f"unknown plane {plane}")


class VkpError(RuntimeError):
    pass


def lib_dir():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")


_vkp = None
_host = None


def load_vkp():
    """Load libvkp.so and declare every prototype of include/vkp.h."""
    global _vkp
    if _vkp is not None:
        return _vkp
    path = os.environ.get("VKP_LIB") or os.path.join(lib_dir(), "libvkp.so")      # VKP_LIB: tuning builds (scripts/variants.sh)
    if not os.path.exists(path):
        raise VkpError(f"{path} is missing: run `make` (or __graft_entry__.build()); there is no fallback path")
    L = C.CDLL(path, mode=C.RTLD_GLOBAL)
    vp, i, u32, sz = C.c_void_p, C.c_int, C.c_uint32, C.c_size_t
    P = C.POINTER
    sig = {
        "vkp_abi_version": (i, []),
        "vkp_last_error": (C.c_char_p, [vp]),
        "vkp_create": (i, [P(vp), i]),
        "vkp_destroy": (None, [vp]),
        "vkp_upload_mesh": (i, [vp, vp, u32, vp, u32, P(i)]),
        "vkp_upload_texture_rgba8": (i, [vp, vp, i, i, P(i)]),
        "vkp_set_env": (i, [vp, i, i]),
        "vkp_set_materials": (i, [vp, vp, i]),
        "vkp_set_lights": (i, [vp, vp, i]),
        "vkp_set_instances": (i, [vp, vp, i]),
        "vkp_build_blas": (i, [vp]),
        "vkp_set_builder": (i, [vp, i]),
        "vkp_build_tlas": (i, [vp, i]),
        "vkp_get_build_stats": (i, [vp, P(T.BuildStats)]),
        "vkp_set_l2_persistence": (i, [vp, i]),
        "vkp_set_camera": (i, [vp, P(T.Camera)]),
        "vkp_set_params": (i, [vp, P(T.Params)]),
        "vkp_set_tiling": (i, [vp, i, i, i, i]),
        "vkp_render": (i, [vp]),
        "vkp_run_stage": (i, [vp, i]),
        "vkp_sync": (i, [vp]),
        "vkp_reset_history": (i, [vp]),
        "vkp_plane_bytes": (i, [vp, i, P(sz)]),
        "vkp_readback": (i, [vp, i, vp, sz]),
        "vkp_readback_async": (i, [vp, i, vp, sz]),
        "vkp_readback_wait": (i, [vp, i]),
        "vkp_upload_plane": (i, [vp, i, vp, sz]),
        "vkp_plane_device_ptr": (i, [vp, i, P(vp), P(sz)]),
        "vkp_get_counters": (i, [vp, P(T.Counters)]),
        "vkp_set_profiling": (i, [vp, i]),
        "vkp_set_counter_accumulation": (i, [vp, i]),
        "vkp_trace_rays": (i, [vp, vp, u32, i, vp]),
        "vkp_bvh_info": (i, [vp, i, P(u32), P(u32)]),
        "vkp_bvh_download": (i, [vp, i, vp, vp]),
        "vkp_timer_begin": (i, [vp]),
        "vkp_timer_end": (i, [vp, P(C.c_float)]),
        "vkp_flush_l2": (i, [vp]),
        "vkp_tiles_packed_bytes": (i, [vp, i, i, i, P(sz)]),
        "vkp_tiles_pack": (i, [vp, i, vp, sz]),
        "vkp_tiles_unpack": (i, [vp, i, i, i, vp, sz]),
        "vkp_set_stream": (i, [vp, vp]),
        "vkp_eval_math": (i, [vp, i, vp, vp, u32, vp]),
        "vkp_trace_stats": (i, [vp, vp, u32, i, vp]),
        "vkp_selftest_sort": (i, [vp, vp, u32, i, vp, vp, vp]),
        "vkp_ipc_export_plane": (i, [vp, i, vp]),
        "vkp_ipc_open": (i, [vp, vp, P(vp)]),
        "vkp_set_output_redirect": (i, [vp, vp, vp]),
        "vkp_set_output_buffer": (i, [vp, i]),
        "vkp_sync_export": (i, [vp, vp]),
        "vkp_sync_attach": (i, [vp, vp]),
        "vkp_sync_arrive": (i, [vp, C.c_uint32]),
        "vkp_sync_wait_arrivals": (i, [vp, C.c_uint32]),
        "vkp_sync_release": (i, [vp, C.c_uint32, i]),
        "vkp_sync_wait_release": (i, [vp, C.c_uint32]),
        "vkp_sync_timed_out": (i, [vp, P(i)]),
        "vkp_group_create": (i, [P(vp), P(i), i, i, i]),
        "vkp_group_destroy": (None, [vp]),
        "vkp_group_size": (i, [vp]),
        "vkp_group_member": (vp, [vp, i]),
        "vkp_group_last_error": (C.c_char_p, [vp]),
        "vkp_group_upload_mesh": (i, [vp, vp, u32, vp, u32, P(i)]),
        "vkp_group_upload_texture_rgba8": (i, [vp, vp, i, i, P(i)]),
        "vkp_group_set_env": (i, [vp, i, i]),
        "vkp_group_set_materials": (i, [vp, vp, i]),
        "vkp_group_set_lights": (i, [vp, vp, i]),
        "vkp_group_set_instances": (i, [vp, vp, i]),
        "vkp_group_build_blas": (i, [vp]),
        "vkp_group_build_tlas": (i, [vp, i]),
        "vkp_group_set_camera": (i, [vp, P(T.Camera)]),
        "vkp_group_set_params": (i, [vp, P(T.Params)]),
        "vkp_group_reset_history": (i, [vp]),
        "vkp_group_render": (i, [vp]),
        "vkp_group_sync": (i, [vp]),
        "vkp_group_readback": (i, [vp, i, vp, sz]),
        "vkp_group_get_counters": (i, [vp, P(T.Counters)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    L._signatures = sig
    _vkp = L
    return L


_scene = None


def load_scene():
    """Load libvkp_scene.so: the host scene model and the image readers / writers.  It depends on neither CUDA nor libvkp.so, so a
    process that only needs scenes (bench.py --impl reference) never maps the GPU library."""
    global _scene
    if _scene is not None:
        return _scene
    path = os.path.join(lib_dir(), "libvkp_scene.so")
    if not os.path.exists(path):
        raise VkpError(f"{path} is missing: run `make`")
    S = C.CDLL(path, mode=C.RTLD_GLOBAL)
    vp, i, u32 = C.c_void_p, C.c_int, C.c_uint32
    P = C.POINTER
    sig = {
        "vkph_last_error": (C.c_char_p, []),
        "vkph_scene_create": (vp, [i, P(i)]),
        "vkph_scene_create_prefab": (vp, [C.c_char_p, P(i)]),
        "vkph_scene_destroy": (None, [vp]),
        "vkph_scene_counts": (i, [vp] + [P(i)] * 7),
        "vkph_scene_mesh": (i, [vp, i, P(vp), P(u32), P(vp), P(u32)]),
        "vkph_scene_instances": (vp, [vp]),
        "vkph_scene_materials": (vp, [vp]),
        "vkph_scene_lights": (vp, [vp]),
        "vkph_scene_texture": (i, [vp, i, P(vp), P(i), P(i)]),
        "vkph_scene_camera": (i, [vp, C.c_float, P(T.Camera), P(T.Camera)]),
        "vkph_scene_camera_move": (i, [vp, i, C.c_float]),
        "vkph_scene_camera_rotate": (i, [vp, C.c_float, C.c_float]),
        "vkph_scene_camera_set": (i, [vp, P(C.c_float * 3), C.c_float, C.c_float]),
        "vkph_scene_set_entity_matrix": (i, [vp, i, P(C.c_float * 16)]),
        "vkph_write_png": (i, [C.c_char_p, vp, i, i]),
        "vkph_write_exr": (i, [C.c_char_p, vp, i, i]),
        "vkph_decode_image_file": (i, [C.c_char_p, P(i), P(i), vp, C.c_size_t]),
    }
    for name, (res, args) in sig.items():
        f = getattr(S, name)
        f.restype, f.argtypes = res, args
    _scene = S
    return S


def load_host():
    """Load libvkp_host.so (host Renderer / VulkanEngine over libvkp.so + libvkp_scene.so, and its C shim)."""
    global _host
    if _host is not None:
        return _host
    load_vkp()
    load_scene()
    path = os.path.join(lib_dir(), "libvkp_host.so")
    if not os.path.exists(path):
        raise VkpError(f"{path} is missing: run `make`")
    H = C.CDLL(path)
    vp, i, u32 = C.c_void_p, C.c_int, C.c_uint32
    P = C.POINTER
    sig = {
        "vkph_engine_create": (vp, [i, i, i, i, P(i)]),
        "vkph_engine_create_multi": (vp, [i, i, i, i, i, P(i)]),
        "vkph_engine_set": (i, [vp, C.c_char_p, i]),
        "vkph_engine_frames": (i, [vp, i]),
        "vkph_engine_camera_move": (i, [vp, i, C.c_float]),
        "vkph_engine_read": (i, [vp, i, vp, C.c_size_t]),
        "vkph_engine_present": (i, [vp, vp, C.c_size_t]),
        "vkph_engine_present_wait": (i, [vp, i]),
        "vkph_engine_counters": (i, [vp, P(T.Counters)]),
        "vkph_engine_state": (i, [vp, P(i), P(i)]),
        "vkph_engine_ctx": (vp, [vp]),
        "vkph_engine_destroy": (None, [vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(H, name)
        f.restype, f.argtypes = res, args
    H.vkph_last_error = load_scene().vkph_last_error
    _host = H
    return H


def decode_image(path):
    """Texture::GET's image loader (host/image_decode.cpp, the stbi_load(..., 4) replacement): H x W x 4 uint8."""
    H = load_scene()
    w, h = C.c_int(), C.c_int()
    if H.vkph_decode_image_file(str(path).encode(), C.byref(w), C.byref(h), None, 0) != 0:
        raise VkpError(H.vkph_last_error().decode())
    out = np.empty((h.value, w.value, 4), dtype=np.uint8)
    if H.vkph_decode_image_file(str(path).encode(), C.byref(w), C.byref(h), out.ctypes.data_as(C.c_void_p), out.nbytes) != 0:
        raise VkpError(H.vkph_last_error().decode())
    return out


def make_params(width, height, mode=MODE_HYBRID, spp=1, shadow_samples=1, max_bounces=8, frame=0, denoise_frame=None,
                faithful_rng=1, denoise_cover_remainder=1, debug_target=0, wavefront=0, fuse_denoise=0, reproject_fix=0, denoiser=0):
    p = T.Params()
    p.width, p.height, p.mode, p.spp, p.shadow_samples, p.max_bounces = width, height, mode, spp, shadow_samples, max_bounces
    p.frame = frame
    p.denoise_frame = frame if denoise_frame is None else denoise_frame
    p.faithful_rng, p.denoise_cover_remainder, p.debug_target = faithful_rng, denoise_cover_remainder, debug_target
    p.wavefront = wavefront
    p.fuse_denoise = fuse_denoise
    p.reproject_fix = reproject_fix
    p.denoiser = denoiser
    return p


def _detail4(detail):
    """detail = (lucy subdivisions, helmet subdivisions, stress subdivisions, stress instances) of the procedural stand-in meshes,
    or None for the full-size scene.  It is an argument of every scene / engine creation, never library state."""
    return None if detail is None else (C.c_int * 4)(*[int(x) for x in detail])


class HostScene:
    """Scene::create_scene(i) of the host library, exported as the flat arrays both the C-ABI and the oracle consume."""

    def __init__(self, index=0, detail=None):
        H = load_scene()
        self._H = H
        self.detail = None if detail is None else tuple(int(x) for x in detail)
        self._h = H.vkph_scene_create(index, _detail4(detail))
        if not self._h:
            raise VkpError(H.vkph_last_error().decode())
        self.index = index
        self._export()

    @classmethod
    def from_prefab(cls, path, detail=None):
        """Scene::create_viewer_scene: the default scene's floor, lights and camera around one .obj / .gltf / .glb file."""
        H = load_scene()
        self = cls.__new__(cls)
        self._H = H
        self.detail = None if detail is None else tuple(int(x) for x in detail)
        self._h = H.vkph_scene_create_prefab(str(path).encode(), _detail4(detail))
        if not self._h:
            raise VkpError(H.vkph_last_error().decode())
        self.index = -1
        self._export()
        return self

    def _export(self):
        H, h = self._H, self._h
        n = [C.c_int() for _ in range(7)]
        H.vkph_scene_counts(h, *[C.byref(x) for x in n])
        n_meshes, n_inst, n_mat, n_lights, n_tex, sky, irr = [x.value for x in n]
        self.meshes = []
        for m in range(n_meshes):
            v, nv, ix, ni = C.c_void_p(), C.c_uint32(), C.c_void_p(), C.c_uint32()
            H.vkph_scene_mesh(h, m, C.byref(v), C.byref(nv), C.byref(ix), C.byref(ni))
            verts = np.frombuffer((C.c_char * (44 * nv.value)).from_address(v.value), dtype=T.vertex_dtype).copy()
            idx = np.frombuffer((C.c_char * (4 * ni.value)).from_address(ix.value), dtype=np.uint32).copy()
            self.meshes.append((verts, idx))

        def arr(ptr, count, dtype):
            if count == 0:
                return np.zeros(0, dtype=dtype)
            return np.frombuffer((C.c_char * (dtype.itemsize * count)).from_address(ptr), dtype=dtype).copy()

        self.instances = arr(H.vkph_scene_instances(h), n_inst, T.instance_dtype)
        self.materials = arr(H.vkph_scene_materials(h), n_mat, T.material_dtype)
        self.lights = arr(H.vkph_scene_lights(h), n_lights, T.light_dtype)
        self.textures = []
        for t in range(n_tex):
            p, w, hh = C.c_void_p(), C.c_int(), C.c_int()
            H.vkph_scene_texture(h, t, C.byref(p), C.byref(w), C.byref(hh))
            self.textures.append(np.frombuffer((C.c_char * (4 * w.value * hh.value)).from_address(p.value), dtype=np.uint8)
                                 .reshape(hh.value, w.value, 4).copy())
        self.sky_tex, self.irradiance_tex = sky, irr

    @property
    def n_triangles(self):
        return int(sum(int(i["index_count"]) // 3 for i in self.instances))

    def camera(self, width, height, prev=None):
        cam = T.Camera()
        self._H.vkph_scene_camera(self._h, C.c_float(width / height), C.byref(prev) if prev is not None else None, C.byref(cam))
        return cam

    def move_camera(self, direction, dt):
        self._H.vkph_scene_camera_move(self._h, direction, C.c_float(dt))

    def rotate_camera(self, dx, dy):
        self._H.vkph_scene_camera_rotate(self._h, C.c_float(dx), C.c_float(dy))

    def set_camera(self, pos, yaw=-90.0, pitch=0.0):
        self._H.vkph_scene_camera_set(self._h, C.byref((C.c_float * 3)(*pos)), C.c_float(yaw), C.c_float(pitch))

    def set_entity_matrix(self, entity, m16):
        r = self._H.vkph_scene_set_entity_matrix(self._h, entity, C.byref((C.c_float * 16)(*[float(x) for x in m16])))
        if r != 0:
            raise VkpError("bad entity index")
        self._export()

    def close(self):
        if self._h:
            self._H.vkph_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One vkp_ctx (one GPU).  Thin: every method is one C-ABI call plus numpy marshalling."""

    def __init__(self, device=0):
        self.L = load_vkp()
        h = C.c_void_p()
        r = self.L.vkp_create(C.byref(h), device)
        if r != 0:
            raise VkpError(f"vkp_create failed ({r}): {self.L.vkp_last_error(None).decode()}")
        self.h = h
        self.width = self.height = 0
        self.n_lights = 0
        self._keep = []

    def _ck(self, r, what):
        if r != 0:
            raise VkpError(f"{what} failed ({r}): {self.L.vkp_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.L.vkp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- scene --------------------------------------------------------------------------------------
    def upload_mesh(self, verts, idx):
        verts = np.ascontiguousarray(verts, dtype=T.vertex_dtype)
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        mid = C.c_int(-1)
        self._ck(self.L.vkp_upload_mesh(self.h, verts.ctypes.data, len(verts), idx.ctypes.data, len(idx), C.byref(mid)), "vkp_upload_mesh")
        return mid.value

    def upload_texture(self, rgba):
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        tid = C.c_int(-1)
        self._ck(self.L.vkp_upload_texture_rgba8(self.h, rgba.ctypes.data, rgba.shape[1], rgba.shape[0], C.byref(tid)), "vkp_upload_texture_rgba8")
        return tid.value

    def set_env(self, sky, irr):
        self._ck(self.L.vkp_set_env(self.h, sky, irr), "vkp_set_env")

    def set_materials(self, mats):
        mats = np.ascontiguousarray(mats, dtype=T.material_dtype)
        self._ck(self.L.vkp_set_materials(self.h, mats.ctypes.data, len(mats)), "vkp_set_materials")

    def set_lights(self, lights):
        lights = np.ascontiguousarray(lights, dtype=T.light_dtype)
        self.n_lights = len(lights)
        self._ck(self.L.vkp_set_lights(self.h, lights.ctypes.data if len(lights) else None, len(lights)), "vkp_set_lights")

    def set_instances(self, inst):
        inst = np.ascontiguousarray(inst, dtype=T.instance_dtype)
        self._ck(self.L.vkp_set_instances(self.h, inst.ctypes.data if len(inst) else None, len(inst)), "vkp_set_instances")

    def set_builder(self, builder):
        self._ck(self.L.vkp_set_builder(self.h, builder), "vkp_set_builder")

    def build_blas(self):
        self._ck(self.L.vkp_build_blas(self.h), "vkp_build_blas")

    def build_tlas(self, update=False):
        self._ck(self.L.vkp_build_tlas(self.h, 1 if update else 0), "vkp_build_tlas")

    def build_stats(self):
        """vkp_get_build_stats as a dict (build / refit milliseconds, heights, arena and L2 window)."""
        b = T.BuildStats()
        self._ck(self.L.vkp_get_build_stats(self.h, C.byref(b)), "vkp_get_build_stats")
        return {k: getattr(b, k) for k, _ in T.BuildStats._fields_ if k != "_pad"}

    def set_l2_persistence(self, on):
        self._ck(self.L.vkp_set_l2_persistence(self.h, int(on)), "vkp_set_l2_persistence")

    def upload_scene(self, scene: HostScene):
        """Renderer::load_data_to_gpu + create_*_acceleration_structure in C-ABI calls."""
        for verts, idx in scene.meshes:
            self.upload_mesh(verts, idx)
        for tex in scene.textures:
            self.upload_texture(tex)
        self.set_env(scene.sky_tex, scene.irradiance_tex)
        self.set_materials(scene.materials)
        self.set_lights(scene.lights)
        self.set_instances(scene.instances)
        self.build_blas()
        self.build_tlas(False)

    # ---- frame --------------------------------------------------------------------------------------
    def set_camera(self, cam):
        self._ck(self.L.vkp_set_camera(self.h, C.byref(cam)), "vkp_set_camera")

    def set_params(self, p):
        self.width, self.height = p.width, p.height
        self._ck(self.L.vkp_set_params(self.h, C.byref(p)), "vkp_set_params")

    def set_tiling(self, rank, world, tile_w=16, tile_h=8):
        self._ck(self.L.vkp_set_tiling(self.h, rank, world, tile_w, tile_h), "vkp_set_tiling")

    def render(self):
        self._ck(self.L.vkp_render(self.h), "vkp_render")

    def run_stage(self, stage):
        self._ck(self.L.vkp_run_stage(self.h, stage), "vkp_run_stage")

    def sync(self):
        self._ck(self.L.vkp_sync(self.h), "vkp_sync")

    def reset_history(self):
        self._ck(self.L.vkp_reset_history(self.h), "vkp_reset_history")

    def set_profiling(self, on):
        self._ck(self.L.vkp_set_profiling(self.h, int(on)), "vkp_set_profiling")

    def set_counter_accumulation(self, on):
        self._ck(self.L.vkp_set_counter_accumulation(self.h, 1 if on else 0), "vkp_set_counter_accumulation")

    def set_stream(self, stream_ptr):
        self._ck(self.L.vkp_set_stream(self.h, C.c_void_p(stream_ptr)), "vkp_set_stream")

    def readback(self, plane, out=None):
        dt, ch = plane_format(plane)
        shape = (self.height, self.width, ch) if ch > 1 else (self.height, self.width)
        if out is None:
            out = np.empty(shape, dtype=dt)
        self._ck(self.L.vkp_readback(self.h, plane, out.ctypes.data, out.nbytes), "vkp_readback")
        return out

    def readback_async(self, plane, dst_ptr, nbytes):
        """Queue the D2H of `plane` into pinned host memory behind the frames submitted so far (frames in flight)."""
        self._ck(self.L.vkp_readback_async(self.h, plane, C.c_void_p(dst_ptr), nbytes), "vkp_readback_async")

    def readback_wait(self, keep_in_flight=0):
        self._ck(self.L.vkp_readback_wait(self.h, keep_in_flight), "vkp_readback_wait")

    def upload_plane(self, plane, arr):
        dt, _ = plane_format(plane)
        arr = np.ascontiguousarray(arr, dtype=dt)
        self._ck(self.L.vkp_upload_plane(self.h, plane, arr.ctypes.data, arr.nbytes), "vkp_upload_plane")

    def plane_device_ptr(self, plane):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.L.vkp_plane_device_ptr(self.h, plane, C.byref(p), C.byref(n)), "vkp_plane_device_ptr")
        return p.value, n.value

    def counters(self):
        c = T.Counters()
        self._ck(self.L.vkp_get_counters(self.h, C.byref(c)), "vkp_get_counters")
        return c

    def trace_rays(self, rays, any_hit=False):
        rays = np.ascontiguousarray(rays, dtype=T.ray_dtype)
        hits = np.empty(len(rays), dtype=T.hit_dtype)
        self._ck(self.L.vkp_trace_rays(self.h, rays.ctypes.data if len(rays) else None, len(rays), 1 if any_hit else 0,
                                       hits.ctypes.data if len(rays) else None), "vkp_trace_rays")
        return hits

    def selftest_sort(self, keys, end_bit=32):
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        sk = np.empty_like(keys); sv = np.empty_like(keys); sc = np.empty(len(keys), dtype=np.uint64)
        self._ck(self.L.vkp_selftest_sort(self.h, keys.ctypes.data, len(keys), end_bit, sk.ctypes.data, sv.ctypes.data, sc.ctypes.data), "vkp_selftest_sort")
        return sk, sv, sc

    def ipc_export_plane(self, plane):
        h = (C.c_ubyte * 64)()
        self._ck(self.L.vkp_ipc_export_plane(self.h, plane, h), "vkp_ipc_export_plane")
        return bytes(h)

    def ipc_open(self, handle):
        p = C.c_void_p()
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        self._ck(self.L.vkp_ipc_open(self.h, buf, C.byref(p)), "vkp_ipc_open")
        return p.value

    def set_output_redirect(self, ptr0, ptr1):
        self._ck(self.L.vkp_set_output_redirect(self.h, C.c_void_p(ptr0) if ptr0 else None, C.c_void_p(ptr1) if ptr1 else None), "vkp_set_output_redirect")

    def set_output_buffer(self, which):
        self._ck(self.L.vkp_set_output_buffer(self.h, which), "vkp_set_output_buffer")

    # frame barrier of the peer composite (device-side flags over NVLink)
    def sync_export(self):
        h = (C.c_ubyte * 64)()
        self._ck(self.L.vkp_sync_export(self.h, h), "vkp_sync_export")
        return bytes(h)

    def sync_attach(self, handle):
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        self._ck(self.L.vkp_sync_attach(self.h, buf), "vkp_sync_attach")

    def sync_arrive(self, frame):
        self._ck(self.L.vkp_sync_arrive(self.h, frame & 0xFFFFFFFF), "vkp_sync_arrive")

    def sync_wait_arrivals(self, frame):
        self._ck(self.L.vkp_sync_wait_arrivals(self.h, frame & 0xFFFFFFFF), "vkp_sync_wait_arrivals")

    def sync_release(self, frame, after_present=False):
        self._ck(self.L.vkp_sync_release(self.h, frame & 0xFFFFFFFF, 1 if after_present else 0), "vkp_sync_release")

    def sync_wait_release(self, frame):
        self._ck(self.L.vkp_sync_wait_release(self.h, frame & 0xFFFFFFFF), "vkp_sync_wait_release")

    def sync_timed_out(self):
        v = C.c_int(0)
        self._ck(self.L.vkp_sync_timed_out(self.h, C.byref(v)), "vkp_sync_timed_out")
        return bool(v.value)

    def trace_stats(self, rays, any_hit=False):
        rays = np.ascontiguousarray(rays, dtype=T.ray_dtype)
        out = np.zeros((len(rays), 3), dtype=np.uint32)
        self._ck(self.L.vkp_trace_stats(self.h, rays.ctypes.data, len(rays), 1 if any_hit else 0, out.ctypes.data), "vkp_trace_stats")
        return out

    def bvh(self, blas_index):
        nn, npr = C.c_uint32(), C.c_uint32()
        self._ck(self.L.vkp_bvh_info(self.h, blas_index, C.byref(nn), C.byref(npr)), "vkp_bvh_info")
        nodes = np.empty(nn.value, dtype=T.bvh_node_dtype)
        prims = np.empty(npr.value, dtype=np.uint32)
        self._ck(self.L.vkp_bvh_download(self.h, blas_index, nodes.ctypes.data, prims.ctypes.data if npr.value else None), "vkp_bvh_download")
        return nodes, prims

    def eval_math(self, func, x, y=None):
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.zeros_like(x) if y is None else np.ascontiguousarray(y, dtype=np.float32)
        out = np.empty_like(x)
        self._ck(self.L.vkp_eval_math(self.h, func, x.ctypes.data, y.ctypes.data, x.size, out.ctypes.data), "vkp_eval_math")
        return out

    def timer_begin(self):
        self._ck(self.L.vkp_timer_begin(self.h), "vkp_timer_begin")

    def timer_end(self):
        ms = C.c_float()
        self._ck(self.L.vkp_timer_end(self.h, C.byref(ms)), "vkp_timer_end")
        return ms.value

    def flush_l2(self):
        self._ck(self.L.vkp_flush_l2(self.h), "vkp_flush_l2")

    def tiles_packed_bytes(self, plane, rank, world):
        n = C.c_size_t()
        self._ck(self.L.vkp_tiles_packed_bytes(self.h, plane, rank, world, C.byref(n)), "vkp_tiles_packed_bytes")
        return n.value

    def tiles_pack(self, plane, device_dst, nbytes):
        self._ck(self.L.vkp_tiles_pack(self.h, plane, C.c_void_p(device_dst), nbytes), "vkp_tiles_pack")

    def tiles_unpack(self, plane, rank, world, device_src, nbytes):
        self._ck(self.L.vkp_tiles_unpack(self.h, plane, rank, world, C.c_void_p(device_src), nbytes), "vkp_tiles_unpack")


class HostEngine:
    """VulkanEngine of the host library (init / run(frames) / cleanup), driven through the shim."""

    def __init__(self, scene_index=0, width=1700, height=900, device=0, detail=None, gpus=1):
        H = load_host()
        self._H = H
        self.width, self.height = width, height
        self._h = H.vkph_engine_create_multi(scene_index, width, height, device, gpus, _detail4(detail))
        if not self._h:
            raise VkpError(H.vkph_last_error().decode())

    def set(self, key, value):
        if self._H.vkph_engine_set(self._h, key.encode(), int(value)) != 0:
            raise VkpError(f"unknown engine key {key}")

    def frames(self, n=1):
        if self._H.vkph_engine_frames(self._h, n) != 0:
            raise VkpError(self._H.vkph_last_error().decode())

    def move_camera(self, direction, dt):
        self._H.vkph_engine_camera_move(self._h, direction, C.c_float(dt))

    def present(self, dst_ptr, nbytes):
        """Queue the D2H of the frame just rendered into pinned memory at dst_ptr; the next frame renders into the other buffer."""
        if self._H.vkph_engine_present(self._h, C.c_void_p(dst_ptr), nbytes) != 0:
            raise VkpError(self._H.vkph_last_error().decode())

    def present_wait(self, keep_in_flight=0):
        if self._H.vkph_engine_present_wait(self._h, keep_in_flight) != 0:
            raise VkpError(self._H.vkph_last_error().decode())

    def read_srgb8(self):
        out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        if self._H.vkph_engine_read(self._h, 0, out.ctypes.data, out.nbytes) != 0:
            raise VkpError(self._H.vkph_last_error().decode())
        return out

    def read_linear(self):
        out = np.empty((self.height, self.width, 4), dtype=np.float32)
        if self._H.vkph_engine_read(self._h, 1, out.ctypes.data, out.nbytes) != 0:
            raise VkpError(self._H.vkph_last_error().decode())
        return out

    def counters(self):
        c = T.Counters()
        if self._H.vkph_engine_counters(self._h, C.byref(c)) != 0:
            raise VkpError(self._H.vkph_last_error().decode())
        return c

    def state(self):
        a, b = C.c_int(), C.c_int()
        self._H.vkph_engine_state(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def close(self):
        if self._h:
            self._H.vkph_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Group:
    """One vkp_group: several GPUs of the box render every frame together in ONE process (tiles dealt over the members, the
    composite lands in rank 0's image over NVLink peer stores).  Thin: every method is one vkp_group_* call."""

    def __init__(self, devices, tile=(32, 16)):
        self.L = load_vkp()
        h = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        r = self.L.vkp_group_create(C.byref(h), arr, len(devices), tile[0], tile[1])
        if r != 0:
            raise VkpError(f"vkp_group_create failed ({r}): {self.L.vkp_last_error(None).decode()}")
        self.h = h
        self.width = self.height = 0

    def _ck(self, r, what):
        if r != 0:
            raise VkpError(f"{what} failed ({r}): {self.L.vkp_group_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.L.vkp_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self):
        return self.L.vkp_group_size(self.h)

    def upload_scene(self, scene):
        for verts, idx in scene.meshes:
            verts = np.ascontiguousarray(verts, dtype=T.vertex_dtype); idx = np.ascontiguousarray(idx, dtype=np.uint32)
            mid = C.c_int(-1)
            self._ck(self.L.vkp_group_upload_mesh(self.h, verts.ctypes.data, len(verts), idx.ctypes.data, len(idx), C.byref(mid)), "vkp_group_upload_mesh")
        for tex in scene.textures:
            rgba = np.ascontiguousarray(tex, dtype=np.uint8)
            tid = C.c_int(-1)
            self._ck(self.L.vkp_group_upload_texture_rgba8(self.h, rgba.ctypes.data, rgba.shape[1], rgba.shape[0], C.byref(tid)), "vkp_group_upload_texture_rgba8")
        self._ck(self.L.vkp_group_set_env(self.h, scene.sky_tex, scene.irradiance_tex), "vkp_group_set_env")
        m = np.ascontiguousarray(scene.materials, dtype=T.material_dtype)
        self._ck(self.L.vkp_group_set_materials(self.h, m.ctypes.data, len(m)), "vkp_group_set_materials")
        self.set_lights(scene.lights)
        inst = np.ascontiguousarray(scene.instances, dtype=T.instance_dtype)
        self._ck(self.L.vkp_group_set_instances(self.h, inst.ctypes.data, len(inst)), "vkp_group_set_instances")
        self._ck(self.L.vkp_group_build_blas(self.h), "vkp_group_build_blas")
        self._ck(self.L.vkp_group_build_tlas(self.h, 0), "vkp_group_build_tlas")

    def set_lights(self, lights):
        l = np.ascontiguousarray(lights, dtype=T.light_dtype)
        self._ck(self.L.vkp_group_set_lights(self.h, l.ctypes.data, len(l)), "vkp_group_set_lights")

    def set_camera(self, cam):
        self._ck(self.L.vkp_group_set_camera(self.h, C.byref(cam)), "vkp_group_set_camera")

    def set_params(self, p):
        self.width, self.height = p.width, p.height
        self._ck(self.L.vkp_group_set_params(self.h, C.byref(p)), "vkp_group_set_params")

    def reset_history(self):
        self._ck(self.L.vkp_group_reset_history(self.h), "vkp_group_reset_history")

    def render(self):
        self._ck(self.L.vkp_group_render(self.h), "vkp_group_render")

    def sync(self):
        self._ck(self.L.vkp_group_sync(self.h), "vkp_group_sync")

    def readback(self, plane):
        dt, ch = plane_format(plane)
        shape = (self.height, self.width, ch) if ch > 1 else (self.height, self.width)
        out = np.empty(shape, dtype=dt)
        self._ck(self.L.vkp_group_readback(self.h, plane, out.ctypes.data, out.nbytes), "vkp_group_readback")
        return out

    def counters(self):
        c = T.Counters()
        self._ck(self.L.vkp_group_get_counters(self.h, C.byref(c)), "vkp_group_get_counters")
        return c

