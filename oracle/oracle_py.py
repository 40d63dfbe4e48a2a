"""TEST INFRASTRUCTURE: ctypes wrapper of the CPU oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this
module.  The product package never does.
"""
import ctypes as C
import os

import numpy as np

from vulkanhybridraytracer_b200 import types as T

_HERE = os.path.dirname(os.path.abspath(__file__))


class OrcMesh(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("n_vertices", C.c_uint32), ("indices", C.c_void_p), ("n_indices", C.c_uint32)]


class OrcTexture(C.Structure):
    _fields_ = [("rgba", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32)]


class OrcSceneDesc(C.Structure):
    _fields_ = [("meshes", C.POINTER(OrcMesh)), ("n_meshes", C.c_int32), ("textures", C.POINTER(OrcTexture)), ("n_textures", C.c_int32),
                ("sky_tex", C.c_int32), ("irradiance_tex", C.c_int32), ("materials", C.c_void_p), ("n_materials", C.c_int32),
                ("lights", C.c_void_p), ("n_lights", C.c_int32), ("instances", C.c_void_p), ("n_instances", C.c_int32)]


class OrcGbuffer(C.Structure):
    _fields_ = [("position", C.c_void_p), ("normal", C.c_void_p), ("albedo", C.c_void_p), ("motion", C.c_void_p),
                ("material", C.c_void_p), ("emissive", C.c_void_p), ("hit_ids", C.c_void_p)]


class OrcCounters(C.Structure):
    _fields_ = [("rays", C.c_uint64 * 4)]


_libs = {}


def load(libm=False):
    key = "libm" if libm else "det"
    if key in _libs:
        return _libs[key]
    path = os.path.join(_HERE, "liboracle_libm.so" if libm else "liboracle.so")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} missing: run `make -C oracle`")
    L = C.CDLL(path)
    vp, i, u32 = C.c_void_p, C.c_int, C.c_uint32
    P = C.POINTER
    L.orc_abi_version.restype = i
    L.orc_num_threads.restype = i
    L.orc_set_num_threads.argtypes = [i]
    L.orc_uses_libm.restype = i
    L.orc_set_row_sampling.argtypes = [i, i]
    L.orc_scene_create.restype = vp; L.orc_scene_create.argtypes = [P(OrcSceneDesc)]
    L.orc_scene_destroy.argtypes = [vp]
    L.orc_scene_set_lights.argtypes = [vp, vp, i]
    L.orc_scene_set_instances.argtypes = [vp, vp, i]
    L.orc_scene_triangles.restype = C.c_uint64; L.orc_scene_triangles.argtypes = [vp]
    L.orc_trace_rays.argtypes = [vp, vp, u32, i, i, vp]
    L.orc_gbuffer_pass.argtypes = [vp, P(T.Camera), P(T.Params), P(OrcGbuffer), P(OrcCounters)]
    L.orc_shadow_pass.argtypes = [vp, P(T.Camera), P(T.Params), P(OrcGbuffer), vp, P(OrcCounters)]
    L.orc_denoise_pass.argtypes = [P(T.Params), i, P(OrcGbuffer), vp, vp]
    L.orc_resolve_pass.argtypes = [vp, P(T.Camera), P(T.Params), P(OrcGbuffer), vp, vp, P(OrcCounters)]
    L.orc_post_pass.argtypes = [P(T.Params), vp, vp]
    L.orc_raytrace_pass.argtypes = [vp, P(T.Camera), P(T.Params), vp, P(OrcCounters)]
    L.orc_deferred_pass.argtypes = [vp, P(T.Camera), P(T.Params), P(OrcGbuffer), vp]
    L.orc_tea.restype = u32; L.orc_tea.argtypes = [u32, u32]
    L.orc_lcg.restype = u32; L.orc_lcg.argtypes = [P(u32)]
    L.orc_rnd.restype = C.c_float; L.orc_rnd.argtypes = [P(u32)]
    L.orc_sample_disk.argtypes = [P(T.Light), P(C.c_float * 3), P(C.c_float * 3), u32, P(C.c_float * 3)]
    L.orc_distribution_ggx.restype = C.c_float; L.orc_distribution_ggx.argtypes = [P(C.c_float * 3), P(C.c_float * 3), C.c_float]
    L.orc_geometry_smith.restype = C.c_float; L.orc_geometry_smith.argtypes = [P(C.c_float * 3)] * 3 + [C.c_float]
    L.orc_fresnel_schlick.argtypes = [C.c_float, P(C.c_float * 3), P(C.c_float * 3)]
    L.orc_env_uv.argtypes = [P(C.c_float * 3), i, P(C.c_float * 2)]
    L.orc_to_half.restype = C.c_uint16; L.orc_to_half.argtypes = [C.c_float]
    L.orc_from_half.restype = C.c_float; L.orc_from_half.argtypes = [C.c_uint16]
    L.orc_to_unorm8.restype = C.c_uint8; L.orc_to_unorm8.argtypes = [C.c_float]
    _libs[key] = L
    return L


class GBuffer:
    """G-buffer planes in the binary layout of enum vkp_plane."""

    def __init__(self, width, height):
        self.width, self.height = width, height
        self.position = np.zeros((height, width, 4), np.float16)
        self.normal = np.zeros((height, width, 4), np.float16)
        self.albedo = np.zeros((height, width, 4), np.uint8)
        self.motion = np.zeros((height, width, 2), np.float16)
        self.material = np.zeros((height, width, 4), np.uint8)
        self.emissive = np.zeros((height, width, 4), np.uint8)
        self.hit_ids = np.full((height, width, 2), -1, np.int32)

    def c(self):
        g = OrcGbuffer()
        for k in ("position", "normal", "albedo", "motion", "material", "emissive", "hit_ids"):
            setattr(g, k, getattr(self, k).ctypes.data)
        return g


class Oracle:
    """The oracle bound to one scene (deep copy on the C side)."""

    def __init__(self, scene, libm=False, threads=None):
        self.L = load(libm)
        if threads:
            self.L.orc_set_num_threads(threads)
        meshes = (OrcMesh * len(scene.meshes))()
        self._keep = []
        for k, (v, ix) in enumerate(scene.meshes):
            v = np.ascontiguousarray(v); ix = np.ascontiguousarray(ix)
            self._keep += [v, ix]
            meshes[k] = OrcMesh(v.ctypes.data, len(v), ix.ctypes.data, len(ix))
        texs = (OrcTexture * max(1, len(scene.textures)))()
        for k, t in enumerate(scene.textures):
            t = np.ascontiguousarray(t)
            self._keep.append(t)
            texs[k] = OrcTexture(t.ctypes.data, t.shape[1], t.shape[0])
        mats = np.ascontiguousarray(scene.materials); lights = np.ascontiguousarray(scene.lights); inst = np.ascontiguousarray(scene.instances)
        d = OrcSceneDesc(meshes, len(scene.meshes), texs, len(scene.textures), scene.sky_tex, scene.irradiance_tex,
                         mats.ctypes.data, len(mats), lights.ctypes.data, len(lights), inst.ctypes.data, len(inst))
        self.h = self.L.orc_scene_create(C.byref(d))
        self.n_lights = len(lights)
        self.counters = OrcCounters()

    def close(self):
        if getattr(self, "h", None):
            self.L.orc_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def threads(self):
        return self.L.orc_num_threads()

    def set_lights(self, lights):
        lights = np.ascontiguousarray(lights, dtype=T.light_dtype)
        self.n_lights = len(lights)
        self.L.orc_scene_set_lights(self.h, lights.ctypes.data, len(lights))

    def set_instances(self, inst):
        inst = np.ascontiguousarray(inst, dtype=T.instance_dtype)
        self.L.orc_scene_set_instances(self.h, inst.ctypes.data, len(inst))

    def set_row_sampling(self, stride, offset=0):
        """bench timing aid: only rows y % stride == offset are processed (1 = every row)."""
        self.L.orc_set_row_sampling(stride, offset)

    def reset_counters(self):
        self.counters = OrcCounters()

    def trace_rays(self, rays, any_hit=False, brute_force=False):
        rays = np.ascontiguousarray(rays, dtype=T.ray_dtype)
        hits = np.empty(len(rays), dtype=T.hit_dtype)
        self.L.orc_trace_rays(self.h, rays.ctypes.data, len(rays), 1 if any_hit else 0, 1 if brute_force else 0, hits.ctypes.data)
        return hits

    def gbuffer(self, cam, p):
        g = GBuffer(p.width, p.height)
        gc = g.c()
        self.L.orc_gbuffer_pass(self.h, C.byref(cam), C.byref(p), C.byref(gc), C.byref(self.counters))
        return g

    def shadow(self, cam, p, g):
        out = np.zeros((self.n_lights, p.height, p.width), np.uint8)
        gc = g.c()
        self.L.orc_shadow_pass(self.h, C.byref(cam), C.byref(p), C.byref(gc), out.ctypes.data, C.byref(self.counters))
        return out

    def denoise(self, p, g, shadow, history):
        hist = np.ascontiguousarray(history, dtype=np.uint8).copy()
        shadow = np.ascontiguousarray(shadow, dtype=np.uint8)
        gc = g.c()
        self.L.orc_denoise_pass(C.byref(p), shadow.shape[0], C.byref(gc), shadow.ctypes.data, hist.ctypes.data)
        return hist

    def resolve(self, cam, p, g, denoised):
        out = np.zeros((p.height, p.width, 4), np.float32)
        denoised = np.ascontiguousarray(denoised, dtype=np.uint8)
        gc = g.c()
        self.L.orc_resolve_pass(self.h, C.byref(cam), C.byref(p), C.byref(gc), denoised.ctypes.data, out.ctypes.data, C.byref(self.counters))
        return out

    def post(self, p, final):
        final = np.ascontiguousarray(final, dtype=np.float32)
        out = np.zeros((p.height, p.width, 4), np.uint8)
        self.L.orc_post_pass(C.byref(p), final.ctypes.data, out.ctypes.data)
        return out

    def raytrace(self, cam, p, accum=None):
        acc = np.zeros((p.height, p.width, 4), np.float32) if accum is None else np.ascontiguousarray(accum, dtype=np.float32).copy()
        self.L.orc_raytrace_pass(self.h, C.byref(cam), C.byref(p), acc.ctypes.data, C.byref(self.counters))
        return acc

    def deferred(self, cam, p, g):
        out = np.zeros((p.height, p.width, 4), np.float32)
        gc = g.c()
        self.L.orc_deferred_pass(self.h, C.byref(cam), C.byref(p), C.byref(gc), out.ctypes.data)
        return out

    def hybrid_frame(self, cam, p, history=None):
        """One Renderer::rasterize_hybrid frame: G-buffer -> shadow -> denoise -> resolve.  Returns a dict of planes."""
        g = self.gbuffer(cam, p)
        sh = self.shadow(cam, p, g)
        if history is None:
            history = np.zeros_like(sh)
        den = self.denoise(p, g, sh, history)
        final = self.resolve(cam, p, g, den)
        return {"g": g, "shadow": sh, "denoised": den, "final": final}
