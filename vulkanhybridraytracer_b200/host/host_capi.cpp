// host_capi.cpp -- flat C shim over VulkanEngine / Renderer (libvkp_host.so, on top of libvkp_scene.so and libvkp.so) so that the
// Python test-suite and bench.py can drive the host API with ctypes.
#include "vkpinut.h"

#include <cstring>
#include <exception>
#include <string>

extern "C" void vkph_set_error(const char* msg);      // scene_capi.cpp: one error slot behind vkph_last_error
namespace { struct ErrSlot { ErrSlot& operator=(const char* m) { vkph_set_error(m); return *this; } ErrSlot& operator=(const std::string& m) { vkph_set_error(m.c_str()); return *this; } } g_err; }

static MeshDetail detail_of(const int* d)
{
    MeshDetail m;
    if (d) {
        if (d[0] >= 0) m.lucy_subdiv = d[0];
        if (d[1] >= 0) m.helmet_subdiv = d[1];
        if (d[2] >= 0) m.stress_subdiv = d[2];
        if (d[3] > 0) m.stress_instances = d[3];
    }
    return m;
}

extern "C" {

#define VKPH_API __attribute__((visibility("default")))

// ---- engine --------------------------------------------------------------------------------------------------
// gpus > 1: GPUs device .. device + gpus - 1 render every frame together (VulkanEngine::_gpus, one process, vkp_group)
VKPH_API void* vkph_engine_create_multi(int scene_index, int width, int height, int device, int gpus, const int* detail4)
{
    VulkanEngine* e = new VulkanEngine();
    try {
        e->_detail = detail_of(detail4);
        e->_sceneIndex = scene_index; e->_width = width; e->_height = height; e->_device = device; e->_gpus = gpus < 1 ? 1 : gpus;
        e->init();
        return e;
    } catch (const std::exception& ex) { g_err = ex.what(); delete e; return nullptr; }
}
VKPH_API void* vkph_engine_create(int scene_index, int width, int height, int device, const int* detail4)
{
    return vkph_engine_create_multi(scene_index, width, height, device, 1, detail4);
}

VKPH_API int vkph_engine_set(void* p, const char* key, int value)
{
    VulkanEngine* e = (VulkanEngine*)p;
    std::string k(key);
    if (k == "mode") e->_mode = (renderMode)value;
    else if (k == "samples") e->_samples = value;
    else if (k == "denoise") e->_denoise = value != 0;
    else if (k == "spp") e->_spp = value;
    else if (k == "max_bounces") e->_maxBounces = value;
    else if (k == "faithful_rng") e->_faithfulRng = value != 0;
    else if (k == "debug_target") e->debugTarget = value;
    else if (k == "chain") e->_chainMode = value;
    else if (k == "reproject_fix") e->_reprojectFix = value != 0;
    else return -1;
    return 0;
}

VKPH_API int vkph_engine_frames(void* p, int frames)
{
    VulkanEngine* e = (VulkanEngine*)p;
    try { e->run(frames); return 0; } catch (const std::exception& ex) { g_err = ex.what(); return -1; }
}

VKPH_API int vkph_engine_camera_move(void* p, int direction, float dt) { ((VulkanEngine*)p)->_scene->_camera->processKeyboard((Camera_Movement)direction, dt); return 0; }

VKPH_API int vkph_engine_read(void* p, int linear, void* dst, size_t bytes)
{
    VulkanEngine* e = (VulkanEngine*)p;
    try {
        size_t need = (size_t)e->_width * e->_height * (linear ? 16 : 4);
        if (bytes < need) { g_err = "vkph_engine_read: destination too small"; return -1; }
        e->renderer->read_into(linear ? VKP_PLANE_FINAL_F32 : VKP_PLANE_FINAL_SRGB8, dst, need);
        return 0;
    } catch (const std::exception& ex) { g_err = ex.what(); return -1; }
}

// present with FRAME_OVERLAP frames in flight: queue the D2H of the frame just rendered into pinned `dst`; wait = all landed
VKPH_API int vkph_engine_present(void* p, void* dst, size_t bytes)
{
    VulkanEngine* e = (VulkanEngine*)p;
    try {
        size_t need = (size_t)e->_width * e->_height * 4;
        if (bytes < need) { g_err = "vkph_engine_present: destination too small"; return -1; }
        e->renderer->present_async(dst, need);
        return 0;
    } catch (const std::exception& ex) { g_err = ex.what(); return -1; }
}
VKPH_API int vkph_engine_present_wait(void* p, int keep_in_flight)
{
    try { ((VulkanEngine*)p)->renderer->present_wait(keep_in_flight); return 0; } catch (const std::exception& ex) { g_err = ex.what(); return -1; }
}

VKPH_API int vkph_engine_counters(void* p, vkp_counters* out)
{
    VulkanEngine* e = (VulkanEngine*)p;
    try { *out = e->renderer->counters(); return 0; } catch (const std::exception& ex) { g_err = ex.what(); return -1; }
}

VKPH_API int vkph_engine_state(void* p, int* denoise_frame, int* frame_number)
{
    VulkanEngine* e = (VulkanEngine*)p;
    *denoise_frame = e->_denoise_frame; *frame_number = e->_frameNumber;
    return 0;
}

VKPH_API void* vkph_engine_ctx(void* p) { return ((VulkanEngine*)p)->renderer->_ctx; }

VKPH_API void vkph_engine_destroy(void* p)
{
    VulkanEngine* e = (VulkanEngine*)p;
    if (!e) return;
    e->cleanup();
    delete e;
}

} // extern "C"
