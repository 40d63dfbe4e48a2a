// scene_capi.cpp -- flat C shim over the host scene model (Scene / Prefab / Mesh / Material / Camera) and the image readers /
// writers: libvkp_scene.so.  No CUDA and no libvkp behind it -- the scene export gives the exact arrays that both the C-ABI
// (include/vkp.h) and the CPU checker consume, so a process that only needs a scene (bench.py --impl reference) maps this library alone.
#include "vkpinut.h"
#include "image_decode.h"

#include <cstring>
#include <exception>
#include <string>

struct vkph_scene {
    Scene* scene;
    SceneBuffers buffers;
};

static thread_local std::string g_err;
// shared with libvkp_host.so (host_capi.cpp), which reports its errors through the same vkph_last_error
extern "C" __attribute__((visibility("default"))) void vkph_set_error(const char* msg) { g_err = msg ? msg : ""; }

extern "C" {

#define VKPH_API __attribute__((visibility("default")))

VKPH_API const char* vkph_last_error(void) { return g_err.c_str(); }

// detail4 = {lucy subdivisions, helmet subdivisions, stress-blob subdivisions, stress instances}; NULL = the full-size scene.
// Negative / zero entries keep the full-size default of that entry.
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

VKPH_API void* vkph_scene_create(int index, const int* detail4)
{
    try {
        Scene::reset_registries();
        vkph_scene* s = new vkph_scene();
        s->scene = new Scene();
        s->scene->_detail = detail_of(detail4);
        s->scene->create_scene(index);
        s->buffers = build_scene_buffers(s->scene);
        return s;
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}

// floor + lights + camera of the default scene around one asset file (.obj / .gltf / .glb): Scene::create_viewer_scene
VKPH_API void* vkph_scene_create_prefab(const char* file, const int* detail4)
{
    try {
        Scene::reset_registries();
        vkph_scene* s = new vkph_scene();
        s->scene = new Scene();
        s->scene->_detail = detail_of(detail4);
        s->scene->create_viewer_scene(file);
        s->buffers = build_scene_buffers(s->scene);
        return s;
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}

VKPH_API void vkph_scene_destroy(void* p)
{
    vkph_scene* s = (vkph_scene*)p;
    if (!s) return;
    delete s->scene;
    delete s;
}

VKPH_API int vkph_scene_counts(void* p, int* n_meshes, int* n_instances, int* n_materials, int* n_lights, int* n_textures, int* sky, int* irr)
{
    vkph_scene* s = (vkph_scene*)p;
    *n_meshes = (int)s->buffers.meshes.size(); *n_instances = (int)s->buffers.instances.size();
    *n_materials = (int)s->buffers.materials.size(); *n_lights = (int)s->buffers.lights.size();
    *n_textures = (int)s->buffers.textures.size(); *sky = s->buffers.sky_tex; *irr = s->buffers.irradiance_tex;
    return 0;
}

VKPH_API int vkph_scene_mesh(void* p, int i, const vkp_vertex** v, uint32_t* nv, const uint32_t** idx, uint32_t* ni)
{
    vkph_scene* s = (vkph_scene*)p;
    if (i < 0 || i >= (int)s->buffers.meshes.size()) return -1;
    Mesh* m = s->buffers.meshes[i];
    *v = reinterpret_cast<const vkp_vertex*>(m->_vertices.data()); *nv = (uint32_t)m->_vertices.size();
    *idx = m->_indices.data(); *ni = (uint32_t)m->_indices.size();
    return 0;
}

VKPH_API const vkp_instance* vkph_scene_instances(void* p) { return ((vkph_scene*)p)->buffers.instances.data(); }
VKPH_API const vkp_material* vkph_scene_materials(void* p) { return ((vkph_scene*)p)->buffers.materials.data(); }
VKPH_API const vkp_light* vkph_scene_lights(void* p) { return ((vkph_scene*)p)->buffers.lights.data(); }

VKPH_API int vkph_scene_texture(void* p, int i, const uint8_t** rgba, int* w, int* h)
{
    vkph_scene* s = (vkph_scene*)p;
    if (i < 0 || i >= (int)s->buffers.textures.size()) return -1;
    Texture* t = s->buffers.textures[i];
    *rgba = t->_rgba.data(); *w = t->_width; *h = t->_height;
    return 0;
}

// camera: static (prev == current) unless have_prev, in which case prev_view/prev_proj are taken from `prev`
VKPH_API int vkph_scene_camera(void* p, float ratio, const vkp_camera* prev, vkp_camera* out)
{
    vkph_scene* s = (vkph_scene*)p;
    if (prev) {
        glm::mat4 pv, pp;
        std::memcpy(pv.data(), prev->view, 64); std::memcpy(pp.data(), prev->proj, 64);
        *out = build_camera(s->scene->_camera, ratio, &pv, &pp);
    } else {
        *out = build_camera(s->scene->_camera, ratio, nullptr, nullptr);
    }
    return 0;
}
VKPH_API int vkph_scene_camera_move(void* p, int direction, float dt) { ((vkph_scene*)p)->scene->_camera->processKeyboard((Camera_Movement)direction, dt); return 0; }
VKPH_API int vkph_scene_camera_rotate(void* p, float dx, float dy) { ((vkph_scene*)p)->scene->_camera->rotate(dx, dy); return 0; }
VKPH_API int vkph_scene_camera_set(void* p, const float pos[3], float yaw, float pitch)
{
    Camera* c = ((vkph_scene*)p)->scene->_camera;
    *c = Camera(glm::vec3(pos[0], pos[1], pos[2]), glm::vec3(0, 1, 0), c->_fov, yaw, pitch);
    return 0;
}
// move entity i (gizmo equivalent, renderer.cpp:761-774) and rebuild the instance records
VKPH_API int vkph_scene_set_entity_matrix(void* p, int entity, const float m[16])
{
    vkph_scene* s = (vkph_scene*)p;
    if (entity < 0 || entity >= (int)s->scene->_entities.size()) return -1;
    std::memcpy(s->scene->_entities[entity]->m_matrix.data(), m, 64);
    s->buffers = build_scene_buffers(s->scene);
    return 0;
}

VKPH_API int vkph_write_png(const char* path, const uint8_t* rgba, int w, int h) { return write_png_rgba8(path, rgba, w, h) ? 0 : -1; }
// stbi_load(path, ..., STBI_rgb_alpha) replacement (image_decode.cpp): dst == NULL only reports the size
VKPH_API int vkph_decode_image_file(const char* path, int* w, int* h, uint8_t* dst, size_t capacity)
{
    std::vector<uint8_t> rgba;
    std::string err;
    if (!imgdec::load_rgba8(path, *w, *h, rgba, &err)) { g_err = err; return -1; }
    if (dst) {
        if (capacity < rgba.size()) { g_err = "vkph_decode_image_file: destination too small"; return -2; }
        std::memcpy(dst, rgba.data(), rgba.size());
    }
    return 0;
}

VKPH_API int vkph_write_exr(const char* path, const float* rgba, int w, int h) { return write_exr_rgba32f(path, rgba, w, h) ? 0 : -1; }

} // extern "C"
