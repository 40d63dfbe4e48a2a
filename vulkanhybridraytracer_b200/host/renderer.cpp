// renderer.cpp -- Renderer and VulkanEngine of the reference (src/renderer.cpp, src/vk_engine.cpp) re-hosted on
// the C-ABI of include/vkp.h.  The 5 vkQueueSubmit + vkQueueWaitIdle pairs of Renderer::rasterize_hybrid
// (renderer.cpp:544-588) collapse into one asynchronous vkp_render on the context stream.
#include "vkpinut.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#define VKP_CHECK(x)                                                                                              \
    do {                                                                                                          \
        int _e = (x);                                                                                             \
        if (_e != VKP_OK) throw std::runtime_error(std::string(#x) + ": " + (_group ? vkp_group_last_error(_group) : vkp_last_error(_ctx))); \
    } while (0)
// one GPU: the context call; several: the group call of the same name (every member)
#define VKP_ALL(name, ...) VKP_CHECK(_group ? vkp_group_##name(_group, __VA_ARGS__) : vkp_##name(_ctx, __VA_ARGS__))
#define VKP_ALL0(name) VKP_CHECK(_group ? vkp_group_##name(_group) : vkp_##name(_ctx))

// ---- Mesh::upload (vk_mesh.cpp upload_to_gpu: vertex + index buffers) -------------------------------------------
void Mesh::upload(vkp_ctx* ctx)
{
    if (_deviceMeshId >= 0) return;
    static_assert(sizeof(Vertex) == sizeof(vkp_vertex), "Vertex / vkp_vertex layout");
    int id = -1;
    if (vkp_upload_mesh(ctx, reinterpret_cast<const vkp_vertex*>(_vertices.data()), (uint32_t)_vertices.size(), _indices.data(),
                        (uint32_t)_indices.size(), &id) == VKP_OK)
        _deviceMeshId = id;
}

// ---- Renderer ---------------------------------------------------------------------------------------------
Renderer::Renderer(Scene* scene) : _scene(scene)
{
    VulkanEngine* e = VulkanEngine::engine;
    _width = e ? e->_width : 1700; _height = e ? e->_height : 900;
    int dev = e ? e->_device : 0;
    const int gpus = e ? e->_gpus : 1;
    if (gpus > 1) {
        std::vector<int> devs;
        const bool same = std::getenv("VKP_GROUP_SAME_DEVICE") != nullptr;     // tests on a one-GPU box: every rank on `dev`
        for (int k = 0; k < gpus; k++) devs.push_back(same ? dev : dev + k);
        if (vkp_group_create(&_group, devs.data(), gpus, 32, 16) != VKP_OK) throw std::runtime_error(std::string("vkp_group_create: ") + vkp_last_error(nullptr));
        _ctx = vkp_group_member(_group, 0);
    } else if (vkp_create(&_ctx, dev) != VKP_OK) throw std::runtime_error(std::string("vkp_create: ") + vkp_last_error(nullptr));
    load_data_to_gpu();
    create_bottom_acceleration_structure();
    create_top_acceleration_structure();
}

Renderer::~Renderer()
{
    if (_group) vkp_group_destroy(_group);
    else if (_ctx) vkp_destroy(_ctx);
}

// renderer.cpp:1477-1530 (lights, materials, matrices) + create_hybrid_descriptors (:2917-3129: vertex /
// index / id / texture / environment bindings)
void Renderer::load_data_to_gpu()
{
    SceneBuffers B = build_scene_buffers(_scene);
    for (Mesh* m : B.meshes) m->_deviceMeshId = -1;
    for (size_t i = 0; i < B.meshes.size(); i++) {
        if (_group) {
            Mesh* m = B.meshes[i];
            int id = -1;
            VKP_CHECK(vkp_group_upload_mesh(_group, reinterpret_cast<const vkp_vertex*>(m->_vertices.data()), (uint32_t)m->_vertices.size(), m->_indices.data(),
                                            (uint32_t)m->_indices.size(), &id));
            m->_deviceMeshId = id;
        } else B.meshes[i]->upload(_ctx);
        if (B.meshes[i]->_deviceMeshId != (int)i) throw std::runtime_error("mesh upload order");
    }
    for (Texture* t : B.textures) {
        int id = -1;
        VKP_ALL(upload_texture_rgba8, t->_rgba.data(), t->_width, t->_height, &id);
    }
    VKP_ALL(set_env, B.sky_tex, B.irradiance_tex);
    VKP_ALL(set_materials, B.materials.data(), (int)B.materials.size());
    VKP_ALL(set_lights, B.lights.data(), (int)B.lights.size());
    _instances = B.instances;
    VKP_ALL(set_instances, _instances.data(), (int)_instances.size());
    _tlas.clear();
    for (size_t i = 0; i < _instances.size(); i++) {
        TlasInstance t; t.blasId = t.instanceId = (uint32_t)i;
        std::memcpy(t.transform.data(), _instances[i].model, 64);
        _tlas.push_back(t);
    }
}

void Renderer::create_bottom_acceleration_structure() { VKP_ALL0(build_blas); }
void Renderer::create_top_acceleration_structure() { VKP_ALL(build_tlas, 0); }
void Renderer::set_lights(const std::vector<vkp_light>& lights) { VKP_ALL(set_lights, lights.data(), (int)lights.size()); }

// Renderer::buildTlas(input, flags, update) (renderer.cpp:1757-1865)
void Renderer::buildTlas(const std::vector<TlasInstance>& input, bool update)
{
    if (input.size() != _instances.size()) throw std::runtime_error("buildTlas: instance count changed; rebuild the renderer");
    bool changed = false;
    for (size_t i = 0; i < input.size(); i++) {
        if (std::memcmp(_instances[i].model, input[i].transform.data(), 64) != 0) {
            glm::mat4 inv = glm::inverse(input[i].transform);
            std::memcpy(_instances[i].model, input[i].transform.data(), 64);
            std::memcpy(_instances[i].inv_model, inv.data(), 64);
            changed = true;
        }
    }
    // the reference refits every frame whether or not anything moved (vk_engine.cpp:172-187): so do we
    if (changed || !update) VKP_ALL(set_instances, _instances.data(), (int)_instances.size());
    VKP_ALL(build_tlas, update ? 1 : 0);
}

void Renderer::recreate_renderer()
{
    VulkanEngine* e = VulkanEngine::engine;
    if (e) { _width = e->_width; _height = e->_height; }
    VKP_ALL0(reset_history);
}

void Renderer::submit(renderMode mode)
{
    VulkanEngine* e = VulkanEngine::engine;
    vkp_params p = e->current_params();
    p.mode = (int)mode;
    VKP_ALL(set_params, &p);
    VKP_ALL(set_camera, &e->_cameraData);
    if (_group) { VKP_CHECK(vkp_group_render(_group)); return; }       // the group alternates the output buffers itself
    VKP_CHECK(vkp_set_output_buffer(_ctx, _frameSlot));
    VKP_CHECK(vkp_render(_ctx));
}

// The swapchain present of the reference with FRAME_OVERLAP = 2 frames in flight (renderer.h:34, get_current_frame renderer.cpp:363):
// the image of the frame just submitted starts its way to `dst` and the next frame renders into the other output buffer meanwhile.
void Renderer::present_async(void* dst, size_t bytes)
{
    VKP_CHECK(vkp_readback_async(_ctx, VKP_PLANE_FINAL_SRGB8, dst, bytes));
    _frameSlot ^= 1;
}
void Renderer::present_wait(int keep_in_flight) { VKP_CHECK(vkp_readback_wait(_ctx, keep_in_flight)); }

void Renderer::render() { submit(DEFERRED); }
void Renderer::raytrace() { submit(RAYTRACING); }
void Renderer::rasterize_hybrid() { submit(HYBRID); }

void Renderer::read_into(int plane, void* dst, size_t bytes)
{
    if (_group) VKP_CHECK(vkp_group_readback(_group, plane, dst, bytes));
    else VKP_CHECK(vkp_readback(_ctx, plane, dst, bytes));
}

std::vector<uint8_t> Renderer::read_final_srgb8()
{
    std::vector<uint8_t> out((size_t)_width * _height * 4);
    VKP_CHECK(vkp_readback(_ctx, VKP_PLANE_FINAL_SRGB8, out.data(), out.size()));
    return out;
}
std::vector<float> Renderer::read_final_linear()
{
    std::vector<float> out((size_t)_width * _height * 4);
    if (_group) VKP_CHECK(vkp_group_readback(_group, VKP_PLANE_FINAL_F32, out.data(), out.size() * sizeof(float)));   // gathers the tiles first
    else VKP_CHECK(vkp_readback(_ctx, VKP_PLANE_FINAL_F32, out.data(), out.size() * sizeof(float)));
    return out;
}
vkp_counters Renderer::counters()
{
    vkp_counters c{};
    if (_group) VKP_CHECK(vkp_group_get_counters(_group, &c));
    else VKP_CHECK(vkp_get_counters(_ctx, &c));
    return c;
}

// ---- VulkanEngine (headless) ---------------------------------------------------------------------------
VulkanEngine* VulkanEngine::engine = nullptr;

// vk_engine.cpp:23-61 minus window / swapchain / ImGui
void VulkanEngine::init()
{
    engine = this;
    Scene::reset_registries();
    _scene = new Scene();
    _scene->_detail = _detail;
    if (!_prefabFile.empty()) _scene->create_viewer_scene(_prefabFile);
    else _scene->create_scene(_sceneIndex);
    renderer = new Renderer(_scene);
    _denoise_frame = 0; _frameNumber = 0; _havePrev = false;
    _refMatrix = glm::mat4(0.0f);
}

vkp_params VulkanEngine::current_params() const
{
    vkp_params p{};
    p.width = _width; p.height = _height; p.mode = (int)_mode;
    p.spp = _spp; p.shadow_samples = _samples; p.max_bounces = _maxBounces;
    p.frame = _denoise ? _denoise_frame : 0;             // vk_engine.cpp:733
    p.denoise_frame = _denoise_frame;                    // vk_engine.cpp:720-723
    p.faithful_rng = _faithfulRng ? 1 : 0;
    p.denoise_cover_remainder = 1;
    p.debug_target = debugTarget;
    p.wavefront = _chainMode;
    p.reproject_fix = _reprojectFix ? 1 : 0;
    p.fuse_denoise = 0;   // measured: the stand-alone filter pass hides under the overlapped reflection chain, fused it lands on the critical path
    return p;
}

// vk_engine.cpp:669-686
void VulkanEngine::updateFrame()
{
    const glm::mat4 m = _scene->_camera->getView();
    if (m != _refMatrix) { resetFrame(); _refMatrix = m; }
    _denoise_frame++;
}
void VulkanEngine::resetFrame() { _denoise_frame = -1; }

// vk_engine.cpp:688-741
void VulkanEngine::updateCameraMatrices()
{
    float ratio = (float)_width / (float)_height;
    glm::mat4 view = _scene->_camera->getView();
    glm::mat4 projection = _scene->_camera->getProjection(ratio);
    projection[1][1] *= -1;
    if (!_havePrev) { _prevView = view; _prevProj = projection; _havePrev = true; }   // App. A #17: defined static-camera motion
    if (_prevView != view || _prevProj != projection) {
        _cameraData = build_camera(_scene->_camera, ratio, &_prevView, &_prevProj);
        _prevView = view; _prevProj = projection;
    } else {
        _cameraData = build_camera(_scene->_camera, ratio, nullptr, nullptr);
    }
}

// vk_engine.cpp:130-188
void VulkanEngine::update(const float)
{
    updateFrame();
    updateCameraMatrices();
    std::vector<vkp_light> lights = build_lights(_scene);
    renderer->set_lights(lights);
    int instanceIndex = 0;
    renderer->_tlas.clear();
    for (Object* obj : _scene->_entities)
        if (!obj->prefab->_root.empty())
            for (Node* root : obj->prefab->_root) root->node_to_instance(renderer->_tlas, instanceIndex, obj->m_matrix);
    renderer->buildTlas(renderer->_tlas, true);
}

// vk_engine.cpp:91-128 for a fixed number of frames
void VulkanEngine::run(int frames)
{
    for (int f = 0; f < frames; f++) {
        if (_onFrame) _onFrame(_frameNumber);
        update(16.6f);
        if (_mode == DEFERRED) renderer->render();
        else if (_mode == RAYTRACING) renderer->raytrace();
        else if (_mode == HYBRID) renderer->rasterize_hybrid();
        _frameNumber++;
    }
}

void VulkanEngine::cleanup()
{
    delete renderer; renderer = nullptr;
    delete _scene; _scene = nullptr;
    if (engine == this) engine = nullptr;
}
