// vkpinut.h -- host-side mirror of the reference's C++ API for the Hybrid hot path.
//
// Same class names, members and call meaning as the reference, minus every
// Vulkan-typed member: the GPU work goes through the C-ABI of include/vkp.h and nothing else.
//   renderMode        src/vk_engine.h:42-46        Material / GPUMaterial   src/material.h:5-37
//   Camera            src/camera.h:7-46            Texture                  src/vk_textures.h:8-15
//   Vertex/Mesh/Primitive/Node/Prefab/TlasInstance src/vk_mesh.h:21-155
//   Entity/Object/Light                            src/entity.h:6-69
//   Scene             src/scene.h:7-22             Renderer                 src/renderer.h:36-281
//   VulkanEngine      src/vk_engine.h:91-228 (headless: no window / swapchain / ImGui)
#pragma once

#include <cstdint>
#include <functional>
#include <string>
#include <unordered_map>
#include <vector>

#include "vkm.h"
#include "../../include/vkp.h"

namespace glm = vkm;   // reference code spells glm::; the subset it uses lives in vkm.h

enum renderMode { DEFERRED, RAYTRACING, HYBRID };

// ---- material.h ---------------------------------------------------------------------------------
struct GPUMaterial {
    glm::vec4 diffuseColor;
    glm::vec4 textures;
    glm::vec4 shadingMetallicRoughness;
};

class Material {
public:
    static std::vector<Material*> _materials;

    int shadingModel{0};   // 0 metallic-roughness ("difus"), 3 mirror ("mirall"), 4 glass ("vidre")
    glm::vec4 diffuseColor = glm::vec4(1);
    float metallicFactor{1.f};
    float roughnessFactor{1.f};
    float ior{1.f};
    float uvFactor{1.f};
    int diffuseTexture{-1};
    int metallicRoughnessTexture{-1};
    int emissiveTexture{-1};
    int normalTexture{-1};

    static int setDefaultMaterial();
    static bool exists(Material* material);
    static int getIndex(Material* material);
    GPUMaterial materialToShader();
    bool operator==(const Material& m) const;
};

// ---- camera.h -----------------------------------------------------------------------------------
enum Camera_Movement { FORWARD, BACKWARD, LEFT, RIGHT, UP, DOWN };

class Camera {
public:
    Camera(glm::vec3 position = glm::vec3(0.0f, 0.0f, 0.0f), glm::vec3 up = glm::vec3(0.0f, 1.0f, 0.0f),
           float fov = 60.0f, float yaw = -90.0f, float pitch = 0.0f);

    glm::vec3 _position, _direction, _up, _right;
    float _yaw, _pitch, _speed, _sensitivity, _fov;

    void processKeyboard(Camera_Movement direction, const float dt);
    void rotate(float xoffset, float yoffset, bool constrainPitch = true);
    glm::mat4 getView();
    glm::mat4 getProjection(const float ratio);

private:
    void updateCameraVectors();
};

// ---- vk_textures.h ------------------------------------------------------------------------------
// Index in _textures == shader texture id.  Everything is 8-bit RGBA (stbi_load(..., 4)).
struct Texture {
    static std::vector<Texture*> _textures;
    static std::unordered_map<std::string, int> _ids;

    std::string _name;
    int _width{0}, _height{0};
    std::vector<uint8_t> _rgba;

    static Texture* GET(const char* filename);                       // registered image, else PNG / JPEG / HDR file on the search paths, else nullptr
    static int get_id(const char* filename);                         // loads on first use like the reference; -1 when there is no such image
    static int add(const char* name, int w, int h, std::vector<uint8_t> rgba);
};

// ---- vk_mesh.h ----------------------------------------------------------------------------------
struct Vertex {
    glm::vec3 position, normal, color;
    glm::vec2 uv;
    bool operator==(const Vertex& o) const { return position == o.position && color == o.color && uv.x == o.uv.x && uv.y == o.uv.y; }
};
static_assert(sizeof(Vertex) == 44, "Vertex must stay 44 bytes (vk_mesh.h:21-27)");

struct TlasInstance {
    uint32_t blasId{0};
    uint32_t instanceId{0};
    uint32_t hitGroupId{0};
    uint32_t mask{0xFF};
    uint32_t flags{0x1};   // VK_GEOMETRY_INSTANCE_TRIANGLE_FACING_CULL_DISABLE_BIT_KHR
    glm::mat4 transform{glm::mat4(1)};
};

struct Primitive {
    uint32_t firstIndex{0}, indexCount{0}, firstVertex{0}, vertexCount{0};
    int32_t materialID{0}, instanceID{0}, transformID{0};
};

struct Mesh {
    static std::unordered_map<std::string, Mesh*> _loadedMeshes;
    std::vector<Vertex> _vertices;
    std::vector<uint32_t> _indices;
    int _deviceMeshId{-1};    // handle from vkp_upload_mesh (replaces _vertexBuffer/_indexBuffer)

    static Mesh* GET(const char* filename);   // OBJ on the search path, else the deterministic stand-in
    static Mesh* get_quad();
    static Mesh* get_triangle();
    static Mesh* get_cube();
    void upload(vkp_ctx* ctx);                // Mesh::upload (vk_mesh.cpp:265): vkp_upload_mesh
    bool load_from_obj(const char* filename);
};

class Node {
public:
    std::string _name;
    Node* _parent{nullptr};
    std::vector<Node*> _children;
    std::vector<Primitive*> _primitives;
    glm::mat4 _matrix{glm::mat4(1)};
    glm::mat4 _global_matrix{glm::mat4(1)};

    glm::mat4 getGlobalMatrix(bool fast = false);
    void node_to_instance(std::vector<TlasInstance>& instances, int& index, glm::mat4 model);
    unsigned int get_number_nodes();
    void fill_matrix_buffer(std::vector<glm::mat4>& buffer, const glm::mat4 model);
    void fill_index_buffer(std::vector<glm::vec4>& buffer);
    void addMaterial(Material* mat);
};

class Prefab {
public:
    static std::unordered_map<std::string, Prefab*> _prefabsMap;
    std::string _name;
    Mesh* _mesh{nullptr};
    std::vector<Node*> _root;

    static Prefab* GET(std::string filename, const bool invertNormals = false);   // "x.obj" / "x.gltf" / "x.glb" (stand-in when absent)
    static Prefab* GET(std::string name, Mesh* mesh);
};

// vkutil::findFile (vk_utils.cpp:30-60) over the search paths of vk_engine.cpp:29-34 plus $VKPINUT_ASSETS; "" when absent
std::string vkutil_find_file(const std::string& filename);
// glTF branch of Prefab::GET (vk_mesh.cpp:857-896), gltf.cpp; throws std::runtime_error on a malformed file
Prefab* load_gltf_prefab(const std::string& path, bool binary, bool invertNormals);

// ---- entity.h -----------------------------------------------------------------------------------
enum lightType { DIRECTIONAL_LIGHT, POINT_LIGHT, SPOT_LIGHT };

class Entity {
public:
    glm::mat4 m_matrix{glm::mat4(1)};
    bool m_visible{true};
    bool m_selected{false};
    virtual ~Entity() = default;
    virtual void update() = 0;
    virtual void setColor(glm::vec3 color) = 0;
};

class Object : public Entity {
public:
    Prefab* prefab{nullptr};
    Material* material{nullptr};
    int materialIdx{0};
    int id{0};
    Object(glm::vec3 position = glm::vec3(0), Mesh* mesh = nullptr, Material* material = nullptr);
    void update() override {}
    void setColor(glm::vec3) override {}
};

class Light : public Entity {
public:
    lightType type;
    glm::vec3 color;
    glm::vec3 position;
    float intensity, maxDistance, radius;
    Light(lightType type = POINT_LIGHT, glm::vec3 color = glm::vec3(1), glm::vec3 position = glm::vec3(0),
          float intensity = 1000.0f, float maxDistance = 500.0f, float radius = 1.0f);
    void update() override;
    void setColor(glm::vec3 c) override { color = c; }
};

// ---- scene.h ------------------------------------------------------------------------------------
// Tessellation of the procedural stand-ins for the git-ignored assets and the instance count of the stress scene.
// A property of the scene being created -- never library-global state: a scene built without an explicit detail is the full one
// (356 110 triangles for scene 0, 10.16 M instanced triangles for scene 2).
struct MeshDetail {
    int lucy_subdiv{7}, helmet_subdiv{5}, stress_subdiv{7}, stress_instances{31};
};

class Scene {
public:
    MeshDetail _detail;                 // installed by create_scene / create_viewer_scene before any mesh is generated
    std::vector<Object*> _entities;
    std::vector<Light*> _lights;
    std::vector<glm::mat4> _matricesVector;
    Camera* _camera{nullptr};
    int _skyTexture{-1}, _irradianceTexture{-1};   // environmentTexture[0], [1] (renderer.cpp:3073-3074)

    unsigned int get_drawable_nodes_size();
    void create_scene(int i);   // 0 default, 1 cornell, 2 instanced stress (cfg4), 3 ring lights (cfg5)
    void create_viewer_scene(const std::string& prefabFile);   // the default scene's floor, lights and camera around one .obj / .gltf / .glb
    static void reset_registries();   // clear the static registries (tests create several scenes)

private:
    void default_scene();
    void cornell_scene();
    void stress_scene();
    void ring_lights_scene();
    void finish_environment();
};

// Flat buffers the renderer binds (hybrid descriptor set 0-12): what both the C-ABI and the
// CPU oracle consume.
struct SceneBuffers {
    std::vector<Mesh*> meshes;                 // distinct meshes, index = mesh id
    std::vector<vkp_instance> instances;       // TLAS order
    std::vector<vkp_material> materials;
    std::vector<vkp_light> lights;
    std::vector<Texture*> textures;
    int sky_tex{-1}, irradiance_tex{-1};
};
SceneBuffers build_scene_buffers(Scene* scene);
vkp_camera build_camera(Camera* cam, float ratio, const glm::mat4* prevView, const glm::mat4* prevProj);
std::vector<vkp_light> build_lights(Scene* scene);   // VulkanEngine::update light packing (vk_engine.cpp:146-164)

// ---- renderer.h ---------------------------------------------------------------------------------
class Renderer {
public:
    explicit Renderer(Scene* scene);
    ~Renderer();

    Scene* _scene;
    vkp_ctx* _ctx{nullptr};            // the device context; with several GPUs: rank 0 of the group (owner of the composited image)
    vkp_group* _group{nullptr};        // VulkanEngine::_gpus > 1: one context per GPU, tiles dealt over them (include/vkp.h "multi-GPU group")
    std::vector<TlasInstance> _tlas;
    int _width, _height;

    void render();             // DEFERRED  (renderer.cpp:403)
    void raytrace();           // RAYTRACING (renderer.cpp:465)
    void rasterize_hybrid();   // HYBRID    (renderer.cpp:530)
    void render_gui() {}       // ImGui is out of scope; its knobs are VulkanEngine members
    void buildTlas(const std::vector<TlasInstance>& input, bool update = false);
    void recreate_renderer();  // resize
    void create_storage_image() {}

    // headless outputs
    std::vector<uint8_t> read_final_srgb8();
    std::vector<float> read_final_linear();
    void read_into(int plane, void* dst, size_t bytes);   // straight into caller (e.g. pinned) memory
    void present_async(void* dst, size_t bytes);          // queue the D2H of the frame just submitted, flip the output buffer
    void present_wait(int keep_in_flight = 0);            // until at most keep_in_flight queued presents are outstanding
    int _frameSlot{0};                                    // _frameNumber % FRAME_OVERLAP (renderer.cpp:363)
    vkp_counters counters();
    void set_lights(const std::vector<vkp_light>& lights);   // binding 4, every frame (vk_engine.cpp:146-164)

private:
    void load_data_to_gpu();
    void create_bottom_acceleration_structure();
    void create_top_acceleration_structure();
    void submit(renderMode mode);
    std::vector<vkp_instance> _instances;
};

// ---- vk_engine.h --------------------------------------------------------------------------------
class VulkanEngine {
public:
    static VulkanEngine* engine;

    Renderer* renderer{nullptr};
    Scene* _scene{nullptr};
    renderMode _mode{HYBRID};
    int _frameNumber{0};
    bool _denoise{true};
    int _denoise_frame{0};
    int _samples{1};
    int debugTarget{0};
    bool _skyboxFollow{true};
    int _width{1700}, _height{900};     // hard-coded window of vk_engine.cpp:27
    int _sceneIndex{0};
    MeshDetail _detail;                 // handed to the Scene in init()
    std::string _prefabFile;            // headless CLI --prefab: Scene::create_viewer_scene instead of create_scene(_sceneIndex)
    int _device{0};
    int _gpus{1};                       // GPUs _device .. _device + _gpus - 1 render one frame together (tile-sharded, bit-identical image)
    int _spp{1};
    int _maxBounces{8};
    bool _faithfulRng{true};
    bool _reprojectFix{false};          // vkp_params.reproject_fix: false = the reference's blur.comp history fetch, true = the previous frame's pixel
    int _chainMode{7};                  // vkp_params.wavefront: how the reflection chain is scheduled (include/vkp_types.h; 7 = split by predicted length, overlapped, late merge)

    // headless replacement of Window::input_update (window.cpp:40-110): called at the top of every frame of run(), before
    // update(); a script moves the camera there (Camera::processKeyboard / rotate), which resets the accumulation like a key press does
    std::function<void(int frame)> _onFrame;

    void init();
    void run(int frames = 1);
    void update(const float dt);
    void cleanup();
    void updateFrame();
    void resetFrame();
    void updateCameraMatrices();

    vkp_camera _cameraData{};
    vkp_params current_params() const;

private:
    glm::mat4 _refMatrix{glm::mat4(0.0f)};
    glm::mat4 _prevView{glm::mat4(0.0f)}, _prevProj{glm::mat4(0.0f)};
    bool _havePrev{false};
};

// image writers (headless replacement of the swapchain present)
bool write_png_rgba8(const char* path, const uint8_t* rgba, int w, int h);
bool write_exr_rgba32f(const char* path, const float* rgba, int w, int h);
