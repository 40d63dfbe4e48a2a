// scene_model.cpp -- host scene model: Material, Camera, Texture, Mesh, Node, Prefab, Entity, Light, Scene.
// Behaviour follows the reference classes cited in vkpinut.h; asset loading falls back to the
// deterministic stand-ins of procedural.cpp because the reference's assets are git-ignored.
#include "vkpinut.h"
#include "procedural.h"
#include "image_decode.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

// ---- Material (src/material.cpp:4-62) -------------------------------------------------------------
std::vector<Material*> Material::_materials;

int Material::setDefaultMaterial()
{
    Material* mat = new Material();
    if (Material::exists(mat)) { int i = getIndex(mat); delete mat; return i; }
    Material::_materials.push_back(mat);
    return (int)Material::_materials.size() - 1;
}

GPUMaterial Material::materialToShader()
{
    GPUMaterial mat;
    mat.diffuseColor = glm::vec4(diffuseColor[0], diffuseColor[1], diffuseColor[2], ior);
    mat.textures = glm::vec4((float)diffuseTexture, (float)normalTexture, (float)emissiveTexture, (float)metallicRoughnessTexture);
    mat.shadingMetallicRoughness = glm::vec4((float)shadingModel, metallicFactor, roughnessFactor, (float)Material::getIndex(this));
    return mat;
}

bool Material::exists(Material* material)
{
    for (Material* m : _materials) if (*m == *material) return true;
    return false;
}

int Material::getIndex(Material* material)
{
    int count = 0;
    for (Material* m : _materials) { if (*m == *material) return count; count++; }
    return -1;
}

bool Material::operator==(const Material& m) const
{
    return shadingModel == m.shadingModel && diffuseColor == m.diffuseColor && diffuseTexture == m.diffuseTexture &&
           normalTexture == m.normalTexture && emissiveTexture == m.emissiveTexture &&
           metallicRoughnessTexture == m.metallicRoughnessTexture && ior == m.ior && metallicFactor == m.metallicFactor &&
           roughnessFactor == m.roughnessFactor;
}

// ---- Camera (src/camera.cpp:4-70) -------------------------------------------------------------------
Camera::Camera(glm::vec3 position, glm::vec3 up, float fov, float yaw, float pitch)
    : _direction(glm::vec3(0, 0, -1)), _speed(0.1f), _sensitivity(0.1f)
{
    _position = position; _up = up; _yaw = yaw; _pitch = pitch; _fov = fov;
    updateCameraVectors();
}

void Camera::processKeyboard(Camera_Movement direction, const float dt)
{
    float movementSpeed = _speed * dt;
    if (direction == FORWARD) _position += _direction * movementSpeed;
    if (direction == BACKWARD) _position -= _direction * movementSpeed;
    if (direction == RIGHT) _position += _right * movementSpeed;
    if (direction == LEFT) _position -= _right * movementSpeed;
    if (direction == UP) _position += _up * movementSpeed;
    if (direction == DOWN) _position -= _up * movementSpeed;
}

void Camera::rotate(float xoffset, float yoffset, bool constrainPitch)
{
    xoffset *= _sensitivity; yoffset *= _sensitivity;
    _yaw -= xoffset; _pitch += yoffset;
    if (constrainPitch) {
        if (_pitch > 89.0f) _pitch = 89.0f;
        if (_pitch < -89.0f) _pitch = -89.9f;     // sic (camera.cpp:42)
    }
    updateCameraVectors();
}

glm::mat4 Camera::getView() { return glm::lookAt(_position, _position + _direction, glm::vec3(0, 1, 0)); }
glm::mat4 Camera::getProjection(const float ratio) { return glm::perspective(glm::radians(_fov), ratio, 0.1f, 1000.0f); }

void Camera::updateCameraVectors()
{
    glm::vec3 front;
    front.x = std::cos(glm::radians(_yaw)) * std::cos(glm::radians(_pitch));
    front.y = std::sin(glm::radians(_pitch));
    front.z = std::sin(glm::radians(_yaw)) * std::cos(glm::radians(_pitch));
    if (_yaw == -90.0f && _pitch == 0.0f) front = glm::vec3(0, 0, -1);   // exact default view; cos(-pi/2) is 6e-17, not 0
    _direction = glm::normalize(front);
    _right = glm::normalize(glm::cross(_direction, glm::vec3(0, 1, 0)));
    _up = glm::normalize(glm::cross(_right, _direction));
}

// ---- Texture (src/vk_textures.cpp:105-168) --------------------------------------------------------------
std::vector<Texture*> Texture::_textures;
std::unordered_map<std::string, int> Texture::_ids;

int Texture::add(const char* name, int w, int h, std::vector<uint8_t> rgba)
{
    auto it = _ids.find(name);
    if (it != _ids.end()) return it->second;
    Texture* t = new Texture();
    t->_name = name; t->_width = w; t->_height = h; t->_rgba = std::move(rgba);
    _textures.push_back(t);
    _ids[name] = (int)_textures.size() - 1;
    return (int)_textures.size() - 1;
}
Texture* Texture::GET(const char* filename) { int id = get_id(filename); return id < 0 ? nullptr : _textures[id]; }

// Texture::get_id (vk_textures.cpp:156-168) loads the file on first use; the registry is keyed by the resolved path
// (:129).  Names registered with Texture::add (procedural stand-ins, embedded glTF images) are found first.
int Texture::get_id(const char* filename)
{
    auto it = _ids.find(filename);
    if (it != _ids.end()) return it->second;
    std::string path = vkutil_find_file(filename);
    if (path.empty()) return -1;
    it = _ids.find(path);
    if (it != _ids.end()) return it->second;
    int w = 0, h = 0;
    std::vector<uint8_t> rgba;
    std::string err;
    if (!imgdec::load_rgba8(path.c_str(), w, h, rgba, &err))
        throw std::runtime_error("Failed to load texture file: " + path + " (" + err + ")");   // vk_textures.cpp:16-19 only prints; we refuse to go on with an undefined image
    return add(path.c_str(), w, h, std::move(rgba));
}

// ---- Mesh (src/vk_mesh.cpp:72-263) ---------------------------------------------------------------------
std::unordered_map<std::string, Mesh*> Mesh::_loadedMeshes;

// detail of the scene under construction: Scene::create_scene / create_viewer_scene copy their own _detail here first, so a mesh
// generated for one scene can never inherit the tessellation an earlier scene (or test) asked for
static thread_local MeshDetail g_detail;

static std::string find_file(const std::string& filename);
static std::string base_name(const std::string& p)
{
    size_t s = p.find_last_of("/\\");
    return s == std::string::npos ? p : p.substr(s + 1);
}

// vkutil::findFile over the search paths of vk_engine.cpp:29-34, plus $VKPINUT_ASSETS
std::string vkutil_find_file(const std::string& filename) { return find_file(filename); }
static std::string find_file(const std::string& filename)
{
    std::vector<std::string> paths = {"", "data/", "data/meshes/", "data/textures/", "../data/meshes/", "../data/textures/"};
    if (const char* env = std::getenv("VKPINUT_ASSETS")) paths.insert(paths.begin(), std::string(env) + "/");
    for (const std::string& p : paths) {
        std::ifstream f(p + filename);
        if (f.good()) return p + filename;
    }
    return "";
}

Mesh* Mesh::GET(const char* filename)
{
    std::string name = base_name(filename);
    auto it = _loadedMeshes.find(name);
    if (it != _loadedMeshes.end() && it->second) return it->second;
    Mesh* mesh = new Mesh();
    std::string path = find_file(filename);
    bool ok = false;
    if (!path.empty() && name.find(".obj") != std::string::npos) ok = mesh->load_from_obj(path.c_str());
    if (!ok) {
        if (name == "sphere.obj") proc::uv_sphere(64, 32, mesh->_vertices, mesh->_indices);
        else if (name == "lucy.obj") proc::lucy_standin(mesh->_vertices, mesh->_indices, g_detail.lucy_subdiv);
        else if (name == "DamagedHelmet.gltf") proc::helmet_standin(mesh->_vertices, mesh->_indices, g_detail.helmet_subdiv);
        else if (name == "stress_blob") proc::displaced_icosphere(mesh->_vertices, mesh->_indices, g_detail.stress_subdiv, 1234u, 0.35f);
        else { delete mesh; return nullptr; }
    }
    _loadedMeshes[name] = mesh;
    return mesh;
}

// Mesh::load_from_obj (vk_mesh.cpp:96-159) without tinyobjloader: v / vn / vt / f, fan triangulation,
// de-duplication on (position, colour, uv) like Vertex::operator==, uv.y = 1 - v, white vertex colour.
bool Mesh::load_from_obj(const char* filename)
{
    std::ifstream in(filename);
    if (!in.good()) return false;
    std::vector<glm::vec3> P, N; std::vector<glm::vec2> T;
    struct Key { int p, t, n; bool operator==(const Key& o) const { return p == o.p && t == o.t && n == o.n; } };
    struct KeyHash { size_t operator()(const Key& k) const { return ((size_t)k.p * 73856093u) ^ ((size_t)k.t * 19349663u) ^ ((size_t)k.n * 83492791u); } };
    std::unordered_map<Key, uint32_t, KeyHash> unique;
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ss(line);
        std::string tag; ss >> tag;
        if (tag == "v") { glm::vec3 p; ss >> p.x >> p.y >> p.z; P.push_back(p); }
        else if (tag == "vn") { glm::vec3 n; ss >> n.x >> n.y >> n.z; N.push_back(n); }
        else if (tag == "vt") { glm::vec2 t; ss >> t.x >> t.y; T.push_back(t); }
        else if (tag == "f") {
            std::vector<uint32_t> poly;
            std::string vtx;
            while (ss >> vtx) {
                Key k{0, 0, 0};
                int a = 0, b = 0, c = 0;
                if (std::sscanf(vtx.c_str(), "%d/%d/%d", &a, &b, &c) == 3) {}
                else if (std::sscanf(vtx.c_str(), "%d//%d", &a, &c) == 2) { b = 0; }
                else if (std::sscanf(vtx.c_str(), "%d/%d", &a, &b) == 2) { c = 0; }
                else { std::sscanf(vtx.c_str(), "%d", &a); }
                k.p = a < 0 ? (int)P.size() + a : a - 1;
                k.t = b == 0 ? -1 : (b < 0 ? (int)T.size() + b : b - 1);
                k.n = c == 0 ? -1 : (c < 0 ? (int)N.size() + c : c - 1);
                if (k.p < 0 || k.p >= (int)P.size()) return false;
                auto it = unique.find(k);
                uint32_t id;
                if (it == unique.end()) {
                    Vertex v;
                    v.position = P[k.p];
                    v.normal = (k.n >= 0 && k.n < (int)N.size()) ? N[k.n] : glm::vec3(0, 1, 0);
                    v.color = glm::vec3(1.0f);
                    if (k.t >= 0 && k.t < (int)T.size()) v.uv = glm::vec2(T[k.t].x, 1.0f - T[k.t].y);
                    id = (uint32_t)_vertices.size();
                    _vertices.push_back(v);
                    unique[k] = id;
                } else id = it->second;
                poly.push_back(id);
            }
            for (size_t i = 2; i < poly.size(); i++) { _indices.push_back(poly[0]); _indices.push_back(poly[i - 1]); _indices.push_back(poly[i]); }
        }
    }
    return !_indices.empty();
}

// vk_mesh.cpp:161-184
Mesh* Mesh::get_quad()
{
    Mesh*& slot = _loadedMeshes["quad"];
    if (!slot) {
        Mesh* mesh = new Mesh();
        const float q[4][5] = {{1, 1, 1, 1, 0}, {-1, 1, 0, 1, 0}, {-1, -1, 0, 0, 0}, {1, -1, 1, 0, 0}};
        for (auto& r : q) { Vertex v; v.position = glm::vec3(r[0], r[1], 0.0f); v.normal = glm::vec3(0, 0, 1); v.color = glm::vec3(1.0f); v.uv = glm::vec2(r[2], r[3]); mesh->_vertices.push_back(v); }
        mesh->_indices = {0, 1, 2, 2, 3, 0};
        slot = mesh;
    }
    return slot;
}

// vk_mesh.cpp:185-206
Mesh* Mesh::get_triangle()
{
    Mesh*& slot = _loadedMeshes["triangle"];
    if (!slot) {
        Mesh* mesh = new Mesh();
        const float t[3][4] = {{1, 1, 1, 1}, {-1, 1, 0, 1}, {0, -1, 1, 0}};
        for (auto& r : t) { Vertex v; v.position = glm::vec3(r[0], r[1], 0.0f); v.normal = glm::vec3(0, 0, 1); v.color = glm::vec3(1.0f); v.uv = glm::vec2(r[2], r[3]); mesh->_vertices.push_back(v); }
        mesh->_indices = {0, 1, 2};
        slot = mesh;
    }
    return slot;
}

// vk_mesh.cpp:208-263: 24 vertices (4 per face, face order -z +z +y -y -x +x), every uv = (1,1)
Mesh* Mesh::get_cube()
{
    Mesh*& slot = _loadedMeshes["cube"];
    if (!slot) {
        Mesh* mesh = new Mesh();
        struct Face { float n[3]; float c[4][3]; };
        const Face faces[6] = {
            {{0, 0, -1}, {{1, 1, -1}, {-1, 1, -1}, {-1, -1, -1}, {1, -1, -1}}},
            {{0, 0, 1}, {{1, 1, 1}, {-1, 1, 1}, {-1, -1, 1}, {1, -1, 1}}},
            {{0, 1, 0}, {{1, 1, 1}, {-1, 1, 1}, {-1, 1, -1}, {1, 1, -1}}},
            {{0, -1, 0}, {{1, -1, 1}, {-1, -1, 1}, {-1, -1, -1}, {1, -1, -1}}},
            {{-1, 0, 0}, {{-1, 1, 1}, {-1, -1, 1}, {-1, -1, -1}, {-1, 1, -1}}},
            {{1, 0, 0}, {{1, 1, 1}, {1, -1, 1}, {1, -1, -1}, {1, 1, -1}}}};
        for (int f = 0; f < 6; f++) {
            for (int k = 0; k < 4; k++) {
                Vertex v; v.position = glm::vec3(faces[f].c[k][0], faces[f].c[k][1], faces[f].c[k][2]);
                v.normal = glm::vec3(faces[f].n[0], faces[f].n[1], faces[f].n[2]); v.color = glm::vec3(1.0f); v.uv = glm::vec2(1.0f, 1.0f);
                mesh->_vertices.push_back(v);
            }
            uint32_t b = 4u * (uint32_t)f;
            const uint32_t q[6] = {b, b + 1, b + 2, b + 2, b + 3, b};
            mesh->_indices.insert(mesh->_indices.end(), q, q + 6);
        }
        slot = mesh;
    }
    return slot;
}

// Mesh::upload lives in renderer.cpp: it is the only member that talks to libvkp, and the scene library must not depend on it

// ---- Node / Prefab (src/vk_mesh.cpp:361-546, :831-912) ---------------------------------------------------
glm::mat4 Node::getGlobalMatrix(bool fast)
{
    if (_parent) _global_matrix = (fast ? _parent->_global_matrix : _parent->getGlobalMatrix()) * _matrix;
    else _global_matrix = _matrix;
    return _global_matrix;
}

void Node::node_to_instance(std::vector<TlasInstance>& instances, int& index, glm::mat4 model)
{
    if (!_primitives.empty()) {
        glm::mat4 matrix = model * getGlobalMatrix(false);
        for (auto& prim : _primitives) {
            TlasInstance instance{};
            instance.transform = matrix; instance.instanceId = (uint32_t)index; instance.blasId = (uint32_t)index;
            instances.emplace_back(instance);
            prim->instanceID = index;
            index++;
        }
    }
    for (Node* child : _children) child->node_to_instance(instances, index, model);
}

unsigned int Node::get_number_nodes()
{
    unsigned int count = _primitives.empty() ? 0u : 1u;
    for (Node* child : _children) count += child->get_number_nodes();
    return count;
}

void Node::fill_matrix_buffer(std::vector<glm::mat4>& buffer, const glm::mat4 model)
{
    if (!_primitives.empty()) {
        glm::mat4 matrix = model * getGlobalMatrix();
        for (auto& prim : _primitives) prim->transformID = (int32_t)buffer.size();
        buffer.push_back(matrix);
    }
    for (Node* n : _children) n->fill_matrix_buffer(buffer, model);
}

void Node::fill_index_buffer(std::vector<glm::vec4>& buffer)
{
    for (const auto& prim : _primitives)
        buffer.push_back(glm::vec4((float)prim->instanceID, (float)prim->materialID, (float)prim->transformID, (float)prim->firstIndex));
    for (Node* n : _children) n->fill_index_buffer(buffer);
}

void Node::addMaterial(Material* mat)
{
    if (Material::exists(mat)) _primitives[0]->materialID = Material::getIndex(mat);
    else { Material::_materials.push_back(mat); _primitives[0]->materialID = (int32_t)Material::_materials.size() - 1; }
}

std::unordered_map<std::string, Prefab*> Prefab::_prefabsMap;

static Prefab* make_single_primitive_prefab(const std::string& name, Mesh* mesh)
{
    // Prefab::createOBJprefab (vk_mesh.cpp:831-841)
    Prefab* prefab = new Prefab();
    prefab->_name = name; prefab->_mesh = mesh;
    Node* node = new Node();
    Primitive* p = new Primitive();
    p->indexCount = mesh ? (uint32_t)mesh->_indices.size() : 0u;
    p->vertexCount = mesh ? (uint32_t)mesh->_vertices.size() : 0u;
    p->materialID = Material::setDefaultMaterial();
    node->_primitives.push_back(p);
    prefab->_root.push_back(node);
    return prefab;
}

Prefab* Prefab::GET(std::string filename, const bool invertNormals)
{
    std::string name = base_name(filename);
    Prefab*& slot = _prefabsMap[name];
    if (slot) return slot;
    const bool gltf = name.find(".gltf") != std::string::npos, glb = name.find(".glb") != std::string::npos;
    if (gltf || glb) {
        // the real asset when it is on the search paths (vk_mesh.cpp:857-896) ...
        std::string path = find_file(filename);
        if (!path.empty()) {
            slot = load_gltf_prefab(path, glb && !gltf, invertNormals);
            slot->_name = name;
            Mesh*& owned = Mesh::_loadedMeshes[name];   // so that reset_registries frees it
            delete owned;
            owned = slot->_mesh;
            return slot;
        }
    }
    // ... else the deterministic stand-in (the reference's assets are git-ignored)
    Mesh* mesh = Mesh::GET(filename.c_str());
    if (!mesh) return nullptr;
    slot = make_single_primitive_prefab(name, mesh);
    if (gltf || glb) {
        // stand-in for the glTF material import (vk_mesh.cpp:777-829): one textured PBR material
        std::vector<uint8_t> a, e, mr, nm;
        proc::helmet_textures(512, 512, a, e, mr, nm);
        Material* m = new Material();
        m->diffuseTexture = Texture::add((name + ":albedo").c_str(), 512, 512, std::move(a));
        m->emissiveTexture = Texture::add((name + ":emissive").c_str(), 512, 512, std::move(e));
        m->metallicRoughnessTexture = Texture::add((name + ":metalRough").c_str(), 512, 512, std::move(mr));
        m->normalTexture = Texture::add((name + ":normal").c_str(), 512, 512, std::move(nm));
        slot->_root[0]->addMaterial(m);
    }
    return slot;
}

// the reference does not cache these (vk_mesh.cpp:903-912): every call makes a new prefab
Prefab* Prefab::GET(std::string name, Mesh* mesh) { return make_single_primitive_prefab(name, mesh); }

// ---- Entity (src/entity.cpp) -----------------------------------------------------------------------------
Object::Object(glm::vec3 position, Mesh* mesh, Material* material_) : material(material_)
{
    prefab = new Prefab(); prefab->_mesh = mesh;
    m_matrix = glm::translate(glm::mat4(1), position);
}

Light::Light(lightType type_, glm::vec3 color_, glm::vec3 position_, float intensity_, float maxDistance_, float radius_)
    : type{type_}, color{color_}, position{position_}, intensity{intensity_}, maxDistance{maxDistance_}, radius{radius_}
{
    m_matrix = glm::translate(glm::mat4(1), position);
    m_selected = false; m_visible = true;
}
void Light::update() { position = (glm::vec3)m_matrix[3]; }

// ---- Scene (src/scene.cpp) ----------------------------------------------------------------------------------
void Scene::reset_registries()
{
    for (Material* m : Material::_materials) delete m;
    Material::_materials.clear();
    for (Texture* t : Texture::_textures) delete t;
    Texture::_textures.clear(); Texture::_ids.clear();
    for (auto& kv : Mesh::_loadedMeshes) delete kv.second;
    Mesh::_loadedMeshes.clear();
    Prefab::_prefabsMap.clear();
}

unsigned int Scene::get_drawable_nodes_size()
{
    unsigned int count = 0;
    for (Object* o : _entities) for (Node* root : o->prefab->_root) count += root->get_number_nodes();
    return count;
}

void Scene::create_scene(int i)
{
    g_detail = _detail;
    switch (i) {
    case 0: default_scene(); break;
    case 1: cornell_scene(); break;
    case 2: stress_scene(); break;
    case 3: ring_lights_scene(); break;
    default: break;
    }
    finish_environment();
}

// environment (renderer.cpp:3073-3074 binds an 8k sky JPG and an .hdr irradiance map, both 8-bit)
void Scene::finish_environment()
{
    // The real files are used when they are on the search paths, else the analytic stand-ins.
    _skyTexture = Texture::get_id("LA_Downtown_Helipad_GoldenHour_8k.jpg");
    _irradianceTexture = Texture::get_id("LA_Downtown_Helipad_GoldenHour_Env.hdr");
    std::vector<uint8_t> sky, irr;
    if (_skyTexture < 0) {
        if (Texture::get_id("sky") < 0) { proc::sky_texture(2048, 1024, sky); Texture::add("sky", 2048, 1024, std::move(sky)); }
        _skyTexture = Texture::get_id("sky");
    }
    if (_irradianceTexture < 0) {
        if (Texture::get_id("irradiance") < 0) { proc::irradiance_texture(64, 32, irr); Texture::add("irradiance", 64, 32, std::move(irr)); }
        _irradianceTexture = Texture::get_id("irradiance");
    }
}

static Object* make_object(Prefab* prefab, const glm::mat4& m)
{
    Object* o = new Object();
    delete o->prefab;
    o->prefab = prefab; o->m_matrix = m;
    o->material = Material::_materials[prefab->_root[0]->_primitives[0]->materialID];
    return o;
}

// Not in the reference: its default lights, floor and camera around one asset file, for the headless CLI (--prefab)
void Scene::create_viewer_scene(const std::string& prefabFile)
{
    g_detail = _detail;
    _camera = new Camera(glm::vec3(0, 5, 10));
    Light* a = new Light(); a->m_matrix = glm::translate(glm::mat4(1), glm::vec3(10, 12, -5)); a->color = glm::vec3{1.0f, 0.8f, 0.5f}; a->intensity = 500.0f; a->radius = 0.1f;
    Light* b = new Light(); b->m_matrix = glm::translate(glm::mat4(1), glm::vec3(-10, 15, -5)); b->color = glm::vec3{0.5f, 1.0f, 1.0f}; b->intensity = 250.0f; b->radius = 0.2f;
    Light* c = new Light(); c->m_matrix = glm::translate(glm::mat4(1), glm::vec3(0, 15, 5)); c->intensity = 500.f; c->radius = 0.09f;
    _lights = {a, b, c};
    for (Light* l : _lights) l->update();
    Material* m_floor = new Material(); m_floor->metallicFactor = 0.1f;
    Prefab* p_quad = Prefab::GET("quad", Mesh::get_quad()); p_quad->_root[0]->addMaterial(m_floor);
    if (vkutil_find_file(prefabFile).empty()) throw std::runtime_error("File not found " + prefabFile);   // vk_utils.cpp:48-50
    Prefab* prefab = Prefab::GET(prefabFile);
    if (!prefab) throw std::runtime_error("No valid filename " + prefabFile);                            // vk_mesh.cpp:863-866
    Object* floor = make_object(p_quad, glm::translate(glm::mat4(1), glm::vec3(0, 0, -5)) *
                                            glm::rotate(glm::mat4(1), glm::radians(-90.0f), glm::vec3(1, 0, 0)) * glm::scale(glm::mat4(1), glm::vec3(15)));
    Object* object = make_object(prefab, glm::translate(glm::mat4(1), glm::vec3(0, 1, -5)));
    _entities = {floor, object};
    finish_environment();
}

// scene.cpp:34-160
void Scene::default_scene()
{
    _camera = new Camera(glm::vec3(0, 5, 10));

    Light* light = new Light();
    light->m_matrix = glm::translate(glm::mat4(1), glm::vec3(10, 12, -5));
    light->color = glm::vec3{1.0f, 0.8f, 0.5f}; light->intensity = 500.0f; light->radius = 0.1f;
    Light* light2 = new Light();
    light2->m_matrix = glm::translate(glm::mat4(1), glm::vec3(-10, 15, -5));
    light2->color = glm::vec3{0.5f, 1.0f, 1.0f}; light2->intensity = 250.0f; light2->radius = 0.2f;
    Light* light3 = new Light();
    light3->m_matrix = glm::translate(glm::mat4(1), glm::vec3(0, 15, 5));
    light3->intensity = 500.f; light3->radius = 0.09f;
    _lights = {light, light2, light3};

    Material* m_mirror = new Material(); m_mirror->shadingModel = 3; m_mirror->metallicFactor = 1.f;
    Material* m_glass = new Material(); m_glass->diffuseColor = glm::vec4{0.7f, 0.7f, 1.0f, 1}; m_glass->shadingModel = 4; m_glass->ior = 1.125f; m_glass->metallicFactor = 1.f;
    Material* m_gold = new Material(); m_gold->diffuseColor = glm::vec4{1.0f, 0.71f, 0.29f, 1.0f}; m_gold->metallicFactor = 0.5f; m_gold->roughnessFactor = 0.1f;
    Material* m_floor = new Material(); m_floor->metallicFactor = 0.1f;

    Prefab* p_quad = Prefab::GET("quad", Mesh::get_quad());             p_quad->_root[0]->addMaterial(m_floor);
    Prefab* p_mirror = Prefab::GET("cube", Mesh::get_cube());           p_mirror->_root[0]->addMaterial(m_mirror);
    Prefab* p_sphere_mirror = Prefab::GET("sphere.obj");                p_sphere_mirror->_root[0]->addMaterial(m_mirror);
    Prefab* p_glass_sphere = Prefab::GET("Glass Sphere", Mesh::GET("sphere.obj")); p_glass_sphere->_root[0]->addMaterial(m_glass);
    Prefab* p_red_sphere = Prefab::GET("Red Sphere", Mesh::GET("sphere.obj"));     p_red_sphere->_root[0]->addMaterial(m_mirror);
    Prefab* p_helmet = Prefab::GET("DamagedHelmet.gltf");
    Prefab* p_lucy = Prefab::GET("lucy", Mesh::GET("lucy.obj"));        p_lucy->_root[0]->addMaterial(m_gold);

    Object* sphere = make_object(p_red_sphere, glm::translate(glm::mat4(1), glm::vec3(5, 1, -5)));
    Object* sphere3 = make_object(p_glass_sphere, glm::translate(glm::mat4(1), glm::vec3(0, 1, -5)));
    Object* floor = make_object(p_quad, glm::translate(glm::mat4(1), glm::vec3(0, 0, -5)) *
                                            glm::rotate(glm::mat4(1), glm::radians(-90.0f), glm::vec3(1, 0, 0)) *
                                            glm::scale(glm::mat4(1), glm::vec3(15)));
    Object* mirror = make_object(p_mirror, glm::translate(glm::mat4(1), glm::vec3(0, 4, -10)) * glm::scale(glm::mat4(1), glm::vec3(4, 4, 1)));
    Object* helmet = make_object(p_helmet, glm::translate(glm::mat4(1), glm::vec3(-5, 1, -5)));
    Object* lucy = make_object(p_lucy, glm::scale(glm::mat4(1), glm::vec3(0.01f)));

    _entities = {floor, sphere, sphere3, mirror, lucy, helmet};   // TLAS order, scene.cpp:151-158
}

// scene.cpp:162-272
void Scene::cornell_scene()
{
    _camera = new Camera(glm::vec3(0, 5, 10));
    Light* light = new Light();
    light->m_matrix = glm::translate(glm::mat4(1), glm::vec3(0, 8.5f, -5));
    light->color = glm::vec3{1.0f, 0.8f, 0.5f}; light->intensity = 250.0f; light->radius = 0.1f;
    _lights = {light};

    Material* m_white = new Material(); m_white->diffuseColor = glm::vec4(1, 1, 1, 1);
    Material* m_red = new Material(); m_red->diffuseColor = glm::vec4(1, 0, 0, 1);
    Material* m_green = new Material(); m_green->diffuseColor = glm::vec4(0, 1, 0, 1);
    Material* m_mirror = new Material(); m_mirror->shadingModel = 3; m_mirror->metallicFactor = 1.f;
    Material* m_glass = new Material(); m_glass->diffuseColor = glm::vec4{0.7f, 0.7f, 1.0f, 1}; m_glass->shadingModel = 4; m_glass->ior = 1.125f; m_glass->metallicFactor = 1.f;

    Prefab* p_quad = Prefab::GET("quad", Mesh::get_quad());                 p_quad->_root[0]->addMaterial(m_white);
    Prefab* p_white_quad = Prefab::GET("white_quad", Mesh::get_cube());     p_white_quad->_root[0]->addMaterial(m_white);
    Prefab* p_red_quad = Prefab::GET("red_quad", Mesh::get_cube());         p_red_quad->_root[0]->addMaterial(m_red);
    Prefab* p_green_quad = Prefab::GET("quad", Mesh::get_cube());           p_green_quad->_root[0]->addMaterial(m_green);
    Prefab* p_helmet = Prefab::GET("DamagedHelmet.gltf");
    Prefab* p_mirror_sphere = Prefab::GET("mirror_sphere", Mesh::GET("sphere.obj")); p_mirror_sphere->_root[0]->addMaterial(m_mirror);
    Prefab* p_glass_sphere = Prefab::GET("glass_sphere", Mesh::GET("sphere.obj"));   p_glass_sphere->_root[0]->addMaterial(m_glass);
    (void)p_quad;

    auto T = [](float x, float y, float z) { return glm::translate(glm::mat4(1), glm::vec3(x, y, z)); };
    auto S = [](float x, float y, float z) { return glm::scale(glm::mat4(1), glm::vec3(x, y, z)); };
    _entities = {
        make_object(p_white_quad, T(0, 0, -5) * S(5, 0.1f, 5)),        // floor
        make_object(p_white_quad, T(0, 5, -10) * S(5, 5, 0.1f)),       // back wall
        make_object(p_red_quad, T(-5, 5, -5) * S(0.1f, 5, 5)),         // left wall
        make_object(p_green_quad, T(5, 5, -5) * S(0.1f, 5, 5)),        // right wall
        make_object(p_white_quad, T(0, 10, -5) * S(5, 0.1f, 5)),       // ceiling
        make_object(p_glass_sphere, T(-3.f, 2.2f, -2.5f) * S(2, 2, 2)),
        make_object(p_mirror_sphere, T(3.5f, 1.f, -2.5f)),
        make_object(p_helmet, T(0, 2.5f, -5)),
    };
}

// BASELINE.json configs[3] / SURVEY 8d cfg4: one displaced icosphere x N instances on a jittered grid + floor
void Scene::stress_scene()
{
    _camera = new Camera(glm::vec3(0, 9, 24));
    _camera->rotate(0.0f, -150.0f);   // pitch -15 degrees
    Light* l0 = new Light(); l0->m_matrix = glm::translate(glm::mat4(1), glm::vec3(20, 30, 10)); l0->color = glm::vec3(1.0f, 0.9f, 0.8f); l0->intensity = 2500.0f; l0->radius = 0.05f;
    Light* l1 = new Light(); l1->m_matrix = glm::translate(glm::mat4(1), glm::vec3(-25, 25, 0)); l1->color = glm::vec3(0.6f, 0.8f, 1.0f); l1->intensity = 1500.0f; l1->radius = 0.1f;
    Light* l2 = new Light(); l2->m_matrix = glm::translate(glm::mat4(1), glm::vec3(0, 35, -30)); l2->intensity = 2500.0f; l2->radius = 0.08f;
    _lights = {l0, l1, l2};

    Material* m_floor = new Material(); m_floor->metallicFactor = 0.1f;
    Prefab* p_floor = Prefab::GET("quad", Mesh::get_quad()); p_floor->_root[0]->addMaterial(m_floor);
    _entities.push_back(make_object(p_floor, glm::translate(glm::mat4(1), glm::vec3(0, 0, -8)) *
                                                 glm::rotate(glm::mat4(1), glm::radians(-90.0f), glm::vec3(1, 0, 0)) * glm::scale(glm::mat4(1), glm::vec3(60))));
    Material* mats[5];
    for (int k = 0; k < 5; k++) {
        mats[k] = new Material();
        const int modes[5] = {0, 0, 0, 3, 4};
        mats[k]->shadingModel = modes[k];
        if (k == 0) { mats[k]->diffuseColor = glm::vec4(0.9f, 0.3f, 0.25f, 1); mats[k]->metallicFactor = 0.2f; mats[k]->roughnessFactor = 0.4f; }
        if (k == 1) { mats[k]->diffuseColor = glm::vec4(1.0f, 0.71f, 0.29f, 1); mats[k]->metallicFactor = 0.5f; mats[k]->roughnessFactor = 0.1f; }
        if (k == 2) { mats[k]->diffuseColor = glm::vec4(0.3f, 0.7f, 0.4f, 1); mats[k]->metallicFactor = 0.0f; mats[k]->roughnessFactor = 0.8f; }
        if (k == 4) { mats[k]->diffuseColor = glm::vec4(0.7f, 0.7f, 1.0f, 1); mats[k]->ior = 1.125f; }
    }
    Mesh* blob = Mesh::GET("stress_blob");
    uint32_t rng = 1234u;
    auto rnd01 = [&rng]() { rng = 1664525u * rng + 1013904223u; return (float)(rng >> 8) / 16777216.0f; };
    const int n = g_detail.stress_instances;
    int cols = 1; while (cols * cols < n) cols++;
    for (int i = 0; i < n; i++) {
        Prefab* p = Prefab::GET("blob" + std::to_string(i), blob);
        p->_root[0]->addMaterial(mats[i % 5]);
        float gx = (float)(i % cols) - 0.5f * (float)(cols - 1), gz = (float)(i / cols) - 0.5f * (float)(cols - 1);
        float s = 0.5f + rnd01();                                           // scale in [0.5, 1.5)
        glm::vec3 pos(gx * 4.2f + (rnd01() - 0.5f) * 1.2f, 1.6f * s, -8.0f + gz * 4.2f + (rnd01() - 0.5f) * 1.2f);
        glm::vec3 axis = glm::normalize(glm::vec3(rnd01() - 0.5f, rnd01() - 0.5f, rnd01() - 0.5f) + glm::vec3(1e-3f, 0, 0));
        glm::mat4 m = glm::translate(glm::mat4(1), pos) * glm::rotate(glm::mat4(1), rnd01() * 6.2831853f, axis) * glm::scale(glm::mat4(1), glm::vec3(1.5f * s));
        _entities.push_back(make_object(p, m));
    }
}

// BASELINE.json configs[4] / SURVEY 8d cfg5: default scene lit by 16 point lights on a ring
void Scene::ring_lights_scene()
{
    default_scene();
    for (Light* l : _lights) delete l;
    _lights.clear();
    uint32_t rng = 1234u;
    auto rnd01 = [&rng]() { rng = 1664525u * rng + 1013904223u; return (float)(rng >> 8) / 16777216.0f; };
    const double PI = 3.14159265358979323846;
    for (int i = 0; i < 16; i++) {
        double a = 2.0 * PI * (double)i / 16.0;
        Light* l = new Light();
        l->m_matrix = glm::translate(glm::mat4(1), glm::vec3((float)(12.0 * std::cos(a)), 12.0f, (float)(-5.0 + 12.0 * std::sin(a))));
        l->color = glm::vec3(0.4f + 0.6f * rnd01(), 0.4f + 0.6f * rnd01(), 0.4f + 0.6f * rnd01());
        l->intensity = 150.0f; l->radius = 0.2f;
        _lights.push_back(l);
    }
}

// ---- flat buffers ---------------------------------------------------------------------------------------------
// What Renderer::load_data_to_gpu (renderer.cpp:1477-1530), create_hybrid_descriptors (:2917-3129) and
// VulkanEngine::update (vk_engine.cpp:146-187) put in GPU buffers, as plain arrays.
SceneBuffers build_scene_buffers(Scene* scene)
{
    SceneBuffers B;
    std::unordered_map<Mesh*, int> meshIds;
    scene->_matricesVector.clear();
    int index = 0;
    for (Object* obj : scene->_entities) {
        if (!obj->prefab || obj->prefab->_root.empty()) continue;
        Mesh* mesh = obj->prefab->_mesh;
        if (!meshIds.count(mesh)) { meshIds[mesh] = (int)B.meshes.size(); B.meshes.push_back(mesh); }
        std::vector<TlasInstance> tl;
        for (Node* root : obj->prefab->_root) {
            root->node_to_instance(tl, index, obj->m_matrix);
            root->fill_matrix_buffer(scene->_matricesVector, obj->m_matrix);
        }
        // one vkp_instance per primitive, in the same running order as node_to_instance
        size_t k = 0;
        std::function<void(Node*)> walk = [&](Node* n) {
            for (Primitive* p : n->_primitives) {
                vkp_instance inst{};
                inst.mesh_id = meshIds[mesh];
                inst.first_index = p->firstIndex; inst.index_count = p->indexCount; inst.material_id = p->materialID;
                glm::mat4 inv = glm::inverse(tl[k].transform);
                std::memcpy(inst.model, tl[k].transform.data(), 64);
                std::memcpy(inst.inv_model, inv.data(), 64);
                B.instances.push_back(inst);
                k++;
            }
            for (Node* c : n->_children) walk(c);
        };
        for (Node* root : obj->prefab->_root) walk(root);
    }
    for (Material* m : Material::_materials) {
        GPUMaterial g = m->materialToShader();
        vkp_material vm;
        std::memcpy(&vm, &g, sizeof(vm));
        B.materials.push_back(vm);
    }
    B.lights = build_lights(scene);
    B.textures = Texture::_textures;
    B.sky_tex = scene->_skyTexture; B.irradiance_tex = scene->_irradianceTexture;
    return B;
}

std::vector<vkp_light> build_lights(Scene* scene)
{
    std::vector<vkp_light> out;
    for (Light* l : scene->_lights) {
        l->update();
        vkp_light u{};
        u.color[0] = l->color.x; u.color[1] = l->color.y; u.color[2] = l->color.z; u.color[3] = l->intensity;
        u.position[0] = l->position.x; u.position[1] = l->position.y; u.position[2] = l->position.z;
        if (l->type == DIRECTIONAL_LIGHT) u.position[3] = -1.0f;       // radius is left unset upstream (vk_engine.cpp:153-156)
        else { u.position[3] = l->maxDistance; u.radius = l->radius; }
        out.push_back(u);
    }
    return out;
}

// VulkanEngine::updateCameraMatrices (vk_engine.cpp:688-741).  prev == nullptr -> static camera: prev := current,
// the defined behaviour for the reference's NaN motion vectors.
vkp_camera build_camera(Camera* cam, float ratio, const glm::mat4* prevView, const glm::mat4* prevProj)
{
    vkp_camera c{};
    glm::mat4 view = cam->getView();
    glm::mat4 projection = cam->getProjection(ratio);
    projection[1][1] *= -1;
    glm::mat4 invView = glm::inverse(view), invProj = glm::inverse(projection);
    std::memcpy(c.view, view.data(), 64); std::memcpy(c.proj, projection.data(), 64);
    std::memcpy(c.prev_view, (prevView ? *prevView : view).data(), 64);
    std::memcpy(c.prev_proj, (prevProj ? *prevProj : projection).data(), 64);
    std::memcpy(c.inv_view, invView.data(), 64); std::memcpy(c.inv_proj, invProj.data(), 64);
    c.position[0] = cam->_position.x; c.position[1] = cam->_position.y; c.position[2] = cam->_position.z; c.position[3] = 1.0f;
    return c;
}
