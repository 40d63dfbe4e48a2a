// gltf.cpp -- glTF 2.0 import for Prefab::GET("x.gltf" / "x.glb") (src/vk_mesh.cpp:843-901), without tinygltf
// (un-vendored upstream, .gitignore:7): node hierarchy, one Primitive per glTF primitive appended to one shared Mesh,
// metallic-roughness materials and their four textures.  The walk follows the reference line by line, quirks included:
//   * Prefab::loadNode (vk_mesh.cpp:646-775) loads a node's children BEFORE the node's own primitives, so vertex / index
//     order in the shared mesh is children-first; the recursive call drops `invertNormals` (:657), i.e. only the scene's
//     root nodes get flipped normals;
//   * get_local_matrix (:611-631): T * R * S built with glm::translate / mat4(quat) / glm::scale, then replaced by
//     `matrix` when the node has one;
//   * vertices: position, normalize(normal) (NaN when the primitive has no NORMAL -- glm::normalize of the zero vector),
//     uv = TEXCOORD_0 unflipped or (0,0), white colour (:700-709); indices are rebased by firstVertex (:713-752);
//   * Prefab::loadMaterial (:777-803): a new Material per primitive, de-duplicated by value against the registry
//     BEFORE its texture slots are rewritten to registry ids; baseColorFactor / metallicFactor / roughnessFactor with
//     tinygltf's defaults (1,1,1,1 / 1 / 1), shading model 0;
//   * Prefab::loadTextures (:805-829) indexes `images[]` with the glTF *texture* index (not textures[i].source) and runs
//     again for every primitive that maps to the same material; order of registration: diffuse, normal, emissive,
//     metallic-roughness.  We keep that indexing (identical to the reference whenever textures[i].source == i, which
//     holds for DamagedHelmet.gltf) and skip a slot instead of reading out of bounds.
// Beyond the reference (it would read garbage there): accessor byteStride is honoured, TEXCOORD_0 may be normalised
// u8 / u16, primitives without indices get 0..n-1, images may be embedded (bufferView or data: URI), buffers may be data: URIs.
#include "vkpinut.h"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>

#include "image_decode.h"
#include "json_min.h"

namespace {

using jsonmin::Value;

struct Model {
    Value doc;
    std::vector<std::vector<uint8_t>> buffers;
    std::string dir;    // directory of the .gltf, with trailing '/', or ""
    std::string name;   // resolved file name
};

[[noreturn]] void gfail(const Model& M, const std::string& what) { throw std::runtime_error("glTF " + M.name + ": " + what); }

bool read_file(const std::string& path, std::vector<uint8_t>& out)
{
    std::ifstream f(path, std::ios::binary);
    if (!f.good()) return false;
    f.seekg(0, std::ios::end);
    std::streamoff n = f.tellg();
    f.seekg(0, std::ios::beg);
    out.resize((size_t)n);
    if (n > 0) f.read(reinterpret_cast<char*>(out.data()), n);
    return f.good() || f.eof();
}

bool is_data_uri(const std::string& uri) { return uri.compare(0, 5, "data:") == 0; }

std::vector<uint8_t> decode_data_uri(const Model& M, const std::string& uri)
{
    size_t comma = uri.find(',');
    if (comma == std::string::npos || uri.substr(0, comma).find(";base64") == std::string::npos) gfail(M, "unsupported data: URI (base64 expected)");
    std::vector<uint8_t> out;
    unsigned acc = 0;
    int bits = 0;
    for (size_t i = comma + 1; i < uri.size(); i++) {
        char c = uri[i];
        int v;
        if (c >= 'A' && c <= 'Z') v = c - 'A';
        else if (c >= 'a' && c <= 'z') v = c - 'a' + 26;
        else if (c >= '0' && c <= '9') v = c - '0' + 52;
        else if (c == '+' || c == '-') v = 62;
        else if (c == '/' || c == '_') v = 63;
        else if (c == '=') break;
        else continue;
        acc = (acc << 6) | (unsigned)v;
        bits += 6;
        if (bits >= 8) { bits -= 8; out.push_back((uint8_t)((acc >> bits) & 0xff)); }
    }
    return out;
}

std::string percent_decode(const std::string& s)
{
    std::string out;
    for (size_t i = 0; i < s.size(); i++) {
        if (s[i] == '%' && i + 2 < s.size() && std::isxdigit((unsigned char)s[i + 1]) && std::isxdigit((unsigned char)s[i + 2])) {
            out += (char)std::strtol(s.substr(i + 1, 2).c_str(), nullptr, 16);
            i += 2;
        } else out += s[i];
    }
    return out;
}

const Value& element(const Model& M, const char* array, int index)
{
    const Value* a = M.doc.find(array);
    if (!a || !a->is_array() || index < 0 || (size_t)index >= a->arr.size()) gfail(M, std::string(array) + "[" + std::to_string(index) + "] does not exist");
    return a->arr[(size_t)index];
}
size_t count_of(const Model& M, const char* array) { const Value* a = M.doc.find(array); return a && a->is_array() ? a->arr.size() : 0; }

struct View {
    const uint8_t* data{nullptr};
    size_t count{0}, stride{0};
    int componentType{0}, ncomp{0};
    bool normalized{false};
};

int component_size(int t) { return t == 5120 || t == 5121 ? 1 : t == 5122 || t == 5123 ? 2 : t == 5125 || t == 5126 ? 4 : 0; }
int type_components(const std::string& t)
{
    if (t == "SCALAR") return 1;
    if (t == "VEC2") return 2;
    if (t == "VEC3") return 3;
    if (t == "VEC4") return 4;
    if (t == "MAT2") return 4;
    if (t == "MAT3") return 9;
    if (t == "MAT4") return 16;
    return 0;
}

// a JSON number used as a byte offset / count: a non-negative integer below 2^53, anything else is a malformed file
size_t json_size(const Model& M, const Value& obj, const char* key, double dflt = 0)
{
    const double d = obj.get_num(key, dflt);
    if (!(d >= 0.0) || d > 9007199254740992.0 || d != std::floor(d)) gfail(M, std::string(key) + " is not a non-negative integer");
    return (size_t)d;
}

// accessor -> bufferView -> buffer, like vk_mesh.cpp:676-679 (data + accessor.byteOffset + view.byteOffset), bounds-checked
View accessor_view(const Model& M, int index)
{
    const Value& acc = element(M, "accessors", index);
    if (acc.find("sparse")) gfail(M, "sparse accessors are not supported");
    const int bv = acc.get_int("bufferView", -1);
    if (bv < 0) gfail(M, "accessor without bufferView");
    const Value& view = element(M, "bufferViews", bv);
    const int buf = view.get_int("buffer", -1);
    if (buf < 0 || (size_t)buf >= M.buffers.size()) gfail(M, "bufferView names a missing buffer");
    View v;
    v.componentType = acc.get_int("componentType", 0);
    v.ncomp = type_components(acc.get_str("type"));
    v.count = json_size(M, acc, "count");
    const Value* nrm = acc.find("normalized");
    v.normalized = nrm && nrm->type == Value::Bool && nrm->boolean;
    const int csz = component_size(v.componentType);
    if (csz == 0 || v.ncomp == 0) gfail(M, "accessor with unknown component type / type");
    const size_t elem = (size_t)csz * v.ncomp;
    const size_t byteStride = json_size(M, view, "byteStride");
    v.stride = byteStride ? byteStride : elem;
    const std::vector<uint8_t>& B = M.buffers[(size_t)buf];
    // the accessor must lie inside its bufferView, the bufferView inside its buffer; every comparison is written so that it cannot wrap
    const size_t view_off = json_size(M, view, "byteOffset"), acc_off = json_size(M, acc, "byteOffset");
    if (view_off > B.size()) gfail(M, "bufferView starts past the end of its buffer");
    size_t view_len = B.size() - view_off;
    if (view.find("byteLength")) {
        const size_t bl = json_size(M, view, "byteLength");
        if (bl > view_len) gfail(M, "bufferView reaches past the end of its buffer");
        view_len = bl;
    }
    if (v.count) {
        if (acc_off > view_len || elem > view_len - acc_off) gfail(M, "accessor reaches past the end of its bufferView");
        if (v.count - 1 > (view_len - acc_off - elem) / v.stride) gfail(M, "accessor reaches past the end of its bufferView");
    }
    v.data = B.data() + view_off + acc_off;
    return v;
}

inline float read_f32(const uint8_t* p) { float f; std::memcpy(&f, p, 4); return f; }

// glm::normalize for vec3: v * inversesqrt(dot(v, v)) with inversesqrt(x) = 1 / sqrt(x) (GLM without SIMD)
glm::vec3 glm_normalize(glm::vec3 v)
{
    float d = v.x * v.x + v.y * v.y + v.z * v.z;
    float inv = 1.0f / std::sqrt(d);
    return glm::vec3(v.x * inv, v.y * inv, v.z * inv);
}

// glm::mat4_cast(quat) (glm/gtc/quaternion.inl mat3_cast)
glm::mat4 mat4_from_quat(float x, float y, float z, float w)
{
    glm::mat4 R(1.0f);
    float qxx = x * x, qyy = y * y, qzz = z * z, qxz = x * z, qxy = x * y, qyz = y * z, qwx = w * x, qwy = w * y, qwz = w * z;
    R[0].x = 1.0f - 2.0f * (qyy + qzz); R[0].y = 2.0f * (qxy + qwz);        R[0].z = 2.0f * (qxz - qwy);
    R[1].x = 2.0f * (qxy - qwz);        R[1].y = 1.0f - 2.0f * (qxx + qzz); R[1].z = 2.0f * (qyz + qwx);
    R[2].x = 2.0f * (qxz + qwy);        R[2].y = 2.0f * (qyz - qwx);        R[2].z = 1.0f - 2.0f * (qxx + qyy);
    return R;
}

// Prefab::get_local_matrix (vk_mesh.cpp:611-631); tinygltf holds the numbers as double, GLM converts to float
glm::mat4 get_local_matrix(const Value& node)
{
    glm::mat4 matrix(1.0f);
    std::vector<double> t = node.get_numbers("translation"), r = node.get_numbers("rotation"), s = node.get_numbers("scale"), m = node.get_numbers("matrix");
    if (t.size() == 3) matrix = glm::translate(matrix, glm::vec3((float)t[0], (float)t[1], (float)t[2]));
    if (r.size() == 4) matrix = matrix * mat4_from_quat((float)r[0], (float)r[1], (float)r[2], (float)r[3]);
    if (s.size() == 3) matrix = glm::scale(matrix, glm::vec3((float)s[0], (float)s[1], (float)s[2]));
    if (m.size() == 16) for (int i = 0; i < 16; i++) matrix.data()[i] = (float)m[i];
    return matrix;
}

std::string resolve_image_path(const Model& M, const std::string& uri)
{
    std::string rel = percent_decode(uri);
    { std::ifstream f(M.dir + rel); if (f.good()) return M.dir + rel; }
    return vkutil_find_file(rel);   // Texture::GET resolves through the search paths (vk_textures.cpp:129)
}

// One image of the document -> id in Texture::_textures (-1: slot skipped)
int register_image(const Model& M, int imageIndex)
{
    if (imageIndex < 0 || (size_t)imageIndex >= count_of(M, "images")) return -1;   // the reference reads out of bounds here
    const Value& img = element(M, "images", imageIndex);
    const std::string uri = img.get_str("uri");
    int w = 0, h = 0;
    std::vector<uint8_t> rgba;
    std::string err;
    if (!uri.empty() && !is_data_uri(uri)) {
        std::string path = resolve_image_path(M, uri);
        if (path.empty()) throw std::runtime_error("Failed to load texture file: " + uri + " (not found next to " + M.name + " or on the search paths)");
        int id = Texture::get_id(path.c_str());   // loads on first use, like Texture::GET + get_id (vk_mesh.cpp:811-812)
        if (id < 0) throw std::runtime_error("Failed to load texture file: " + path);
        return id;
    }
    const std::string key = M.name + "#image" + std::to_string(imageIndex);
    int id = Texture::get_id(key.c_str());
    if (id >= 0) return id;
    if (!uri.empty()) {
        std::vector<uint8_t> bytes = decode_data_uri(M, uri);
        if (!imgdec::decode_rgba8(bytes.data(), bytes.size(), w, h, rgba, &err)) gfail(M, "image " + std::to_string(imageIndex) + ": " + err);
    } else {
        const int bv = img.get_int("bufferView", -1);
        if (bv < 0) gfail(M, "image without uri or bufferView");
        const Value& view = element(M, "bufferViews", bv);
        const int buf = view.get_int("buffer", -1);
        if (buf < 0 || (size_t)buf >= M.buffers.size()) gfail(M, "image bufferView names a missing buffer");
        const size_t off = json_size(M, view, "byteOffset"), len = json_size(M, view, "byteLength");
        if (off > M.buffers[(size_t)buf].size() || len > M.buffers[(size_t)buf].size() - off) gfail(M, "image bufferView reaches past the end of its buffer");
        if (!imgdec::decode_rgba8(M.buffers[(size_t)buf].data() + off, len, w, h, rgba, &err)) gfail(M, "image " + std::to_string(imageIndex) + ": " + err);
    }
    return Texture::add(key.c_str(), w, h, std::move(rgba));
}

int texture_index(const Value* info) { return info && info->is_object() ? info->get_int("index", -1) : -1; }

// Prefab::loadMaterial (vk_mesh.cpp:777-803)
int load_material(const Model& M, int index)
{
    Material* mat = new Material();
    if (index > -1) {
        const Value& tmat = element(M, "materials", index);
        const Value* pbr = tmat.find("pbrMetallicRoughness");
        std::vector<double> base = pbr ? pbr->get_numbers("baseColorFactor") : std::vector<double>();
        if (base.size() != 4) base = {1.0, 1.0, 1.0, 1.0};
        mat->diffuseColor = glm::vec4((float)base[0], (float)base[1], (float)base[2], (float)base[3]);
        mat->metallicFactor = (float)(pbr ? pbr->get_num("metallicFactor", 1.0) : 1.0);
        mat->roughnessFactor = (float)(pbr ? pbr->get_num("roughnessFactor", 1.0) : 1.0);
        mat->diffuseTexture = texture_index(pbr ? pbr->find("baseColorTexture") : nullptr);
        mat->metallicRoughnessTexture = texture_index(pbr ? pbr->find("metallicRoughnessTexture") : nullptr);
        mat->emissiveTexture = texture_index(tmat.find("emissiveTexture"));
        mat->normalTexture = texture_index(tmat.find("normalTexture"));
    }
    if (Material::exists(mat)) {
        int id = Material::getIndex(mat);
        delete mat;
        return id;
    }
    Material::_materials.push_back(mat);
    return (int)Material::_materials.size() - 1;
}

// Prefab::loadTextures (vk_mesh.cpp:805-829): slot values are used as indices into images[]
void load_textures(const Model& M, int materialIndex)
{
    Material* mat = Material::_materials[(size_t)materialIndex];
    if (mat->diffuseTexture > -1) mat->diffuseTexture = register_image(M, mat->diffuseTexture);
    if (mat->normalTexture > -1) mat->normalTexture = register_image(M, mat->normalTexture);
    if (mat->emissiveTexture > -1) mat->emissiveTexture = register_image(M, mat->emissiveTexture);
    if (mat->metallicRoughnessTexture > -1) mat->metallicRoughnessTexture = register_image(M, mat->metallicRoughnessTexture);
}

// Prefab::loadNode (vk_mesh.cpp:646-775)
void load_node(const Model& M, Prefab* prefab, const Value& tnode, Node* parent, bool invertNormals, int depth)
{
    if (depth > 256) gfail(M, "node hierarchy too deep (cycle?)");
    Node* node = new Node();
    node->_name = tnode.get_str("name");
    node->_matrix = get_local_matrix(tnode);

    if (const Value* children = tnode.find("children"))
        if (children->is_array())
            for (const Value& c : children->arr) load_node(M, prefab, element(M, "nodes", c.is_number() ? (int)c.number : -1), node, false, depth + 1);

    const int meshIndex = tnode.get_int("mesh", -1);
    if (meshIndex > -1) {
        Mesh* mesh = prefab->_mesh;
        const Value& tmesh = element(M, "meshes", meshIndex);
        const Value* prims = tmesh.find("primitives");
        for (size_t i = 0; prims && prims->is_array() && i < prims->arr.size(); i++) {
            const Value& tprim = prims->arr[i];
            const uint32_t firstIndex = (uint32_t)mesh->_indices.size();
            const uint32_t firstVertex = (uint32_t)mesh->_vertices.size();
            uint32_t indexCount = 0, vertexCount = 0;

            const Value* attrs = tprim.find("attributes");
            View pos, nrm, uv;
            bool havePos = false, haveNrm = false, haveUv = false;
            if (attrs && attrs->is_object()) {
                if (attrs->find("POSITION")) { pos = accessor_view(M, attrs->get_int("POSITION", -1)); havePos = true; }
                if (attrs->find("NORMAL")) { nrm = accessor_view(M, attrs->get_int("NORMAL", -1)); haveNrm = true; }
                if (attrs->find("TEXCOORD_0")) { uv = accessor_view(M, attrs->get_int("TEXCOORD_0", -1)); haveUv = true; }
            }
            if (havePos) {
                if (pos.componentType != 5126 || pos.ncomp != 3) gfail(M, "POSITION must be float VEC3");
                vertexCount = (uint32_t)pos.count;
            }
            if (haveNrm && (nrm.componentType != 5126 || nrm.ncomp != 3 || nrm.count < vertexCount)) gfail(M, "NORMAL must be float VEC3 with one entry per vertex");
            if (haveUv && (uv.ncomp != 2 || uv.count < vertexCount || !(uv.componentType == 5126 || uv.componentType == 5121 || uv.componentType == 5123)))
                gfail(M, "TEXCOORD_0 must be VEC2 (float or normalised u8 / u16) with one entry per vertex");
            mesh->_vertices.reserve(mesh->_vertices.size() + vertexCount);
            for (uint32_t v = 0; v < vertexCount; v++) {
                const uint8_t* pp = pos.data + pos.stride * v;
                glm::vec3 n(0.0f);
                if (haveNrm) { const uint8_t* np = nrm.data + nrm.stride * v; n = glm::vec3(read_f32(np), read_f32(np + 4), read_f32(np + 8)); }
                glm::vec3 normal = glm_normalize(n);
                Vertex vert{};
                vert.position = glm::vec3(read_f32(pp), read_f32(pp + 4), read_f32(pp + 8));
                vert.normal = invertNormals ? glm::vec3(normal.x * -1.0f, normal.y * -1.0f, normal.z * -1.0f) : normal;
                if (haveUv) {
                    const uint8_t* up = uv.data + uv.stride * v;
                    if (uv.componentType == 5126) vert.uv = glm::vec2(read_f32(up), read_f32(up + 4));
                    else if (uv.componentType == 5121) vert.uv = glm::vec2(up[0] / 255.0f, up[1] / 255.0f);
                    else { uint16_t a, b; std::memcpy(&a, up, 2); std::memcpy(&b, up + 2, 2); vert.uv = glm::vec2(a / 65535.0f, b / 65535.0f); }
                } else vert.uv = glm::vec2(0.0f, 0.0f);
                vert.color = glm::vec3(1.0f);
                mesh->_vertices.push_back(vert);
            }

            const int indicesAcc = tprim.get_int("indices", -1);
            if (indicesAcc > -1) {
                View idx = accessor_view(M, indicesAcc);
                if (idx.ncomp != 1) gfail(M, "indices must be SCALAR");
                indexCount += (uint32_t)idx.count;
                mesh->_indices.reserve(mesh->_indices.size() + idx.count);
                for (size_t k = 0; k < idx.count; k++) {
                    const uint8_t* ip = idx.data + idx.stride * k;
                    uint32_t value;
                    switch (idx.componentType) {
                    case 5125: std::memcpy(&value, ip, 4); break;
                    case 5123: { uint16_t s; std::memcpy(&s, ip, 2); value = s; break; }
                    case 5121: value = ip[0]; break;
                    default: gfail(M, "Index component type " + std::to_string(idx.componentType) + " not supported!");
                    }
                    if (value >= vertexCount) gfail(M, "index out of range");
                    mesh->_indices.push_back(value + firstVertex);
                }
            } else {
                indexCount = vertexCount;
                for (uint32_t k = 0; k < vertexCount; k++) mesh->_indices.push_back(k + firstVertex);
            }
            const int mode = tprim.get_int("mode", 4);
            if (mode != 4) gfail(M, "only TRIANGLES primitives are supported (mode " + std::to_string(mode) + ")");
            if (indexCount % 3) gfail(M, "index count is not a multiple of 3");

            Primitive* prim = new Primitive();
            prim->firstIndex = firstIndex;
            prim->indexCount = indexCount;
            prim->firstVertex = firstVertex;
            prim->vertexCount = vertexCount;
            prim->materialID = load_material(M, tprim.get_int("material", -1));
            load_textures(M, prim->materialID);
            node->_primitives.push_back(prim);
        }
    }

    if (parent) { node->_parent = parent; parent->_children.push_back(node); }
    else prefab->_root.push_back(node);
}

void load_document(Model& M, const std::string& path, bool binary)
{
    std::vector<uint8_t> file;
    if (!read_file(path, file)) gfail(M, "cannot read file");
    std::vector<uint8_t> glbBin;
    bool haveGlbBin = false;
    if (binary) {
        auto u32 = [&](size_t o) -> uint32_t { uint32_t v; std::memcpy(&v, file.data() + o, 4); return v; };
        if (file.size() < 20 || u32(0) != 0x46546C67u) gfail(M, "not a GLB container");
        if (u32(4) != 2) gfail(M, "unsupported GLB version");
        size_t total = u32(8);
        if (total > file.size()) gfail(M, "GLB length exceeds the file");
        size_t p = 12;
        std::string json;
        while (p + 8 <= total) {
            uint32_t len = u32(p), type = u32(p + 4);
            if (p + 8 + (size_t)len > total) gfail(M, "GLB chunk exceeds the file");
            if (type == 0x4E4F534Au && json.empty()) json.assign(reinterpret_cast<const char*>(file.data() + p + 8), len);
            else if (type == 0x004E4942u && !haveGlbBin) { glbBin.assign(file.begin() + (long)(p + 8), file.begin() + (long)(p + 8 + len)); haveGlbBin = true; }
            p += 8 + (size_t)len;
            p = (p + 3) & ~(size_t)3;
        }
        if (json.empty()) gfail(M, "GLB without a JSON chunk");
        M.doc = jsonmin::parse(json);
    } else {
        M.doc = jsonmin::parse(std::string(reinterpret_cast<const char*>(file.data()), file.size()));
    }
    if (!M.doc.is_object()) gfail(M, "document is not a JSON object");
    if (const Value* asset = M.doc.find("asset")) {
        std::string version = asset->get_str("version");
        if (!version.empty() && version[0] != '2') gfail(M, "only glTF 2.x is supported (asset.version " + version + ")");
    }
    const size_t nb = count_of(M, "buffers");
    M.buffers.resize(nb);
    for (size_t i = 0; i < nb; i++) {
        const Value& b = element(M, "buffers", (int)i);
        const std::string uri = b.get_str("uri");
        if (uri.empty()) {
            if (i != 0 || !haveGlbBin) gfail(M, "buffer " + std::to_string(i) + " has no uri");
            M.buffers[i] = std::move(glbBin);
        } else if (is_data_uri(uri)) M.buffers[i] = decode_data_uri(M, uri);
        else {
            std::string rel = percent_decode(uri);
            if (!read_file(M.dir + rel, M.buffers[i])) {
                std::string alt = vkutil_find_file(rel);
                if (alt.empty() || !read_file(alt, M.buffers[i])) gfail(M, "cannot read buffer " + rel);
            }
        }
        const size_t want = (size_t)b.get_num("byteLength", 0);
        if (M.buffers[i].size() < want) gfail(M, "buffer " + std::to_string(i) + " is shorter than its byteLength");
    }
}

} // namespace

// The body of the glTF branch of Prefab::GET (vk_mesh.cpp:857-896).  Throws std::runtime_error like the reference
// does on a tinygltf error (:880-881).
Prefab* load_gltf_prefab(const std::string& path, bool binary, bool invertNormals)
{
    Model M;
    M.name = path;
    size_t slash = path.find_last_of("/\\");
    M.dir = slash == std::string::npos ? "" : path.substr(0, slash + 1);
    load_document(M, path, binary);

    Prefab* prefab = new Prefab();
    prefab->_mesh = new Mesh();
    const Value* scenes = M.doc.find("scenes");
    if (!scenes || !scenes->is_array() || scenes->arr.empty()) gfail(M, "no scenes");
    const Value* nodes = scenes->arr[0].find("nodes");   // gltfModel.scenes[0] (:885), not the `scene` property
    if (nodes && nodes->is_array())
        for (const Value& n : nodes->arr) load_node(M, prefab, element(M, "nodes", n.is_number() ? (int)n.number : -1), nullptr, invertNormals, 0);
    return prefab;
}
