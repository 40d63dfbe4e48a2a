// procedural.h -- deterministic stand-ins for the reference's missing assets (see procedural.cpp)
#pragma once
#include <cstdint>
#include <vector>
#include "vkpinut.h"

namespace proc {
void uv_sphere(int segments, int rings, std::vector<Vertex>& v, std::vector<uint32_t>& idx);
void lucy_standin(std::vector<Vertex>& v, std::vector<uint32_t>& idx, int subdiv = 7);
void helmet_standin(std::vector<Vertex>& v, std::vector<uint32_t>& idx, int subdiv = 5);
void displaced_icosphere(std::vector<Vertex>& v, std::vector<uint32_t>& idx, int subdiv, uint32_t seed, float amount);
void sky_texture(int w, int h, std::vector<uint8_t>& rgba);
void irradiance_texture(int w, int h, std::vector<uint8_t>& rgba);
void helmet_textures(int w, int h, std::vector<uint8_t>& albedo, std::vector<uint8_t>& emissive, std::vector<uint8_t>& metal_rough, std::vector<uint8_t>& normal);
}
