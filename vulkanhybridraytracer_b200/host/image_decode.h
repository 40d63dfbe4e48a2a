// image_decode.h -- 8-bit RGBA image loading for Texture::GET (src/vk_textures.cpp:12-22 calls
// stbi_load(filename, &w, &h, &n, STBI_rgb_alpha); stb_image is un-vendored upstream, .gitignore:7).
// PNG (all colour types / bit depths / Adam7), JPEG (baseline + progressive Huffman, 8-bit, grey or YCbCr / RGB)
// and Radiance .hdr (tone-mapped to 8 bits the way stbi_load does for an .hdr file).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace imgdec {

// Decodes an image held in memory.  On failure returns false and describes the reason in *err (when given).
bool decode_rgba8(const uint8_t* bytes, size_t n, int& w, int& h, std::vector<uint8_t>& rgba, std::string* err = nullptr);
bool load_rgba8(const char* path, int& w, int& h, std::vector<uint8_t>& rgba, std::string* err = nullptr);

} // namespace imgdec
