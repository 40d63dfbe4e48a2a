// output.cpp -- headless replacement of the swapchain present (window.cpp / renderer.cpp:590-599):
// PNG for the gamma-corrected 8-bit image, uncompressed scan-line OpenEXR for linear radiance.
#include "vkpinut.h"

#include <cstdio>
#include <cstring>
#include <zlib.h>

static void put_be32(std::vector<uint8_t>& b, uint32_t v) { b.push_back(v >> 24); b.push_back(v >> 16); b.push_back(v >> 8); b.push_back(v); }

static void png_chunk(std::vector<uint8_t>& out, const char tag[4], const std::vector<uint8_t>& data)
{
    put_be32(out, (uint32_t)data.size());
    size_t start = out.size();
    out.insert(out.end(), tag, tag + 4);
    out.insert(out.end(), data.begin(), data.end());
    uint32_t crc = crc32(0L, out.data() + start, (uInt)(out.size() - start));
    put_be32(out, crc);
}

bool write_png_rgba8(const char* path, const uint8_t* rgba, int w, int h)
{
    std::vector<uint8_t> raw((size_t)h * ((size_t)w * 4 + 1));
    for (int y = 0; y < h; y++) {
        raw[(size_t)y * (w * 4 + 1)] = 0;   // filter: none
        std::memcpy(&raw[(size_t)y * (w * 4 + 1) + 1], rgba + (size_t)y * w * 4, (size_t)w * 4);
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
    comp.resize(clen);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, (uint32_t)w); put_be32(ihdr, (uint32_t)h);
    ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    png_chunk(out, "IHDR", ihdr);
    png_chunk(out, "IDAT", comp);
    png_chunk(out, "IEND", {});
    FILE* f = std::fopen(path, "wb");
    if (!f) return false;
    bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    std::fclose(f);
    return ok;
}

namespace {
struct Exr {
    std::vector<uint8_t> b;
    void bytes(const void* p, size_t n) { const uint8_t* c = (const uint8_t*)p; b.insert(b.end(), c, c + n); }
    void str(const char* s) { bytes(s, std::strlen(s) + 1); }
    void i32(int32_t v) { bytes(&v, 4); }
    void u64(uint64_t v) { bytes(&v, 8); }
    void f32(float v) { bytes(&v, 4); }
    void attr(const char* name, const char* type, const std::vector<uint8_t>& v) { str(name); str(type); i32((int32_t)v.size()); bytes(v.data(), v.size()); }
};
}

// OpenEXR 2 single-part scan-line file, NO_COMPRESSION, three FLOAT channels B, G, R (alphabetical) + A
bool write_exr_rgba32f(const char* path, const float* rgba, int w, int h)
{
    Exr e;
    e.i32(20000630); e.i32(2);
    {   // channels
        Exr c;
        const char* names[4] = {"A", "B", "G", "R"};
        for (const char* n : names) { c.str(n); c.i32(2); c.b.push_back(0); c.b.push_back(0); c.b.push_back(0); c.b.push_back(0); c.i32(1); c.i32(1); }
        c.b.push_back(0);
        e.attr("channels", "chlist", c.b);
    }
    { Exr c; c.b.push_back(0); e.attr("compression", "compression", c.b); }
    { Exr c; c.i32(0); c.i32(0); c.i32(w - 1); c.i32(h - 1); e.attr("dataWindow", "box2i", c.b); e.attr("displayWindow", "box2i", c.b); }
    { Exr c; c.b.push_back(0); e.attr("lineOrder", "lineOrder", c.b); }
    { Exr c; c.f32(1.0f); e.attr("pixelAspectRatio", "float", c.b); }
    { Exr c; c.f32(0.0f); c.f32(0.0f); e.attr("screenWindowCenter", "v2f", c.b); }
    { Exr c; c.f32(1.0f); e.attr("screenWindowWidth", "float", c.b); }
    e.b.push_back(0);
    const size_t table = e.b.size();
    const size_t line_bytes = 8 + (size_t)w * 4 * 4;
    for (int y = 0; y < h; y++) e.u64((uint64_t)(table + (size_t)h * 8 + (size_t)y * line_bytes));
    std::vector<float> plane((size_t)w);
    const int order[4] = {3, 2, 1, 0};   // A, B, G, R from RGBA
    for (int y = 0; y < h; y++) {
        e.i32(y); e.i32((int32_t)((size_t)w * 4 * 4));
        for (int c = 0; c < 4; c++) {
            for (int x = 0; x < w; x++) plane[x] = rgba[((size_t)y * w + x) * 4 + order[c]];
            e.bytes(plane.data(), (size_t)w * 4);
        }
    }
    FILE* f = std::fopen(path, "wb");
    if (!f) return false;
    bool ok = std::fwrite(e.b.data(), 1, e.b.size(), f) == e.b.size();
    std::fclose(f);
    return ok;
}
