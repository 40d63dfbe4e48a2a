// image_decode.cpp -- PNG / JPEG / Radiance-HDR decoders producing the 8-bit RGBA pixels the reference gets from
// stbi_load(..., STBI_rgb_alpha) (src/vk_textures.cpp:12-22).  stb_image (public domain, un-vendored upstream,
// version not pinned) is absent from this image, so its published algorithms are restated here:
//   PNG   lossless: RFC 2083 filters + zlib inflate; 16-bit samples keep their high byte, low-depth greys are
//         scaled by 255/(2^d-1), tRNS gives the alpha, no gamma handling.
//   JPEG  the entropy-decoded coefficients are fixed by ITU T.81; the lossy back end follows stb_image's integer
//         pipeline (12-bit fixed-point IDCT with 2 guard bits between the passes, the 3:1 "triangle" chroma
//         up-sampling of its resample_row_* functions, BT.601 full-range YCbCr in 20-bit fixed point) so that texels
//         match what the reference samples rather than what libjpeg would produce.
//   HDR   RGBE scanlines (flat or new-style RLE), then stb's hdr-to-ldr: pow(v, 1/2.2) * 255 + 0.5, clamped.
// tests/test_host_assets.py checks PNG bit-exactly and JPEG within the inter-decoder tolerance against Pillow.
#include "image_decode.h"

#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace imgdec {
namespace {

struct DecodeError : std::runtime_error { using std::runtime_error::runtime_error; };
[[noreturn]] void fail(const char* what) { throw DecodeError(what); }

// ================================================================================================ PNG
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

int paeth(int a, int b, int c)
{
    int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    if (pa <= pb && pa <= pc) return a;
    return pb <= pc ? b : c;
}

// One (sub)image of `w` x `h` pixels: undo the per-scanline filters in place, return rows of packed samples.
void png_unfilter(uint8_t* raw, size_t raw_len, int w, int h, int channels, int depth, std::vector<uint8_t>& rows)
{
    const size_t stride = ((size_t)w * channels * depth + 7) / 8;
    const int bpp = channels * depth >= 8 ? channels * depth / 8 : 1;
    if (raw_len < (stride + 1) * (size_t)h) fail("png: not enough pixel data");
    rows.assign(stride * (size_t)h, 0);
    std::vector<uint8_t> zero(stride, 0);
    for (int y = 0; y < h; y++) {
        const uint8_t* src = raw + (stride + 1) * (size_t)y;
        uint8_t* cur = rows.data() + stride * (size_t)y;
        const uint8_t* up = y ? cur - stride : zero.data();
        const int filter = src[0];
        ++src;
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = up[i], c = i >= (size_t)bpp ? up[i - bpp] : 0;
            int v = src[i];
            switch (filter) {
            case 0: break;
            case 1: v += a; break;
            case 2: v += b; break;
            case 3: v += (a + b) >> 1; break;
            case 4: v += paeth(a, b, c); break;
            default: fail("png: invalid filter");
            }
            cur[i] = (uint8_t)v;
        }
    }
}

struct PngInfo {
    int w{0}, h{0}, depth{0}, color{0}, interlace{0};
    std::vector<uint8_t> palette;   // rgb triples
    std::vector<uint8_t> trns;      // palette alpha, or the 16-bit colour key (2 or 6 bytes)
    bool has_trns{false};
};

// Writes the pixel whose samples are s[0..channels) (raw sample values of `depth` bits) as RGBA8.
inline void png_store(const PngInfo& I, const unsigned* s, uint8_t* out)
{
    static const int scale[9] = {0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01};
    auto to8 = [&](unsigned v) -> uint8_t { return I.depth == 16 ? (uint8_t)(v >> 8) : (uint8_t)(v * (unsigned)scale[I.depth]); };
    switch (I.color) {
    case 0: {   // grey
        uint8_t g = to8(s[0]);
        out[0] = out[1] = out[2] = g; out[3] = 255;
        if (I.has_trns && I.trns.size() >= 2) {
            unsigned key = ((unsigned)I.trns[0] << 8) | I.trns[1];
            if (I.depth != 16) key &= 255u;
            if (s[0] == key) out[3] = 0;
        }
        break;
    }
    case 2: {   // rgb
        out[0] = to8(s[0]); out[1] = to8(s[1]); out[2] = to8(s[2]); out[3] = 255;
        if (I.has_trns && I.trns.size() >= 6) {
            bool match = true;
            for (int k = 0; k < 3; k++) {
                unsigned key = ((unsigned)I.trns[2 * k] << 8) | I.trns[2 * k + 1];
                if (I.depth != 16) key &= 255u;
                match = match && s[k] == key;
            }
            if (match) out[3] = 0;
        }
        break;
    }
    case 3: {   // palette
        unsigned idx = s[0];
        if (idx * 3 + 2 < I.palette.size()) { out[0] = I.palette[idx * 3]; out[1] = I.palette[idx * 3 + 1]; out[2] = I.palette[idx * 3 + 2]; }
        else { out[0] = out[1] = out[2] = 0; }
        out[3] = I.has_trns && idx < I.trns.size() ? I.trns[idx] : 255;
        break;
    }
    case 4: { uint8_t g = to8(s[0]); out[0] = out[1] = out[2] = g; out[3] = to8(s[1]); break; }
    case 6: out[0] = to8(s[0]); out[1] = to8(s[1]); out[2] = to8(s[2]); out[3] = to8(s[3]); break;
    default: fail("png: bad colour type");
    }
}

void png_expand_pass(const PngInfo& I, const std::vector<uint8_t>& rows, int pw, int ph, int channels,
                     int x0, int y0, int dx, int dy, std::vector<uint8_t>& rgba)
{
    const size_t stride = ((size_t)pw * channels * I.depth + 7) / 8;
    for (int y = 0; y < ph; y++) {
        const uint8_t* r = rows.data() + stride * (size_t)y;
        for (int x = 0; x < pw; x++) {
            unsigned s[4] = {0, 0, 0, 0};
            for (int c = 0; c < channels; c++) {
                const size_t k = (size_t)x * channels + c;
                if (I.depth == 8) s[c] = r[k];
                else if (I.depth == 16) s[c] = ((unsigned)r[2 * k] << 8) | r[2 * k + 1];
                else {
                    const size_t bit = k * I.depth;
                    s[c] = (r[bit >> 3] >> (8 - I.depth - (bit & 7))) & ((1u << I.depth) - 1);
                }
            }
            png_store(I, s, rgba.data() + 4 * ((size_t)(y0 + y * dy) * I.w + (x0 + x * dx)));
        }
    }
}

bool is_png(const uint8_t* b, size_t n) { static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10}; return n >= 8 && std::memcmp(b, sig, 8) == 0; }

void decode_png(const uint8_t* b, size_t n, int& w, int& h, std::vector<uint8_t>& rgba)
{
    PngInfo I;
    std::vector<uint8_t> idat;
    size_t p = 8;
    bool seen_ihdr = false, done = false;
    while (!done) {
        if (p + 8 > n) fail("png: truncated chunk header");
        const uint32_t len = be32(b + p);
        const uint8_t* type = b + p + 4;
        const uint8_t* d = b + p + 8;
        if (p + 8 + (size_t)len > n) fail("png: truncated chunk");   // the trailing CRC is not checked (stb does not either)
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13) fail("png: bad IHDR");
            I.w = (int)be32(d); I.h = (int)be32(d + 4); I.depth = d[8]; I.color = d[9]; I.interlace = d[12];
            if (I.w <= 0 || I.h <= 0 || I.w > (1 << 24) || I.h > (1 << 24)) fail("png: bad size");
            if (I.depth != 1 && I.depth != 2 && I.depth != 4 && I.depth != 8 && I.depth != 16) fail("png: bad bit depth");
            if (I.color != 0 && I.color != 2 && I.color != 3 && I.color != 4 && I.color != 6) fail("png: bad colour type");
            if (I.color == 3 && I.depth == 16) fail("png: bad palette depth");
            if ((I.color == 2 || I.color == 4 || I.color == 6) && I.depth < 8) fail("png: bad bit depth for colour type");
            if (d[10] != 0 || d[11] != 0 || I.interlace > 1) fail("png: bad compression / filter / interlace method");
            seen_ihdr = true;
        } else if (!seen_ihdr) fail("png: first chunk is not IHDR");
        else if (!std::memcmp(type, "PLTE", 4)) { if (len > 768 || len % 3) fail("png: bad PLTE"); I.palette.assign(d, d + len); }
        else if (!std::memcmp(type, "tRNS", 4)) { I.trns.assign(d, d + len); I.has_trns = true; }
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), d, d + len);
        else if (!std::memcmp(type, "IEND", 4)) done = true;
        p += 12 + (size_t)len;
    }
    if (!seen_ihdr || idat.empty()) fail("png: no image data");
    if (I.color == 3 && I.palette.empty()) fail("png: missing palette");
    const int channels = I.color == 0 ? 1 : I.color == 2 ? 3 : I.color == 3 ? 1 : I.color == 4 ? 2 : 4;

    // size of the filtered stream
    auto pass_bytes = [&](int pw, int ph) -> size_t { return pw > 0 && ph > 0 ? ((((size_t)pw * channels * I.depth + 7) / 8) + 1) * (size_t)ph : 0; };
    static const int xo[7] = {0, 4, 0, 2, 0, 1, 0}, yo[7] = {0, 0, 4, 0, 2, 0, 1}, xs[7] = {8, 8, 4, 4, 2, 2, 1}, ys[7] = {8, 8, 8, 4, 4, 2, 2};
    size_t total = 0;
    if (!I.interlace) total = pass_bytes(I.w, I.h);
    else for (int k = 0; k < 7; k++) total += pass_bytes((I.w - xo[k] + xs[k] - 1) / xs[k], (I.h - yo[k] + ys[k] - 1) / ys[k]);

    std::vector<uint8_t> raw(total);
    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    if (inflateInit(&zs) != Z_OK) fail("png: zlib init failed");
    zs.next_in = idat.data(); zs.avail_in = (uInt)idat.size();
    zs.next_out = raw.data(); zs.avail_out = (uInt)raw.size();
    int zr = inflate(&zs, Z_FINISH);
    size_t got = zs.total_out;
    inflateEnd(&zs);
    if (zr != Z_STREAM_END && !(zr == Z_BUF_ERROR && got == total) && !(zr == Z_OK && got == total)) fail("png: corrupt zlib stream");
    if (got < total) fail("png: not enough pixel data");

    w = I.w; h = I.h;
    rgba.assign((size_t)w * h * 4, 0);
    std::vector<uint8_t> rows;
    if (!I.interlace) {
        png_unfilter(raw.data(), raw.size(), w, h, channels, I.depth, rows);
        png_expand_pass(I, rows, w, h, channels, 0, 0, 1, 1, rgba);
    } else {
        size_t off = 0;
        for (int k = 0; k < 7; k++) {
            const int pw = (I.w - xo[k] + xs[k] - 1) / xs[k], ph = (I.h - yo[k] + ys[k] - 1) / ys[k];
            if (pw <= 0 || ph <= 0) continue;
            png_unfilter(raw.data() + off, raw.size() - off, pw, ph, channels, I.depth, rows);
            png_expand_pass(I, rows, pw, ph, channels, xo[k], yo[k], xs[k], ys[k], rgba);
            off += pass_bytes(pw, ph);
        }
    }
}

// ================================================================================================ JPEG
const uint8_t kDezigzag[64 + 15] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
    // a corrupt run may step past 63: land on the last coefficient
    63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

struct Huffman {
    uint8_t values[256];
    uint8_t size[257];
    unsigned maxcode[18];
    int delta[17];
    bool present{false};

    void build(const int* count)
    {
        int k = 0;
        for (int i = 0; i < 16; i++) for (int j = 0; j < count[i]; j++) { if (k >= 256) fail("jpeg: bad huffman table"); size[k++] = (uint8_t)(i + 1); }
        size[k] = 0;
        unsigned code = 0;
        k = 0;
        for (int j = 1; j <= 16; j++) {
            delta[j] = k - (int)code;
            if (size[k] == j) {
                while (size[k] == j) { ++k; ++code; }
                if (code - 1 >= (1u << j)) fail("jpeg: bad code lengths");
            }
            maxcode[j] = code << (16 - j);
            code <<= 1;
        }
        maxcode[17] = 0xffffffffu;
        present = true;
    }
};

inline uint8_t clamp8(int x) { return (unsigned)x > 255 ? (x < 0 ? 0 : 255) : (uint8_t)x; }

// stb_image's integer IDCT (a fixed-point AAN-free LLM butterfly): constants scaled by 4096, two extra bits kept
// between the column and the row pass, +128 level shift and rounding folded into the row pass.
#define IMGDEC_F2F(x) ((int)(((x) * 4096 + 0.5)))
#define IMGDEC_FSH(x) ((x) * 4096)
#define IMGDEC_IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)          \
    int t0, t1, t2, t3, p1, p2, p3, p4, p5, x0, x1, x2, x3;      \
    p2 = s2; p3 = s6;                                             \
    p1 = (p2 + p3) * IMGDEC_F2F(0.5411961f);                      \
    t2 = p1 + p3 * IMGDEC_F2F(-1.847759065f);                     \
    t3 = p1 + p2 * IMGDEC_F2F(0.765366865f);                      \
    p2 = s0; p3 = s4;                                             \
    t0 = IMGDEC_FSH(p2 + p3);                                     \
    t1 = IMGDEC_FSH(p2 - p3);                                     \
    x0 = t0 + t3; x3 = t0 - t3; x1 = t1 + t2; x2 = t1 - t2;       \
    t0 = s7; t1 = s5; t2 = s3; t3 = s1;                           \
    p3 = t0 + t2; p4 = t1 + t3; p1 = t0 + t3; p2 = t1 + t2;       \
    p5 = (p3 + p4) * IMGDEC_F2F(1.175875602f);                    \
    t0 = t0 * IMGDEC_F2F(0.298631336f);                           \
    t1 = t1 * IMGDEC_F2F(2.053119869f);                           \
    t2 = t2 * IMGDEC_F2F(3.072711026f);                           \
    t3 = t3 * IMGDEC_F2F(1.501321110f);                           \
    p1 = p5 + p1 * IMGDEC_F2F(-0.899976223f);                     \
    p2 = p5 + p2 * IMGDEC_F2F(-2.562915447f);                     \
    p3 = p3 * IMGDEC_F2F(-1.961570560f);                          \
    p4 = p4 * IMGDEC_F2F(-0.390180644f);                          \
    t3 += p1 + p4; t2 += p2 + p3; t1 += p2 + p4; t0 += p1 + p3;

void idct_block(uint8_t* out, int out_stride, const short data[64])
{
    int val[64];
    int* v = val;
    const short* d = data;
    for (int i = 0; i < 8; ++i, ++d, ++v) {
        if (d[8] == 0 && d[16] == 0 && d[24] == 0 && d[32] == 0 && d[40] == 0 && d[48] == 0 && d[56] == 0) {
            int dcterm = d[0] * 4;
            v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = dcterm;
        } else {
            IMGDEC_IDCT_1D(d[0], d[8], d[16], d[24], d[32], d[40], d[48], d[56])
            x0 += 512; x1 += 512; x2 += 512; x3 += 512;
            v[0] = (x0 + t3) >> 10; v[56] = (x0 - t3) >> 10;
            v[8] = (x1 + t2) >> 10; v[48] = (x1 - t2) >> 10;
            v[16] = (x2 + t1) >> 10; v[40] = (x2 - t1) >> 10;
            v[24] = (x3 + t0) >> 10; v[32] = (x3 - t0) >> 10;
        }
    }
    v = val;
    uint8_t* o = out;
    for (int i = 0; i < 8; ++i, v += 8, o += out_stride) {
        IMGDEC_IDCT_1D(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7])
        x0 += 65536 + (128 << 17); x1 += 65536 + (128 << 17); x2 += 65536 + (128 << 17); x3 += 65536 + (128 << 17);
        o[0] = clamp8((x0 + t3) >> 17); o[7] = clamp8((x0 - t3) >> 17);
        o[1] = clamp8((x1 + t2) >> 17); o[6] = clamp8((x1 - t2) >> 17);
        o[2] = clamp8((x2 + t1) >> 17); o[5] = clamp8((x2 - t1) >> 17);
        o[3] = clamp8((x3 + t0) >> 17); o[4] = clamp8((x3 - t0) >> 17);
    }
}

struct Jpeg {
    const uint8_t* p;
    const uint8_t* end;
    int W{0}, H{0}, ncomp{0};
    struct Comp {
        int id{0}, h{1}, v{1}, tq{0}, hd{0}, ha{0}, dc_pred{0};
        int x{0}, y{0}, w2{0}, h2{0}, coeff_w{0}, coeff_h{0};
        std::vector<uint8_t> data;
        std::vector<short> coeff;
    } comp[4];
    int hmax{1}, vmax{1}, mcu_w{8}, mcu_h{8}, mcux{0}, mcuy{0};
    uint16_t dequant[4][64];
    Huffman hdc[4], hac[4];
    bool progressive{false};
    int spec_start{0}, spec_end{63}, succ_high{0}, succ_low{0}, eob_run{0};
    int restart_interval{0}, todo{0};
    uint32_t code_buffer{0};
    int code_bits{0};
    int marker{0xff};   // 0xff = none pending
    bool nomore{false};
    int scan_n{0}, order[4]{0, 0, 0, 0};
    int app14_transform{-1};
    bool jfif{false};
    int rgb{0};

    int get8() { return p < end ? *p++ : 0; }
    int get16() { int a = get8(); return (a << 8) | get8(); }
    bool eof() const { return p >= end; }

    void grow()
    {
        do {
            unsigned b = nomore ? 0u : (unsigned)get8();
            if (b == 0xff) {
                int c = get8();
                while (c == 0xff) c = get8();   // fill bytes
                if (c != 0) { marker = c; nomore = true; return; }
            }
            code_buffer |= b << (24 - code_bits);
            code_bits += 8;
        } while (code_bits <= 24);
    }
    int decode(const Huffman& hf)
    {
        if (code_bits < 16) grow();
        unsigned temp = code_buffer >> 16;
        int k = 1;
        while (temp >= hf.maxcode[k]) ++k;
        if (k == 17) { code_bits -= 16; fail("jpeg: bad huffman code"); }
        if (k > code_bits) fail("jpeg: bad huffman code");
        int c = (int)((code_buffer >> (32 - k)) & ((1u << k) - 1)) + hf.delta[k];
        if (c < 0 || c >= 256) fail("jpeg: bad huffman code");
        code_bits -= k;
        code_buffer <<= k;
        return hf.values[c];
    }
    // n-bit two's-complement-ish value of T.81 F.2.2.1 (EXTEND)
    int extend_receive(int n)
    {
        if (n == 0) return 0;
        if (code_bits < n) grow();
        if (code_bits < n) return 0;
        unsigned sgn = code_buffer >> 31;
        unsigned k = code_buffer >> (32 - n);
        code_buffer <<= n;
        code_bits -= n;
        return (int)k + (sgn ? 0 : (int)((~0u << n) + 1u));
    }
    int get_bits(int n)
    {
        if (n == 0) return 0;
        if (code_bits < n) grow();
        if (code_bits < n) return 0;
        unsigned k = code_buffer >> (32 - n);
        code_buffer <<= n;
        code_bits -= n;
        return (int)k;
    }
    int get_bit() { return get_bits(1); }

    void reset()
    {
        code_bits = 0; code_buffer = 0; nomore = false;
        for (int i = 0; i < 4; i++) comp[i].dc_pred = 0;
        marker = 0xff;
        todo = restart_interval ? restart_interval : 0x7fffffff;
        eob_run = 0;
    }

    void decode_block(short data[64], const Huffman& dc, const Huffman& ac, int b, const uint16_t* dq)
    {
        if (code_bits < 16) grow();
        int t = decode(dc);
        if (t > 15) fail("jpeg: bad DC category");
        std::memset(data, 0, 64 * sizeof(short));
        int diff = t ? extend_receive(t) : 0;
        int dcv = comp[b].dc_pred + diff;
        comp[b].dc_pred = dcv;
        data[0] = (short)(dcv * dq[0]);
        int k = 1;
        do {
            int rs = decode(ac);
            int s = rs & 15, r = rs >> 4;
            if (s == 0) {
                if (rs != 0xf0) break;
                k += 16;
            } else {
                k += r;
                unsigned zig = kDezigzag[k++];
                data[zig] = (short)(extend_receive(s) * dq[zig]);
            }
        } while (k < 64);
    }
    void decode_block_prog_dc(short data[64], const Huffman& dc, int b)
    {
        if (spec_end != 0) fail("jpeg: can't merge DC and AC");
        if (code_bits < 16) grow();
        if (succ_high == 0) {
            std::memset(data, 0, 64 * sizeof(short));
            int t = decode(dc);
            if (t > 15) fail("jpeg: bad DC category");
            int diff = t ? extend_receive(t) : 0;
            int dcv = comp[b].dc_pred + diff;
            comp[b].dc_pred = dcv;
            data[0] = (short)(dcv * (1 << succ_low));
        } else if (get_bit()) data[0] += (short)(1 << succ_low);
    }
    void decode_block_prog_ac(short data[64], const Huffman& ac)
    {
        if (spec_start == 0) fail("jpeg: can't merge DC and AC");
        if (succ_high == 0) {
            const int shift = succ_low;
            if (eob_run) { --eob_run; return; }
            int k = spec_start;
            do {
                int rs = decode(ac);
                int s = rs & 15, r = rs >> 4;
                if (s == 0) {
                    if (r < 15) {
                        eob_run = 1 << r;
                        if (r) eob_run += get_bits(r);
                        --eob_run;
                        break;
                    }
                    k += 16;
                } else {
                    k += r;
                    unsigned zig = kDezigzag[k++];
                    data[zig] = (short)(extend_receive(s) * (1 << shift));
                }
            } while (k <= spec_end);
        } else {
            const short bit = (short)(1 << succ_low);
            auto refine = [&](short* q) {
                if (get_bit() && (*q & bit) == 0) { if (*q > 0) *q += bit; else *q -= bit; }
            };
            if (eob_run) {
                --eob_run;
                for (int k = spec_start; k <= spec_end; ++k) { short* q = &data[kDezigzag[k]]; if (*q != 0) refine(q); }
            } else {
                int k = spec_start;
                do {
                    int rs = decode(ac);
                    int s = rs & 15, r = rs >> 4;
                    if (s == 0) {
                        if (r < 15) {
                            eob_run = (1 << r) - 1;
                            if (r) eob_run += get_bits(r);
                            r = 64;   // force end of block
                        }
                    } else {
                        if (s != 1) fail("jpeg: bad refinement code");
                        s = get_bit() ? bit : -bit;
                    }
                    while (k <= spec_end) {
                        short* q = &data[kDezigzag[k++]];
                        if (*q != 0) refine(q);
                        else {
                            if (r == 0) { *q = (short)s; break; }
                            --r;
                        }
                    }
                } while (k <= spec_end);
            }
        }
    }

    static bool is_restart(int m) { return m >= 0xd0 && m <= 0xd7; }
    bool after_unit()   // false = stop the scan (a non-restart marker is pending)
    {
        if (--todo <= 0) {
            if (code_bits < 24) grow();
            if (!is_restart(marker)) return false;
            reset();
        }
        return true;
    }

    void entropy_coded_data()
    {
        reset();
        short block[64];
        if (!progressive) {
            if (scan_n == 1) {
                const int n = order[0];
                Comp& c = comp[n];
                const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
                for (int j = 0; j < bh; ++j) for (int i = 0; i < bw; ++i) {
                    decode_block(block, hdc[c.hd], hac[c.ha], n, dequant[c.tq]);
                    idct_block(c.data.data() + (size_t)c.w2 * j * 8 + i * 8, c.w2, block);
                    if (!after_unit()) return;
                }
            } else {
                for (int j = 0; j < mcuy; ++j) for (int i = 0; i < mcux; ++i) {
                    for (int k = 0; k < scan_n; ++k) {
                        const int n = order[k];
                        Comp& c = comp[n];
                        for (int y = 0; y < c.v; ++y) for (int x = 0; x < c.h; ++x) {
                            const int x2 = (i * c.h + x) * 8, y2 = (j * c.v + y) * 8;
                            decode_block(block, hdc[c.hd], hac[c.ha], n, dequant[c.tq]);
                            idct_block(c.data.data() + (size_t)c.w2 * y2 + x2, c.w2, block);
                        }
                    }
                    if (!after_unit()) return;
                }
            }
        } else {
            if (scan_n == 1) {
                const int n = order[0];
                Comp& c = comp[n];
                const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
                for (int j = 0; j < bh; ++j) for (int i = 0; i < bw; ++i) {
                    short* data = c.coeff.data() + 64 * ((size_t)i + (size_t)j * c.coeff_w);
                    if (spec_start == 0) decode_block_prog_dc(data, hdc[c.hd], n);
                    else decode_block_prog_ac(data, hac[c.ha]);
                    if (!after_unit()) return;
                }
            } else {
                for (int j = 0; j < mcuy; ++j) for (int i = 0; i < mcux; ++i) {
                    for (int k = 0; k < scan_n; ++k) {
                        const int n = order[k];
                        Comp& c = comp[n];
                        for (int y = 0; y < c.v; ++y) for (int x = 0; x < c.h; ++x) {
                            const int x2 = i * c.h + x, y2 = j * c.v + y;
                            short* data = c.coeff.data() + 64 * ((size_t)x2 + (size_t)y2 * c.coeff_w);
                            decode_block_prog_dc(data, hdc[c.hd], n);
                        }
                    }
                    if (!after_unit()) return;
                }
            }
        }
    }

    void finish_progressive()
    {
        for (int n = 0; n < ncomp; ++n) {
            Comp& c = comp[n];
            const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
            for (int j = 0; j < bh; ++j) for (int i = 0; i < bw; ++i) {
                short* data = c.coeff.data() + 64 * ((size_t)i + (size_t)j * c.coeff_w);
                const uint16_t* dq = dequant[c.tq];
                for (int k = 0; k < 64; ++k) data[k] = (short)(data[k] * dq[k]);
                idct_block(c.data.data() + (size_t)c.w2 * j * 8 + i * 8, c.w2, data);
            }
        }
    }

    void process_marker(int m)
    {
        switch (m) {
        case 0xff: fail("jpeg: expected marker");
        case 0xdd:   // DRI
            if (get16() != 4) fail("jpeg: bad DRI length");
            restart_interval = get16();
            return;
        case 0xdb: {   // DQT
            int L = get16() - 2;
            while (L > 0) {
                int q = get8();
                int prec = q >> 4, t = q & 15;
                if (prec > 1 || t > 3) fail("jpeg: bad DQT");
                for (int i = 0; i < 64; ++i) dequant[t][kDezigzag[i]] = (uint16_t)(prec ? get16() : get8());
                L -= prec ? 129 : 65;
            }
            if (L != 0) fail("jpeg: bad DQT length");
            return;
        }
        case 0xc4: {   // DHT
            int L = get16() - 2;
            while (L > 0) {
                int sizes[16], n = 0;
                int q = get8();
                int tc = q >> 4, th = q & 15;
                if (tc > 1 || th > 3) fail("jpeg: bad DHT header");
                for (int i = 0; i < 16; ++i) { sizes[i] = get8(); n += sizes[i]; }
                if (n > 256) fail("jpeg: bad DHT header");
                L -= 17;
                Huffman& hf = tc == 0 ? hdc[th] : hac[th];
                hf.build(sizes);
                for (int i = 0; i < n; ++i) hf.values[i] = (uint8_t)get8();
                L -= n;
            }
            if (L != 0) fail("jpeg: bad DHT length");
            return;
        }
        default: break;
        }
        if ((m >= 0xe0 && m <= 0xef) || m == 0xfe) {   // APPn / COM
            int L = get16();
            if (L < 2) fail("jpeg: bad segment length");
            L -= 2;
            if (m == 0xe0 && L >= 5) {
                static const char tag[5] = {'J', 'F', 'I', 'F', '\0'};
                bool ok = true;
                for (int i = 0; i < 5; ++i) if (get8() != tag[i]) ok = false;
                L -= 5;
                if (ok) jfif = true;
            } else if (m == 0xee && L >= 12) {
                static const char tag[6] = {'A', 'd', 'o', 'b', 'e', '\0'};
                bool ok = true;
                for (int i = 0; i < 6; ++i) if (get8() != tag[i]) ok = false;
                L -= 6;
                if (ok) {
                    get8(); get16(); get16();         // version, flags0, flags1
                    app14_transform = get8();
                    L -= 6;
                }
            }
            if (L > end - p) fail("jpeg: truncated segment");
            p += L;
            return;
        }
        fail("jpeg: unknown marker");
    }

    int next_marker()
    {
        if (marker != 0xff) { int m = marker; marker = 0xff; return m; }
        int x = get8();
        if (x != 0xff) return 0xff;
        while (x == 0xff) x = get8();
        return x;
    }

    void frame_header()
    {
        int Lf = get16();
        if (Lf < 11) fail("jpeg: bad SOF length");
        if (get8() != 8) fail("jpeg: only 8-bit samples");
        H = get16(); W = get16();
        if (H == 0 || W == 0) fail("jpeg: zero-sized image");
        ncomp = get8();
        if (ncomp != 1 && ncomp != 3) fail(ncomp == 4 ? "jpeg: CMYK / YCCK not supported" : "jpeg: bad component count");
        if (Lf != 8 + 3 * ncomp) fail("jpeg: bad SOF length");
        rgb = 0;
        for (int i = 0; i < ncomp; ++i) {
            static const int ids[3] = {'R', 'G', 'B'};
            comp[i].id = get8();
            if (ncomp == 3 && comp[i].id == ids[i]) ++rgb;
            int q = get8();
            comp[i].h = q >> 4; comp[i].v = q & 15;
            if (comp[i].h < 1 || comp[i].h > 4 || comp[i].v < 1 || comp[i].v > 4) fail("jpeg: bad sampling factors");
            comp[i].tq = get8();
            if (comp[i].tq > 3) fail("jpeg: bad quantisation table index");
        }
        hmax = vmax = 1;
        for (int i = 0; i < ncomp; ++i) { if (comp[i].h > hmax) hmax = comp[i].h; if (comp[i].v > vmax) vmax = comp[i].v; }
        for (int i = 0; i < ncomp; ++i) if (hmax % comp[i].h != 0 || vmax % comp[i].v != 0) fail("jpeg: fractional sampling ratios not supported");
        mcu_w = hmax * 8; mcu_h = vmax * 8;
        mcux = (W + mcu_w - 1) / mcu_w; mcuy = (H + mcu_h - 1) / mcu_h;
        for (int i = 0; i < ncomp; ++i) {
            Comp& c = comp[i];
            c.x = (W * c.h + hmax - 1) / hmax;
            c.y = (H * c.v + vmax - 1) / vmax;
            c.w2 = mcux * c.h * 8;
            c.h2 = mcuy * c.v * 8;
            c.data.assign((size_t)c.w2 * c.h2, 0);
            if (progressive) {
                c.coeff_w = c.w2 / 8; c.coeff_h = c.h2 / 8;
                c.coeff.assign((size_t)c.w2 * c.h2, 0);
            }
        }
    }

    void scan_header()
    {
        int Ls = get16();
        scan_n = get8();
        if (scan_n < 1 || scan_n > 4 || scan_n > ncomp) fail("jpeg: bad SOS component count");
        if (Ls != 6 + 2 * scan_n) fail("jpeg: bad SOS length");
        for (int i = 0; i < scan_n; ++i) {
            int id = get8(), q = get8(), which = 0;
            for (; which < ncomp; ++which) if (comp[which].id == id) break;
            if (which == ncomp) fail("jpeg: scan names an unknown component");
            comp[which].hd = q >> 4; comp[which].ha = q & 15;
            if (comp[which].hd > 3 || comp[which].ha > 3) fail("jpeg: bad huffman table index");
            order[i] = which;
        }
        spec_start = get8(); spec_end = get8();
        int aa = get8();
        succ_high = aa >> 4; succ_low = aa & 15;
        if (progressive) {
            if (spec_start > 63 || spec_end > 63 || spec_start > spec_end || succ_high > 13 || succ_low > 13) fail("jpeg: bad SOS");
        } else {
            if (spec_start != 0 || succ_high != 0 || succ_low != 0) fail("jpeg: bad SOS");
            spec_end = 63;
        }
        for (int i = 0; i < scan_n; ++i) {
            const Comp& c = comp[order[i]];
            if ((!progressive || spec_start == 0) && !(progressive && succ_high) && !hdc[c.hd].present) fail("jpeg: missing DC huffman table");
            if ((!progressive || spec_start != 0) && !hac[c.ha].present) fail("jpeg: missing AC huffman table");
        }
    }

    void decode_all()
    {
        if (get8() != 0xff || get8() != 0xd8) fail("jpeg: no SOI");
        int m = next_marker();
        while (!(m == 0xc0 || m == 0xc1 || m == 0xc2)) {
            if (m >= 0xc3 && m <= 0xcf && m != 0xc4 && m != 0xc8 && m != 0xcc) fail("jpeg: unsupported coding process (lossless / hierarchical / arithmetic)");
            process_marker(m);
            m = next_marker();
            while (m == 0xff) { if (eof()) fail("jpeg: no SOF"); m = next_marker(); }
        }
        progressive = m == 0xc2;
        frame_header();
        m = next_marker();
        while (m != 0xd9) {
            if (m == 0xda) {
                scan_header();
                entropy_coded_data();
                if (marker == 0xff) {
                    // padding after the entropy-coded segment: find the next marker
                    while (!eof()) {
                        int x = get8();
                        if (x == 0xff) { int y = get8(); while (y == 0xff) y = get8(); if (y != 0) { marker = y; break; } }
                    }
                    if (marker == 0xff) break;   // ran out of data: treat as EOI
                }
            } else if (m == 0xdc) {   // DNL
                int Ld = get16(), NL = get16();
                if (Ld != 4 || NL != H) fail("jpeg: bad DNL");
            } else if (m == 0xff) {
                if (eof()) break;
            } else if (!is_restart(m)) process_marker(m);
            m = next_marker();
        }
        if (progressive) finish_progressive();
    }
};

// chroma up-sampling rows, exactly stb_image's resample_row_* (3:1 triangle filter, rounding as written)
const uint8_t* resample_1(uint8_t*, const uint8_t* in_near, const uint8_t*, int, int) { return in_near; }
const uint8_t* resample_v2(uint8_t* out, const uint8_t* in_near, const uint8_t* in_far, int w, int)
{
    for (int i = 0; i < w; ++i) out[i] = (uint8_t)((3 * in_near[i] + in_far[i] + 2) >> 2);
    return out;
}
const uint8_t* resample_h2(uint8_t* out, const uint8_t* in_near, const uint8_t*, int w, int)
{
    const uint8_t* input = in_near;
    if (w == 1) { out[0] = out[1] = input[0]; return out; }
    out[0] = input[0];
    out[1] = (uint8_t)((input[0] * 3 + input[1] + 2) >> 2);
    int i;
    for (i = 1; i < w - 1; ++i) {
        int n = 3 * input[i] + 2;
        out[i * 2 + 0] = (uint8_t)((n + input[i - 1]) >> 2);
        out[i * 2 + 1] = (uint8_t)((n + input[i + 1]) >> 2);
    }
    out[i * 2 + 0] = (uint8_t)((input[w - 2] * 3 + input[w - 1] + 2) >> 2);   // sic: stb weights in[w-2] by 3 here
    out[i * 2 + 1] = input[w - 1];
    return out;
}
const uint8_t* resample_hv2(uint8_t* out, const uint8_t* in_near, const uint8_t* in_far, int w, int)
{
    if (w == 1) { out[0] = out[1] = (uint8_t)((3 * in_near[0] + in_far[0] + 2) >> 2); return out; }
    int t1 = 3 * in_near[0] + in_far[0], t0;
    out[0] = (uint8_t)((t1 + 2) >> 2);
    for (int i = 1; i < w; ++i) {
        t0 = t1;
        t1 = 3 * in_near[i] + in_far[i];
        out[i * 2 - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
        out[i * 2] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
    }
    out[w * 2 - 1] = (uint8_t)((t1 + 2) >> 2);
    return out;
}
const uint8_t* resample_generic(uint8_t* out, const uint8_t* in_near, const uint8_t*, int w, int hs)
{
    for (int i = 0; i < w; ++i) for (int j = 0; j < hs; ++j) out[i * hs + j] = in_near[i];
    return out;
}

#define IMGDEC_FLOAT2FIXED(x) (((int)((x) * 4096.0f + 0.5f)) << 8)
void ycbcr_to_rgba_row(uint8_t* out, const uint8_t* y, const uint8_t* pcb, const uint8_t* pcr, int count)
{
    for (int i = 0; i < count; ++i) {
        int y_fixed = (y[i] << 20) + (1 << 19);
        int cr = pcr[i] - 128, cb = pcb[i] - 128;
        int r = y_fixed + cr * IMGDEC_FLOAT2FIXED(1.40200f);
        int g = y_fixed + (cr * -IMGDEC_FLOAT2FIXED(0.71414f)) + (int)((unsigned)(cb * -IMGDEC_FLOAT2FIXED(0.34414f)) & 0xffff0000u);
        int b = y_fixed + cb * IMGDEC_FLOAT2FIXED(1.77200f);
        r >>= 20; g >>= 20; b >>= 20;
        out[0] = clamp8(r); out[1] = clamp8(g); out[2] = clamp8(b); out[3] = 255;
        out += 4;
    }
}

bool is_jpeg(const uint8_t* b, size_t n) { return n >= 3 && b[0] == 0xff && b[1] == 0xd8 && b[2] == 0xff; }

void decode_jpeg(const uint8_t* bytes, size_t n, int& w, int& h, std::vector<uint8_t>& rgba)
{
    Jpeg* j = new Jpeg();
    struct Guard { Jpeg* j; ~Guard() { delete j; } } guard{j};
    j->p = bytes; j->end = bytes + n;
    std::memset(j->dequant, 0, sizeof(j->dequant));
    j->decode_all();
    w = j->W; h = j->H;
    rgba.assign((size_t)w * h * 4, 255);

    typedef const uint8_t* (*resample_fn)(uint8_t*, const uint8_t*, const uint8_t*, int, int);
    struct Res { resample_fn fn; const uint8_t* line0; const uint8_t* line1; int hs, vs, w_lores, ystep, ypos; std::vector<uint8_t> linebuf; } res[3];
    const int nc = j->ncomp;
    for (int k = 0; k < nc; ++k) {
        Res& r = res[k];
        r.hs = j->hmax / j->comp[k].h; r.vs = j->vmax / j->comp[k].v;
        r.ystep = r.vs >> 1;
        r.w_lores = (w + r.hs - 1) / r.hs;
        r.ypos = 0;
        r.line0 = r.line1 = j->comp[k].data.data();
        r.linebuf.assign((size_t)w + 3 + 8, 0);
        if (r.hs == 1 && r.vs == 1) r.fn = resample_1;
        else if (r.hs == 1 && r.vs == 2) r.fn = resample_v2;
        else if (r.hs == 2 && r.vs == 1) r.fn = resample_h2;
        else if (r.hs == 2 && r.vs == 2) r.fn = resample_hv2;
        else r.fn = resample_generic;
    }
    const bool is_rgb = nc == 3 && (j->rgb == 3 || (j->app14_transform == 0 && !j->jfif));
    const uint8_t* co[3] = {nullptr, nullptr, nullptr};
    for (int y = 0; y < h; ++y) {
        uint8_t* out = rgba.data() + (size_t)y * w * 4;
        for (int k = 0; k < nc; ++k) {
            Res& r = res[k];
            const bool y_bot = r.ystep >= (r.vs >> 1);
            co[k] = r.fn(r.linebuf.data(), y_bot ? r.line1 : r.line0, y_bot ? r.line0 : r.line1, r.w_lores, r.hs);
            if (++r.ystep >= r.vs) {
                r.ystep = 0;
                r.line0 = r.line1;
                if (++r.ypos < j->comp[k].y) r.line1 += j->comp[k].w2;
            }
        }
        if (nc == 3) {
            if (is_rgb) for (int i = 0; i < w; ++i) { out[4 * i] = co[0][i]; out[4 * i + 1] = co[1][i]; out[4 * i + 2] = co[2][i]; out[4 * i + 3] = 255; }
            else ycbcr_to_rgba_row(out, co[0], co[1], co[2], w);
        } else {
            for (int i = 0; i < w; ++i) { out[4 * i] = out[4 * i + 1] = out[4 * i + 2] = co[0][i]; out[4 * i + 3] = 255; }
        }
    }
}

// ================================================================================================ Radiance HDR
bool is_hdr(const uint8_t* b, size_t n) { return (n >= 11 && !std::memcmp(b, "#?RADIANCE\n", 11)) || (n >= 7 && !std::memcmp(b, "#?RGBE\n", 7)); }

void decode_hdr(const uint8_t* b, size_t n, int& w, int& h, std::vector<uint8_t>& rgba)
{
    size_t p = 0;
    auto line = [&]() -> std::string {
        std::string s;
        while (p < n && b[p] != '\n') { if (s.size() < 1023) s += (char)b[p]; ++p; }
        if (p < n) ++p;
        return s;
    };
    line();   // magic
    bool valid = false;
    while (true) {
        std::string s = line();
        if (s.empty()) break;
        if (s == "FORMAT=32-bit_rle_rgbe") valid = true;
        if (p >= n) break;
    }
    if (!valid) fail("hdr: unsupported format");
    std::string dims = line();
    int hh = 0, ww = 0;
    if (std::sscanf(dims.c_str(), "-Y %d +X %d", &hh, &ww) != 2 || hh <= 0 || ww <= 0) fail("hdr: unsupported data layout");
    w = ww; h = hh;
    std::vector<float> lin((size_t)w * h * 3);
    auto convert = [](float* out, const uint8_t* in) {
        if (in[3] != 0) {
            float f1 = (float)std::ldexp(1.0f, in[3] - (int)(128 + 8));
            out[0] = in[0] * f1; out[1] = in[1] * f1; out[2] = in[2] * f1;
        } else out[0] = out[1] = out[2] = 0.0f;
    };
    auto need = [&](size_t k) { if (p + k > n) fail("hdr: truncated"); };
    bool flat = w < 8 || w >= 32768;
    std::vector<uint8_t> scan((size_t)w * 4);
    for (int y = 0; y < h && !flat; ++y) {
        need(4);
        int c1 = b[p], c2 = b[p + 1], len = b[p + 2];
        if (c1 != 2 || c2 != 2 || (len & 0x80)) {
            if (y != 0) fail("hdr: mixed scanline encodings");
            flat = true;   // not run-length encoded at all
            break;
        }
        len = (len << 8) | b[p + 3];
        p += 4;
        if (len != w) fail("hdr: invalid decoded scanline length");
        for (int k = 0; k < 4; ++k) {
            int i = 0;
            while (i < w) {
                need(1);
                int count = b[p++];
                if (count > 128) {
                    count -= 128;
                    need(1);
                    if (count == 0 || count > w - i) fail("hdr: corrupt run");
                    uint8_t v = b[p++];
                    for (int z = 0; z < count; ++z) scan[(size_t)(i++) * 4 + k] = v;
                } else {
                    if (count == 0 || count > w - i) fail("hdr: corrupt run");
                    need((size_t)count);
                    for (int z = 0; z < count; ++z) scan[(size_t)(i++) * 4 + k] = b[p++];
                }
            }
        }
        for (int i = 0; i < w; ++i) convert(&lin[((size_t)y * w + i) * 3], &scan[(size_t)i * 4]);
    }
    if (flat) {
        need((size_t)w * h * 4);
        for (size_t i = 0; i < (size_t)w * h; ++i) convert(&lin[i * 3], b + p + i * 4);
    }
    // stbi hdr-to-ldr with its default gamma 2.2 and scale 1
    rgba.assign((size_t)w * h * 4, 255);
    const float gamma_i = 1.0f / 2.2f, scale_i = 1.0f;
    for (size_t i = 0; i < (size_t)w * h; ++i) for (int k = 0; k < 3; ++k) {
        float z = (float)std::pow(lin[i * 3 + k] * scale_i, gamma_i) * 255 + 0.5f;
        if (z < 0) z = 0;
        if (z > 255) z = 255;
        rgba[i * 4 + k] = (uint8_t)(int)z;
    }
}

} // namespace

bool decode_rgba8(const uint8_t* bytes, size_t n, int& w, int& h, std::vector<uint8_t>& rgba, std::string* err)
{
    try {
        if (is_png(bytes, n)) decode_png(bytes, n, w, h, rgba);
        else if (is_jpeg(bytes, n)) decode_jpeg(bytes, n, w, h, rgba);
        else if (is_hdr(bytes, n)) decode_hdr(bytes, n, w, h, rgba);
        else fail("unknown image type (PNG, JPEG and Radiance HDR are supported)");
        return true;
    } catch (const std::exception& e) {
        if (err) *err = e.what();
        rgba.clear(); w = h = 0;
        return false;
    }
}

bool load_rgba8(const char* path, int& w, int& h, std::vector<uint8_t>& rgba, std::string* err)
{
    FILE* f = std::fopen(path, "rb");
    if (!f) { if (err) *err = std::string("cannot open ") + path; return false; }
    std::vector<uint8_t> bytes;
    uint8_t buf[65536];
    size_t k;
    while ((k = std::fread(buf, 1, sizeof(buf), f)) > 0) bytes.insert(bytes.end(), buf, buf + k);
    std::fclose(f);
    return decode_rgba8(bytes.data(), bytes.size(), w, h, rgba, err);
}

} // namespace imgdec
