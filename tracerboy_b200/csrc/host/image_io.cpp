// image_io.cpp — image files out of the library: the role of the reference's frame capture
// (D3D12App.cpp:341-363: CaptureTexture of the back buffer + SaveToWICFile as PNG, DirectXTex, third party) plus a
// float format for the accumulation buffer. Self-contained writers, no third-party code:
//   .png  8-bit RGBA, zlib stream made of stored (uncompressed) deflate blocks
//   .exr  OpenEXR 2.0 scanline file, NO_COMPRESSION, 32-bit float channels A B G R (or B G R)
//   .pfm  Portable Float Map, little endian, bottom row first
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "scene.h"

namespace tb {

namespace {
uint32_t crc_table[256];
bool crc_ready = false;
uint32_t crc32(uint32_t crc, const uint8_t* p, size_t n) {
    if (!crc_ready) {
        for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1; crc_table[i] = c; }
        crc_ready = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; i++) crc = crc_table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return ~crc;
}
void be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
void chunk(std::vector<uint8_t>& out, const char* type, const std::vector<uint8_t>& data) {
    be32(out, (uint32_t)data.size());
    size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), data.begin(), data.end());
    be32(out, crc32(0, out.data() + start, out.size() - start));
}
bool write_all(const std::string& path, const void* p, size_t n, std::string& err) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { err = "cannot open for write: " + path; return false; }
    bool ok = fwrite(p, 1, n, f) == n;
    fclose(f);
    if (!ok) err = "short write: " + path;
    return ok;
}
} // namespace

bool save_png_rgba8(const std::string& path, const uint8_t* rgba, uint32_t w, uint32_t h, std::string& err) {
    std::vector<uint8_t> raw; // filter byte 0 + row
    raw.reserve((size_t)h * (4 * w + 1));
    for (uint32_t y = 0; y < h; y++) { raw.push_back(0); raw.insert(raw.end(), rgba + (size_t)y * w * 4, rgba + (size_t)(y + 1) * w * 4); }
    std::vector<uint8_t> z;
    z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, b = 0; // adler32
    for (size_t off = 0; off < raw.size() || off == 0;) {
        size_t n = raw.size() - off < 65535 ? raw.size() - off : 65535;
        z.push_back(off + n >= raw.size() ? 1 : 0);
        z.push_back(n & 0xff); z.push_back(n >> 8); z.push_back(~n & 0xff); z.push_back((~n >> 8) & 0xff);
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
        for (size_t i = 0; i < n; i++) { a = (a + raw[off + i]) % 65521u; b = (b + a) % 65521u; }
        off += n;
        if (n == 0) break;
    }
    be32(z, (b << 16) | a);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr;
    be32(ihdr, w); be32(ihdr, h);
    ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(out, "IHDR", ihdr);
    chunk(out, "sRGB", std::vector<uint8_t>{0}); // WIC_FLAGS_FORCE_SRGB in the reference's capture
    chunk(out, "IDAT", z);
    chunk(out, "IEND", {});
    return write_all(path, out.data(), out.size(), err);
}

// channels: 3 (rgb) or 4 (rgba) interleaved floats per pixel, top row first
bool save_exr_f32(const std::string& path, const float* px, uint32_t w, uint32_t h, int channels, std::string& err) {
    std::vector<uint8_t> o;
    auto u32 = [&](uint32_t x) { for (int i = 0; i < 4; i++) o.push_back((x >> (8 * i)) & 0xff); };
    auto u64 = [&](uint64_t x) { for (int i = 0; i < 8; i++) o.push_back((x >> (8 * i)) & 0xff); };
    auto str = [&](const char* s) { o.insert(o.end(), s, s + strlen(s) + 1); };
    auto f32 = [&](float f) { uint32_t u; memcpy(&u, &f, 4); u32(u); };
    auto attr = [&](const char* name, const char* type, uint32_t size) { str(name); str(type); u32(size); };
    u32(20000630u); u32(2u);
    const char* names4[4] = {"A", "B", "G", "R"};
    const char* names3[3] = {"B", "G", "R"};
    const char** names = channels == 4 ? names4 : names3;
    attr("channels", "chlist", (uint32_t)(18 * channels + 1));
    for (int c = 0; c < channels; c++) { str(names[c]); u32(2u /*FLOAT*/); o.push_back(0); o.push_back(0); o.push_back(0); o.push_back(0); u32(1); u32(1); }
    o.push_back(0);
    attr("compression", "compression", 1); o.push_back(0);
    attr("dataWindow", "box2i", 16); u32(0); u32(0); u32(w - 1); u32(h - 1);
    attr("displayWindow", "box2i", 16); u32(0); u32(0); u32(w - 1); u32(h - 1);
    attr("lineOrder", "lineOrder", 1); o.push_back(0);
    attr("pixelAspectRatio", "float", 4); f32(1.0f);
    attr("screenWindowCenter", "v2f", 8); f32(0.0f); f32(0.0f);
    attr("screenWindowWidth", "float", 4); f32(1.0f);
    o.push_back(0);
    const uint64_t lineBytes = (uint64_t)w * 4 * channels;
    uint64_t first = o.size() + 8ull * h;
    for (uint32_t y = 0; y < h; y++) u64(first + y * (8 + lineBytes));
    const int src4[4] = {3, 2, 1, 0}, src3[3] = {2, 1, 0}; // file channel order (alphabetical) -> index in the pixel
    const int* src = channels == 4 ? src4 : src3;
    for (uint32_t y = 0; y < h; y++) {
        u32(y); u32((uint32_t)lineBytes);
        for (int c = 0; c < channels; c++)
            for (uint32_t x = 0; x < w; x++) f32(px[((size_t)y * w + x) * channels + src[c]]);
    }
    return write_all(path, o.data(), o.size(), err);
}

bool save_pfm_rgb(const std::string& path, const float* px, uint32_t w, uint32_t h, int channels, std::string& err) {
    std::string hdr = "PF\n" + std::to_string(w) + " " + std::to_string(h) + "\n-1.0\n";
    std::vector<uint8_t> o(hdr.begin(), hdr.end());
    size_t at = o.size();
    o.resize(at + (size_t)w * h * 12);
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++)
            memcpy(o.data() + at + (((size_t)(h - 1 - y) * w + x) * 12), px + ((size_t)y * w + x) * channels, 12);
    return write_all(path, o.data(), o.size(), err);
}

} // namespace tb
