// image_decode.cpp — image textures other than Radiance .hdr: PNG and TGA, decoded to what the reference's loaders
// hand to the GPU (TracerBoy::InitializeTexture, TracerBoy.cpp:2186-2246: DirectX::LoadFromTGAFile for .tga,
// DirectX::LoadFromWICFile(WIC_FLAGS_NONE) for everything else, then CreateTextureEx with the format the loader chose).
//
// The format decides what a shader sees, so the loaders' rules are followed (DirectXTex/DirectXTexWIC.cpp:34-81 pixel
// format table, :582-645 sRGB metadata; DirectXTex/DirectXTexTGA.cpp):
//   * 8-bit RGB / RGBA / palette PNG, 24 / 32-bit TGA   -> R8G8B8A8_UNORM        (tb::Image format 1)
//   * the same with an sRGB chunk, or gAMA == 45455      -> R8G8B8A8_UNORM_SRGB   (format 2: the sampler linearises)
//   * 8-bit (and 1/2/4-bit) greyscale                    -> R8_UNORM: the shader reads (v, 0, 0, 1); stored as RGBA8
//   * 16 bits per channel                                -> R16G16B16A16_UNORM    (stored as float4 = v / 65535, format 0)
//   * greyscale + alpha PNG is converted by WIC to 32bppRGBA (v, v, v, a)
// Self-contained: inflate (RFC 1951), the PNG filters (RFC 2083 §6) and Adam7 are implemented here; there is no zlib
// or libpng dependency. JPEG / BMP / DDS are not implemented (TB_ERR_NOT_IMPL through the importer's error text).
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include "scene.h"

namespace tb {
namespace {

bool read_file(const std::string& path, std::vector<uint8_t>& out, std::string& err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open: " + path; return false; }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (n < 0 || n > (1l << 31)) { fclose(f); err = "not a readable image file: " + path; return false; } // e.g. a directory
    out.resize((size_t)n);
    bool ok = n >= 0 && fread(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    if (!ok) err = "short read: " + path;
    return ok;
}

// ------------------------------------------------------------------ inflate
struct BitReader {
    const uint8_t* p; size_t n, pos = 0; uint32_t acc = 0; int bits = 0;
    BitReader(const uint8_t* d, size_t len) : p(d), n(len) {}
    uint32_t get(int count) {
        while (bits < count) { if (pos >= n) throw std::runtime_error("truncated deflate stream"); acc |= (uint32_t)p[pos++] << bits; bits += 8; }
        uint32_t v = acc & ((count < 32 ? (1u << count) : 0u) - 1u);
        acc >>= count; bits -= count;
        return v;
    }
    void align() { acc = 0; bits = 0; }
};
struct Huffman {
    uint16_t count[16], symbol[288];
    void build(const uint8_t* lengths, int n) {
        memset(count, 0, sizeof(count));
        for (int i = 0; i < n; i++) count[lengths[i]]++;
        count[0] = 0;
        uint16_t offs[16];
        offs[1] = 0;
        for (int i = 1; i < 15; i++) offs[i + 1] = offs[i] + count[i];
        for (int i = 0; i < n; i++) if (lengths[i]) symbol[offs[lengths[i]]++] = (uint16_t)i;
    }
    int decode(BitReader& br) const {
        int code = 0, first = 0, index = 0;
        for (int len = 1; len <= 15; len++) {
            code |= (int)br.get(1);
            int c = count[len];
            if (code - c < first) return symbol[index + (code - first)];
            index += c; first += c; first <<= 1; code <<= 1;
        }
        throw std::runtime_error("bad Huffman code");
    }
};
void inflate(const uint8_t* data, size_t n, std::vector<uint8_t>& out) {
    static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    if (n < 2) throw std::runtime_error("truncated zlib stream");
    BitReader br(data + 2, n - 2); // zlib header: CMF, FLG (the Adler-32 trailer is not checked; the PNG CRCs are not either)
    bool last;
    do {
        last = br.get(1) != 0;
        uint32_t type = br.get(2);
        if (type == 0) {
            br.align();
            if (br.pos + 4 > br.n) throw std::runtime_error("truncated stored block");
            uint32_t len = br.p[br.pos] | (br.p[br.pos + 1] << 8);
            br.pos += 4;
            if (br.pos + len > br.n) throw std::runtime_error("truncated stored block");
            out.insert(out.end(), br.p + br.pos, br.p + br.pos + len);
            br.pos += len;
        } else if (type == 1 || type == 2) {
            Huffman lit, dist;
            uint8_t lengths[320];
            if (type == 1) {
                for (int i = 0; i < 144; i++) lengths[i] = 8;
                for (int i = 144; i < 256; i++) lengths[i] = 9;
                for (int i = 256; i < 280; i++) lengths[i] = 7;
                for (int i = 280; i < 288; i++) lengths[i] = 8;
                lit.build(lengths, 288);
                for (int i = 0; i < 30; i++) lengths[i] = 5;
                dist.build(lengths, 30);
            } else {
                int nlen = (int)br.get(5) + 257, ndist = (int)br.get(5) + 1, ncode = (int)br.get(4) + 4;
                if (nlen > 286 || ndist > 30) throw std::runtime_error("bad deflate header");
                uint8_t cl[19] = {0};
                for (int i = 0; i < ncode; i++) cl[order[i]] = (uint8_t)br.get(3);
                Huffman lencode;
                lencode.build(cl, 19);
                int idx = 0;
                while (idx < nlen + ndist) {
                    int sym = lencode.decode(br);
                    if (sym < 16) lengths[idx++] = (uint8_t)sym;
                    else {
                        uint8_t prev = 0; int rep;
                        if (sym == 16) { if (!idx) throw std::runtime_error("bad deflate lengths"); prev = lengths[idx - 1]; rep = 3 + (int)br.get(2); }
                        else if (sym == 17) rep = 3 + (int)br.get(3);
                        else rep = 11 + (int)br.get(7);
                        if (idx + rep > nlen + ndist) throw std::runtime_error("bad deflate lengths");
                        while (rep--) lengths[idx++] = prev;
                    }
                }
                lit.build(lengths, nlen);
                dist.build(lengths + nlen, ndist);
            }
            while (true) {
                int sym = lit.decode(br);
                if (sym < 256) out.push_back((uint8_t)sym);
                else if (sym == 256) break;
                else {
                    sym -= 257;
                    if (sym >= 29) throw std::runtime_error("bad length symbol");
                    uint32_t len = lbase[sym] + br.get(lext[sym]);
                    int ds = dist.decode(br);
                    if (ds >= 30) throw std::runtime_error("bad distance symbol");
                    uint32_t d = dbase[ds] + br.get(dext[ds]);
                    if (d > out.size()) throw std::runtime_error("deflate distance too far back");
                    size_t from = out.size() - d;
                    for (uint32_t i = 0; i < len; i++) out.push_back(out[from + i]);
                }
            }
        } else throw std::runtime_error("bad deflate block type");
    } while (!last);
}

// ---------------------------------------------------------------------- PNG
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline int paeth(int a, int b, int c) {
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
// un-filter `rows` scanlines of `rowBytes` bytes (each preceded by its filter byte) in place; returns bytes consumed
size_t unfilter(uint8_t* src, size_t avail, uint32_t rows, size_t rowBytes, uint32_t bpp, std::vector<uint8_t>& out) {
    out.assign((size_t)rows * rowBytes, 0);
    if ((rowBytes + 1) * (size_t)rows > avail) throw std::runtime_error("truncated PNG image data");
    for (uint32_t y = 0; y < rows; y++) {
        const uint8_t* in = src + (size_t)y * (rowBytes + 1);
        uint8_t* cur = out.data() + (size_t)y * rowBytes;
        const uint8_t* up = y ? cur - rowBytes : nullptr;
        const uint8_t ft = in[0];
        in++;
        for (size_t x = 0; x < rowBytes; x++) {
            int a = x >= bpp ? cur[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0, v = in[x];
            switch (ft) {
            case 0: break;
            case 1: v += a; break;
            case 2: v += b; break;
            case 3: v += (a + b) >> 1; break;
            case 4: v += paeth(a, b, c); break;
            default: throw std::runtime_error("bad PNG filter type");
            }
            cur[x] = (uint8_t)v;
        }
    }
    return (rowBytes + 1) * (size_t)rows;
}

bool decode_png(const std::vector<uint8_t>& file, Image& img, bool* hasAlpha, std::string& err) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 8 || memcmp(file.data(), sig, 8) != 0) { err = "not a PNG file"; return false; }
    uint32_t w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    bool srgbChunk = false, gama22 = false, gotHeader = false;
    size_t pos = 8;
    while (pos + 12 <= file.size()) {
        uint32_t len = be32(&file[pos]);
        const uint8_t* type = &file[pos + 4];
        const uint8_t* data = &file[pos + 8];
        if (pos + 12 + (size_t)len > file.size()) { err = "truncated PNG chunk"; return false; }
        if (!memcmp(type, "IHDR", 4) && len >= 13) {
            w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
            gotHeader = true;
        } else if (!memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
        else if (!memcmp(type, "tRNS", 4)) trns.assign(data, data + len);
        else if (!memcmp(type, "sRGB", 4)) srgbChunk = true;
        else if (!memcmp(type, "gAMA", 4) && len >= 4) gama22 = be32(data) == 45455;
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!memcmp(type, "IEND", 4)) break;
        pos += 12 + (size_t)len;
    }
    if (!gotHeader || w == 0 || h == 0 || w > 65536 || h > 65536) { err = "bad PNG header"; return false; }
    const uint32_t channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!channels || (depth != 1 && depth != 2 && depth != 4 && depth != 8 && depth != 16) || (ctype == 3 && depth == 16) ||
        ((ctype == 2 || ctype == 4 || ctype == 6) && depth < 8) || interlace > 1) { err = "unsupported PNG colour type / bit depth"; return false; }
    std::vector<uint8_t> raw;
    try { inflate(idat.data(), idat.size(), raw); } catch (const std::exception& e) { err = std::string("PNG: ") + e.what(); return false; }
    const uint32_t bitsPerPixel = channels * depth, bpp = bitsPerPixel >= 8 ? bitsPerPixel / 8 : 1;
    static const uint32_t xs[7] = {0, 4, 0, 2, 0, 1, 0}, ys[7] = {0, 0, 4, 0, 2, 0, 1}, dxs[7] = {8, 8, 4, 4, 2, 2, 1}, dys[7] = {8, 8, 8, 4, 4, 2, 2}; // Adam7
    auto pass_width = [&](int p) { return w > xs[p] ? (w - xs[p] + dxs[p] - 1) / dxs[p] : 0u; };
    auto pass_height = [&](int p) { return h > ys[p] ? (h - ys[p] + dys[p] - 1) / dys[p] : 0u; };
    // the header is not trusted with an allocation: the inflated stream has to hold every scanline it announces
    size_t expected = 0;
    if (!interlace) expected = (size_t)h * (1 + ((size_t)w * bitsPerPixel + 7) / 8);
    else for (int p = 0; p < 7; p++) if (pass_width(p) && pass_height(p)) expected += (size_t)pass_height(p) * (1 + ((size_t)pass_width(p) * bitsPerPixel + 7) / 8);
    if (raw.size() < expected) { err = "PNG: image data shorter than the header's dimensions"; return false; }
    // samples[(y * w + x) * channels + c] as 16-bit values at the file's bit depth
    std::vector<uint16_t> samples((size_t)w * h * channels);
    auto unpack = [&](const std::vector<uint8_t>& rows, uint32_t pw, uint32_t ph, uint32_t x0, uint32_t y0, uint32_t dx, uint32_t dy) {
        const size_t rowBytes = ((size_t)pw * bitsPerPixel + 7) / 8;
        for (uint32_t y = 0; y < ph; y++)
            for (uint32_t x = 0; x < pw; x++)
                for (uint32_t c = 0; c < channels; c++) {
                    const uint8_t* r = rows.data() + (size_t)y * rowBytes;
                    uint16_t v;
                    if (depth == 16) v = (uint16_t)((r[(x * channels + c) * 2] << 8) | r[(x * channels + c) * 2 + 1]);
                    else if (depth == 8) v = r[x * channels + c];
                    else { const size_t bit = (size_t)x * depth; v = (r[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u); }
                    samples[((size_t)(y0 + y * dy) * w + (x0 + x * dx)) * channels + c] = v;
                }
    };
    try {
        std::vector<uint8_t> rows;
        if (!interlace) {
            unfilter(raw.data(), raw.size(), h, ((size_t)w * bitsPerPixel + 7) / 8, bpp, rows);
            unpack(rows, w, h, 0, 0, 1, 1);
        } else { // Adam7
            size_t off = 0;
            for (int p = 0; p < 7; p++) {
                const uint32_t pw = pass_width(p), ph = pass_height(p);
                if (!pw || !ph) continue;
                off += unfilter(raw.data() + off, raw.size() - off, ph, ((size_t)pw * bitsPerPixel + 7) / 8, bpp, rows);
                unpack(rows, pw, ph, xs[p], ys[p], dxs[p], dys[p]);
            }
        }
    } catch (const std::exception& e) { err = std::string("PNG: ") + e.what(); return false; }

    const size_t n = (size_t)w * h;
    const uint32_t maxv = (1u << depth) - 1u;
    bool alphaSeen = false;
    img.width = w; img.height = h;
    if (depth == 16) { // R16G16B16A16_UNORM (48bppRGB -> 64bppRGBA, 16bppGray -> R16_UNORM: shader reads (v, 0, 0, 1))
        img.format = 0;
        img.unorm16 = true;
        img.data.resize(n * 16);
        float* o = (float*)img.data.data();
        for (size_t i = 0; i < n; i++) {
            const uint16_t* s = &samples[i * channels];
            float r, g, b, a = 1.0f;
            if (ctype == 0) { r = s[0] / 65535.0f; g = b = 0.0f; if (trns.size() >= 2 && s[0] == ((trns[0] << 8) | trns[1])) a = 0.0f; }
            else if (ctype == 4) { r = g = b = s[0] / 65535.0f; a = s[1] / 65535.0f; }
            else { r = s[0] / 65535.0f; g = s[1] / 65535.0f; b = s[2] / 65535.0f; if (ctype == 6) a = s[3] / 65535.0f;
                   else if (trns.size() >= 6 && s[0] == ((trns[0] << 8) | trns[1]) && s[1] == ((trns[2] << 8) | trns[3]) && s[2] == ((trns[4] << 8) | trns[5])) a = 0.0f; }
            o[4 * i] = r; o[4 * i + 1] = g; o[4 * i + 2] = b; o[4 * i + 3] = a;
            alphaSeen = alphaSeen || a != 1.0f;
        }
    } else {
        img.format = (srgbChunk || gama22) ? 2u : 1u; // DirectXTexWIC.cpp:582-645: sRGB chunk, or gAMA == 1/2.2
        img.data.resize(n * 4);
        uint8_t* o = img.data.data();
        for (size_t i = 0; i < n; i++) {
            const uint16_t* s = &samples[i * channels];
            uint8_t r, g, b, a = 255;
            if (ctype == 3) {
                const uint32_t pi = s[0];
                if (3 * (size_t)pi + 2 < plte.size()) { r = plte[3 * pi]; g = plte[3 * pi + 1]; b = plte[3 * pi + 2]; } else r = g = b = 0;
                if (pi < trns.size()) a = trns[pi];
            } else if (ctype == 0) { // greyscale: 1/2/4-bit are widened to 8bppGray by WIC (value * 255 / max), -> R8_UNORM
                r = (uint8_t)((s[0] * 255u + maxv / 2) / maxv); g = b = 0;
                if (trns.size() >= 2 && s[0] == (uint16_t)((trns[0] << 8) | trns[1])) a = 0;
            } else if (ctype == 4) { r = g = b = (uint8_t)s[0]; a = (uint8_t)s[1]; }
            else { r = (uint8_t)s[0]; g = (uint8_t)s[1]; b = (uint8_t)s[2];
                   if (ctype == 6) a = (uint8_t)s[3];
                   else if (trns.size() >= 6 && s[0] == trns[1] && s[1] == trns[3] && s[2] == trns[5]) a = 0; }
            o[4 * i] = r; o[4 * i + 1] = g; o[4 * i + 2] = b; o[4 * i + 3] = a;
            alphaSeen = alphaSeen || a != 255;
        }
        if (ctype == 0) img.format = 1; // R8_UNORM has no sRGB variant (MakeSRGB leaves it alone)
    }
    if (hasAlpha) *hasAlpha = alphaSeen; // !scratchImage.IsAlphaAllOpaque()
    return true;
}

// ---------------------------------------------------------------------- TGA
bool decode_tga(const std::vector<uint8_t>& f, Image& img, bool* hasAlpha, std::string& err) {
    if (f.size() < 18) { err = "truncated TGA header"; return false; }
    const uint32_t idLen = f[0], cmapType = f[1], type = f[2], cmapLen = f[5] | (f[6] << 8), cmapBits = f[7];
    const uint32_t w = f[12] | (f[13] << 8), h = f[14] | (f[15] << 8), bits = f[16], desc = f[17];
    const bool rle = type >= 9;
    const uint32_t base = rle ? type - 8 : type;
    if (w == 0 || h == 0 || (base != 1 && base != 2 && base != 3)) { err = "unsupported TGA image type"; return false; }
    if (base == 1 && (cmapType != 1 || bits != 8 || (cmapBits != 24 && cmapBits != 32))) { err = "unsupported TGA colour map"; return false; }
    if (base == 2 && bits != 16 && bits != 24 && bits != 32) { err = "unsupported TGA bit depth"; return false; }
    if (base == 3 && bits != 8) { err = "unsupported TGA greyscale depth"; return false; }
    size_t pos = 18 + idLen;
    const uint8_t* cmap = nullptr;
    const uint32_t cmapBytes = cmapType ? cmapLen * ((cmapBits + 7) / 8) : 0;
    if (pos + cmapBytes > f.size()) { err = "truncated TGA colour map"; return false; }
    if (cmapType) { cmap = &f[pos]; pos += cmapBytes; }
    const uint32_t px = bits / 8;
    const size_t n = (size_t)w * h;
    std::vector<uint8_t> pix(n * px);
    if (!rle) {
        if (pos + n * px > f.size()) { err = "truncated TGA pixels"; return false; }
        memcpy(pix.data(), &f[pos], n * px);
    } else {
        size_t o = 0;
        while (o < n) {
            if (pos >= f.size()) { err = "truncated TGA RLE stream"; return false; }
            const uint32_t hd = f[pos++], cnt = (hd & 127u) + 1;
            if (o + cnt > n) { err = "TGA RLE packet overruns the image"; return false; }
            if (hd & 128u) {
                if (pos + px > f.size()) { err = "truncated TGA RLE stream"; return false; }
                for (uint32_t k = 0; k < cnt; k++) memcpy(&pix[(o + k) * px], &f[pos], px);
                pos += px;
            } else {
                if (pos + (size_t)cnt * px > f.size()) { err = "truncated TGA RLE stream"; return false; }
                memcpy(&pix[o * px], &f[pos], (size_t)cnt * px);
                pos += (size_t)cnt * px;
            }
            o += cnt;
        }
    }
    img.width = w; img.height = h; img.format = 1;
    img.data.resize(n * 4);
    const bool topDown = (desc & 0x20) != 0, rightLeft = (desc & 0x10) != 0;
    bool alphaSeen = false, alphaNonZero = false;
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) {
            const uint8_t* s = &pix[((size_t)y * w + x) * px];
            uint8_t r, g, b, a = 255;
            if (base == 3) { r = s[0]; g = b = 0; } // R8_UNORM
            else if (base == 1) { const uint8_t* c = cmap + (size_t)(s[0] < cmapLen ? s[0] : 0) * (cmapBits / 8); b = c[0]; g = c[1]; r = c[2]; if (cmapBits == 32) a = c[3]; }
            else if (bits == 16) { const uint32_t v = s[0] | (s[1] << 8); r = (uint8_t)((((v >> 10) & 31u) * 255u + 15u) / 31u); g = (uint8_t)((((v >> 5) & 31u) * 255u + 15u) / 31u);
                                   b = (uint8_t)(((v & 31u) * 255u + 15u) / 31u); a = (v & 0x8000u) ? 255 : 0; }
            else { b = s[0]; g = s[1]; r = s[2]; if (bits == 32) a = s[3]; }
            const uint32_t dy = topDown ? y : h - 1 - y, dx = rightLeft ? w - 1 - x : x;
            uint8_t* o = &img.data[((size_t)dy * w + dx) * 4];
            o[0] = r; o[1] = g; o[2] = b; o[3] = a;
            alphaSeen = alphaSeen || a != 255;
            alphaNonZero = alphaNonZero || a != 0;
        }
    // DirectXTexTGA.cpp: a 32-bit / 16-bit image whose alpha channel is entirely zero is treated as opaque
    if (!alphaNonZero) { for (size_t i = 0; i < n; i++) img.data[4 * i + 3] = 255; alphaSeen = false; }
    if (hasAlpha) *hasAlpha = alphaSeen;
    return true;
}

} // namespace

bool load_image_file(const std::string& path, Image& img, bool* hasAlpha, std::string& err) {
    std::string ext = path.size() >= 4 ? path.substr(path.size() - 4) : "";
    for (auto& c : ext) c = (char)tolower(c);
    if (ext == ".hdr") { if (hasAlpha) *hasAlpha = false; return load_hdr(path, img, err); }
    std::vector<uint8_t> file;
    if (!read_file(path, file, err)) return false;
    if (ext == ".tga") return decode_tga(file, img, hasAlpha, err);
    if (ext == ".png") return decode_png(file, img, hasAlpha, err);
    err = "unsupported texture format (supported: .hdr, .png, .tga): " + path;
    return false;
}

} // namespace tb
