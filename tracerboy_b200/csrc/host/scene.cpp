// scene.cpp — flattened-scene container, .tbscene cache I/O, .hdr loader and the
// built-in procedural scenes. Host-only code; see scene.h.
#include "scene.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include "../common/tb_math.h"

namespace tb {

// ------------------------------------------------------------------ helpers
static inline TbFloat3 sub3(TbFloat3 a, TbFloat3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline float dot3(TbFloat3 a, TbFloat3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline TbFloat3 cross3(TbFloat3 a, TbFloat3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static inline TbFloat3 normalize3(TbFloat3 a) {
    float l = 1.0f / sqrtf(dot3(a, a));
    return {a.x * l, a.y * l, a.z * l};
}

TbMaterial default_material(TbFloat3 emissive) {
    // TracerBoy.cpp:275-283
    TbMaterial m;
    memset(&m, 0, sizeof(m));
    m.IOR = 1.5f;
    m.albedoIndex = m.alphaIndex = m.normalMapIndex = m.emissiveIndex = m.specularMapIndex =
        TB_INVALID_TEXTURE;
    m.emissive = emissive;
    float avg = (float)((emissive.x + emissive.y + emissive.z) / 3.0);
    m.Flags = avg > 0.0 ? TB_LIGHT_MATERIAL_FLAG : TB_DEFAULT_MATERIAL_FLAG;
    return m;
}

uint32_t append_geometry(Scene& s, const TbFloat3* pos, const TbFloat3* nrm, const TbFloat2* uv,
                         const TbFloat3* tan, uint32_t nverts, const uint32_t* idx, uint32_t nidx,
                         uint32_t material) {
    TbGeometryRecord g;
    memset(&g, 0, sizeof(g));
    g.MaterialIndex = material;
    g.VertexFirst = (uint32_t)s.positions.size();
    g.VertexCount = nverts;
    g.IndexFirst = (uint32_t)s.indices.size();
    g.IndexCount = nidx;
    g.GeometryFlags = 1; // D3D12_RAYTRACING_GEOMETRY_FLAG_OPAQUE (USE_ANYHIT off, TracerBoy.cpp:1761)
    g.GeometryIndex = (uint32_t)s.geoms.size();
    s.positions.insert(s.positions.end(), pos, pos + nverts);
    size_t v0 = s.vertices.size();
    s.vertices.resize(v0 + nverts);
    for (uint32_t v = 0; v < nverts; v++) {
        TbVertex& o = s.vertices[v0 + v];
        o.Normal = nrm ? nrm[v] : TbFloat3{0, 1, 0}; // TracerBoy.cpp:1644
        o.Tangent = tan ? tan[v] : TbFloat3{0, 0, 1};
        o.UV = uv ? uv[v] : TbFloat2{0, 0};
    }
    s.indices.insert(s.indices.end(), idx, idx + nidx);
    if (!nrm) {
        // flat normal written into the (shared) vertices, last face wins: TracerBoy.cpp:1710-1729
        for (uint32_t i = 0; i + 2 < nidx; i += 3) {
            uint32_t a = idx[i], b = idx[i + 1], c = idx[i + 2];
            TbFloat3 e1 = sub3(pos[c], pos[a]);
            TbFloat3 e2 = sub3(pos[c], pos[b]);
            TbFloat3 n = cross3(e1, e2);
            if (dot3(n, n) <= 0.0000000001f) n = {0, 1, 0};
            else n = normalize3(n);
            s.vertices[v0 + a].Normal = n;
            s.vertices[v0 + b].Normal = n;
            s.vertices[v0 + c].Normal = n;
        }
    }
    s.geoms.push_back(g);
    return g.GeometryIndex;
}

void append_area_lights(Scene& s, const TbFloat3* pos, const TbFloat3* nrm, const uint32_t* idx,
                        uint32_t nidx, TbFloat3 emissive) {
    // TracerBoy.cpp:1531-1575
    for (uint32_t i = 0; i + 2 < nidx; i += 3) {
        TbLight l;
        memset(&l, 0, sizeof(l));
        l.LightType = TB_LIGHT_TYPE_AREA;
        l.LightColor = emissive;
        TbFloat3 p0 = pos[idx[i]], p1 = pos[idx[i + 1]], p2 = pos[idx[i + 2]];
        TbFloat3 v0 = sub3(p1, p0), v1 = sub3(p2, p0);
        float l0 = sqrtf(dot3(v0, v0)), l1 = sqrtf(dot3(v1, v1));
        float angle = acosf(dot3(v0, v1) / (l0 * l1));
        l.SurfaceArea = (float)(l0 * l1 * sinf(angle) / 2.0);
        l.P0 = p0; l.P1 = p1; l.P2 = p2;
        if (nrm) {
            l.N0 = nrm[idx[i]]; l.N1 = nrm[idx[i + 1]]; l.N2 = nrm[idx[i + 2]];
        } else {
            TbFloat3 n = normalize3(cross3(sub3(p1, p0), sub3(p2, p0)));
            l.N0 = l.N1 = l.N2 = n;
        }
        s.lights.push_back(l);
    }
}

// ------------------------------------------------------------- .tbscene I/O
struct FileHeader {
    char magic[8];
    uint32_t version;
    uint32_t flipTextureUVs;
    uint32_t numGeoms, numVerts, numIndices, numMaterials, numLights, numTextures, numImages;
    int32_t envImage;
    TbCamera camera;
    TbFloat4 envTransform[3];
    TbFloat3 envColorScale;
    uint32_t reserved[8];
};

template <class T>
static bool wr(FILE* f, const T* p, size_t n) { return n == 0 || fwrite(p, sizeof(T), n, f) == n; }
template <class T>
static bool rd(FILE* f, T* p, size_t n) { return n == 0 || fread(p, sizeof(T), n, f) == n; }

bool save_tbscene(const Scene& s, const std::string& path, std::string& err) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { err = "cannot open for write: " + path; return false; }
    FileHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, "TBSCENE1", 8);
    h.version = 1;
    h.flipTextureUVs = s.flipTextureUVs;
    h.numGeoms = (uint32_t)s.geoms.size();
    h.numVerts = (uint32_t)s.positions.size();
    h.numIndices = (uint32_t)s.indices.size();
    h.numMaterials = (uint32_t)s.materials.size();
    h.numLights = (uint32_t)s.lights.size();
    h.numTextures = (uint32_t)s.textures.size();
    h.numImages = (uint32_t)s.images.size();
    h.envImage = s.envImage;
    h.camera = s.camera;
    memcpy(h.envTransform, s.envTransform, sizeof(h.envTransform));
    h.envColorScale = s.envColorScale;
    bool ok = wr(f, &h, 1) && wr(f, s.geoms.data(), s.geoms.size()) &&
              wr(f, s.positions.data(), s.positions.size()) &&
              wr(f, s.vertices.data(), s.vertices.size()) &&
              wr(f, s.indices.data(), s.indices.size()) &&
              wr(f, s.materials.data(), s.materials.size()) &&
              wr(f, s.lights.data(), s.lights.size()) &&
              wr(f, s.textures.data(), s.textures.size());
    for (size_t i = 0; ok && i < s.materials.size(); i++) {
        char name[64];
        memset(name, 0, sizeof(name));
        if (i < s.materialNames.size()) strncpy(name, s.materialNames[i].c_str(), 63);
        ok = wr(f, name, 64);
    }
    for (size_t i = 0; ok && i < s.images.size(); i++) {
        uint32_t ih[4] = {s.images[i].width, s.images[i].height, s.images[i].format,
                          (uint32_t)s.images[i].data.size()};
        ok = wr(f, ih, 4) && wr(f, s.images[i].data.data(), s.images[i].data.size());
    }
    fclose(f);
    if (!ok) err = "short write: " + path;
    return ok;
}

// Every index the kernels dereference without a bounds check. The reference leans on D3D12's robust buffer access
// (out-of-range reads return 0); here a stale or malformed file would be an illegal-address fault that poisons the
// CUDA context of the whole process, so it is rejected on the host instead.
bool validate_material(const Scene& s, const TbMaterial& m, std::string& err) {
    const uint32_t nt = (uint32_t)s.textures.size();
    const uint32_t tex[5] = {m.albedoIndex, m.alphaIndex, m.normalMapIndex, m.emissiveIndex, m.specularMapIndex};
    for (uint32_t t : tex)
        if (t != TB_INVALID_TEXTURE && t >= nt) { err = "material references a texture that does not exist"; return false; }
    if (m.Flags & TB_MIX_MATERIAL_FLAG) { // the two mixed material ids travel in albedo.x / albedo.y (TracerBoy.cpp:452-470)
        const float n = (float)s.materials.size();
        if (!(m.albedo.x >= 0.0f && m.albedo.x < n) || !(m.albedo.y >= 0.0f && m.albedo.y < n)) {
            err = "mix material references a material that does not exist"; return false;
        }
    }
    return true;
}

bool validate_scene(const Scene& s, std::string& err) {
    if (s.vertices.size() != s.positions.size()) { err = "vertex attribute count differs from position count"; return false; }
    for (const auto& g : s.geoms) {
        if ((uint64_t)g.VertexFirst + g.VertexCount > s.positions.size() ||
            (uint64_t)g.IndexFirst + g.IndexCount > s.indices.size() || g.IndexCount % 3 ||
            g.MaterialIndex >= s.materials.size()) { err = "geometry range or material index out of range"; return false; }
        for (uint32_t i = 0; i < g.IndexCount; i++)
            if (s.indices[g.IndexFirst + i] >= g.VertexCount) { err = "vertex index out of range"; return false; }
    }
    for (const auto& m : s.materials)
        if (!validate_material(s, m, err)) return false;
    for (const auto& t : s.textures) {
        if (t.TextureType == TB_IMAGE_TEXTURE_TYPE && t.DescriptorHeapIndex >= s.images.size()) { err = "texture references an image that does not exist"; return false; }
        if (t.TextureType == TB_SCALE_TEXTURE_TYPE && (t.TextureIndex1 >= s.textures.size() || t.TextureIndex2 >= s.textures.size())) {
            err = "scale texture references a texture that does not exist"; return false;
        }
    }
    if (s.envImage < -1 || s.envImage >= (int64_t)s.images.size()) { err = "environment image index out of range"; return false; }
    for (const auto& im : s.images) {
        const uint64_t bpp = im.format == 0 ? 16 : 4;
        if (im.format > 2 || im.width == 0 || im.height == 0 || (uint64_t)im.width * im.height * bpp != im.data.size()) {
            err = "image size does not match its pixel data"; return false;
        }
    }
    return true;
}

bool load_tbscene(Scene& s, const std::string& path, std::string& err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open: " + path; return false; }
    FileHeader h;
    if (!rd(f, &h, 1) || memcmp(h.magic, "TBSCENE1", 8) != 0 || h.version != 1) {
        fclose(f);
        err = "not a .tbscene v1 file: " + path;
        return false;
    }
    // the counts drive the allocations below: bound them by what the file can hold before resizing anything
    fseek(f, 0, SEEK_END);
    const uint64_t fileBytes = (uint64_t)ftell(f);
    fseek(f, (long)sizeof(FileHeader), SEEK_SET);
    const uint64_t declared = sizeof(FileHeader) + (uint64_t)h.numGeoms * sizeof(TbGeometryRecord) +
                              (uint64_t)h.numVerts * (sizeof(TbFloat3) + sizeof(TbVertex)) + (uint64_t)h.numIndices * 4 +
                              (uint64_t)h.numMaterials * (sizeof(TbMaterial) + 64) + (uint64_t)h.numLights * sizeof(TbLight) +
                              (uint64_t)h.numTextures * sizeof(TbTextureData) + (uint64_t)h.numImages * 16;
    if (declared > fileBytes) { fclose(f); err = "corrupt .tbscene (header counts exceed the file size): " + path; return false; }
    s.clear();
    s.flipTextureUVs = h.flipTextureUVs;
    s.envImage = h.envImage;
    s.camera = h.camera;
    memcpy(s.envTransform, h.envTransform, sizeof(h.envTransform));
    s.envColorScale = h.envColorScale;
    s.geoms.resize(h.numGeoms);
    s.positions.resize(h.numVerts);
    s.vertices.resize(h.numVerts);
    s.indices.resize(h.numIndices);
    s.materials.resize(h.numMaterials);
    s.lights.resize(h.numLights);
    s.textures.resize(h.numTextures);
    s.images.resize(h.numImages);
    bool ok = rd(f, s.geoms.data(), s.geoms.size()) && rd(f, s.positions.data(), s.positions.size()) &&
              rd(f, s.vertices.data(), s.vertices.size()) && rd(f, s.indices.data(), s.indices.size()) &&
              rd(f, s.materials.data(), s.materials.size()) && rd(f, s.lights.data(), s.lights.size()) &&
              rd(f, s.textures.data(), s.textures.size());
    for (size_t i = 0; ok && i < s.materials.size(); i++) {
        char name[64];
        ok = rd(f, name, 64);
        name[63] = 0;
        s.materialNames.push_back(name);
    }
    for (size_t i = 0; ok && i < s.images.size(); i++) {
        uint32_t ih[4];
        ok = rd(f, ih, 4);
        if (!ok) break;
        if ((uint64_t)ih[3] > fileBytes) { ok = false; break; }
        s.images[i].width = ih[0];
        s.images[i].height = ih[1];
        s.images[i].format = ih[2];
        s.images[i].data.resize(ih[3]);
        ok = rd(f, s.images[i].data.data(), ih[3]);
    }
    fclose(f);
    if (!ok) { err = "short read: " + path; return false; }
    std::string why;
    if (!validate_scene(s, why)) { err = "corrupt .tbscene (" + why + "): " + path; return false; }
    return true;
}

// ------------------------------------------------------------------- .hdr
bool load_hdr(const std::string& path, Image& img, std::string& err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open: " + path; return false; }
    char line[512];
    bool gotFormat = false;
    int w = 0, h = 0;
    if (!fgets(line, sizeof(line), f) || strncmp(line, "#?", 2) != 0) {
        fclose(f); err = "not a Radiance .hdr: " + path; return false;
    }
    while (fgets(line, sizeof(line), f)) {
        if (line[0] == '\n' || line[0] == '\r') break;
        if (strncmp(line, "FORMAT=32-bit_rle_rgbe", 22) == 0) gotFormat = true;
    }
    if (!fgets(line, sizeof(line), f) || sscanf(line, "-Y %d +X %d", &h, &w) != 2 || !gotFormat ||
        w <= 0 || h <= 0) {
        fclose(f); err = "unsupported .hdr header: " + path; return false;
    }
    img.width = w; img.height = h; img.format = 0;
    img.data.resize((size_t)w * h * 16);
    float* out = (float*)img.data.data();
    std::vector<uint8_t> scan((size_t)w * 4);
    for (int y = 0; y < h; y++) {
        uint8_t hd[4];
        if (fread(hd, 1, 4, f) != 4) { fclose(f); err = "truncated .hdr"; return false; }
        if (hd[0] == 2 && hd[1] == 2 && !(hd[2] & 0x80) && ((hd[2] << 8) | hd[3]) == w && w >= 8 && w < 32768) {
            for (int c = 0; c < 4; c++) {
                int x = 0;
                while (x < w) {
                    int n = fgetc(f);
                    if (n == EOF) { fclose(f); err = "truncated .hdr"; return false; }
                    if (n > 128) {
                        n -= 128;
                        int v = fgetc(f);
                        while (n-- && x < w) scan[(size_t)x++ * 4 + c] = (uint8_t)v;
                    } else {
                        while (n-- && x < w) scan[(size_t)x++ * 4 + c] = (uint8_t)fgetc(f);
                    }
                }
            }
        } else { // flat scanline
            memcpy(scan.data(), hd, 4);
            if (fread(scan.data() + 4, 1, (size_t)(w - 1) * 4, f) != (size_t)(w - 1) * 4) {
                fclose(f); err = "truncated .hdr"; return false;
            }
        }
        for (int x = 0; x < w; x++) {
            const uint8_t* p = &scan[(size_t)x * 4];
            float* o = out + ((size_t)y * w + x) * 4;
            if (p[3]) {
                float sc = ldexpf(1.0f, (int)p[3] - (128 + 8));
                o[0] = p[0] * sc; o[1] = p[1] * sc; o[2] = p[2] * sc;
            } else o[0] = o[1] = o[2] = 0.0f;
            o[3] = 1.0f;
        }
    }
    fclose(f);
    return true;
}

// ------------------------------------------------------- procedural scenes
static inline uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }
static inline float unit(uint32_t& s) { return (float)(lcg(s) >> 8) * (1.0f / 16777216.0f); }

static std::map<std::string, std::string> parse_query(const std::string& q) {
    std::map<std::string, std::string> m;
    size_t i = 0;
    while (i < q.size()) {
        size_t a = q.find('&', i);
        if (a == std::string::npos) a = q.size();
        std::string kv = q.substr(i, a - i);
        size_t e = kv.find('=');
        if (e != std::string::npos) m[kv.substr(0, e)] = kv.substr(e + 1);
        i = a + 1;
    }
    return m;
}

static void add_quad(Scene& s, TbFloat3 a, TbFloat3 b, TbFloat3 c, TbFloat3 d, uint32_t mat,
                     bool light, TbFloat3 emissive) {
    TbFloat3 p[4] = {a, b, c, d};
    TbFloat3 n = normalize3(cross3(sub3(b, a), sub3(c, a)));
    TbFloat3 nn[4] = {n, n, n, n};
    TbFloat2 uv[4] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}};
    uint32_t idx[6] = {0, 1, 2, 0, 2, 3};
    append_geometry(s, p, nn, uv, nullptr, 4, idx, 6, mat);
    if (light) append_area_lights(s, p, nn, idx, 6, emissive);
}

// A displaced lat-long sphere with exactly 2*rings*segs - 2*segs... triangles; we build
// (rings x segs) quads (poles are degenerate-free: ring 0 / ring R are single vertices).
static void add_blob(Scene& s, TbFloat3 center, float radius, uint32_t rings, uint32_t segs,
                     uint32_t seed, uint32_t mat, float displacement = 1.0f) {
    std::vector<TbFloat3> pos, nrm;
    std::vector<TbFloat2> uv;
    std::vector<uint32_t> idx;
    const float PI_F = 3.14159265358979f;
    // low-frequency displacement: 3 random lobes
    uint32_t rs = seed * 2654435761u + 12345u;
    float ax[3], ph[3], am[3];
    for (int k = 0; k < 3; k++) { ax[k] = 2.0f + floorf(unit(rs) * 5.0f); ph[k] = unit(rs) * 6.28f; am[k] = displacement * (0.04f + 0.06f * unit(rs)); }
    for (uint32_t r = 0; r <= rings; r++) {
        float th = PI_F * (float)r / (float)rings;
        float st = tbm::sin_(th), ct = tbm::cos_(th);
        for (uint32_t g = 0; g <= segs; g++) {
            float phi = 2.0f * PI_F * (float)(g % segs) / (float)segs;
            float sp = tbm::sin_(phi), cp = tbm::cos_(phi);
            float d = 1.0f + am[0] * tbm::sin_(ax[0] * th + ph[0]) + am[1] * tbm::sin_(ax[1] * phi + ph[1]) * st +
                      am[2] * tbm::sin_(ax[2] * (th + phi) + ph[2]) * st;
            TbFloat3 dir = {st * cp, ct, st * sp};
            pos.push_back({center.x + radius * d * dir.x, center.y + radius * d * dir.y, center.z + radius * d * dir.z});
            nrm.push_back(dir);
            uv.push_back({(float)g / (float)segs, (float)r / (float)rings});
        }
    }
    uint32_t stride = segs + 1;
    for (uint32_t r = 0; r < rings; r++)
        for (uint32_t g = 0; g < segs; g++) {
            uint32_t a = r * stride + g, b = a + 1, c = a + stride, d = c + 1;
            if (r != 0) { idx.push_back(a); idx.push_back(b); idx.push_back(c); }
            if (r != rings - 1) { idx.push_back(b); idx.push_back(d); idx.push_back(c); }
        }
    append_geometry(s, pos.data(), nrm.data(), uv.data(), nullptr, (uint32_t)pos.size(), idx.data(),
                    (uint32_t)idx.size(), mat);
}

static uint32_t add_mat(Scene& s, const char* name, TbMaterial m) {
    m.Flags |= TB_NO_ALPHA_MATERIAL_FLAG;
    s.materials.push_back(m);
    s.materialNames.push_back(name);
    return (uint32_t)s.materials.size() - 1;
}

bool make_synthetic(Scene& s, const std::string& spec, std::string& err) {
    // spec = "synthetic:<name>?k=v&k=v"
    std::string body = spec.substr(strlen("synthetic:"));
    std::string name = body, query;
    size_t qm = body.find('?');
    if (qm != std::string::npos) { name = body.substr(0, qm); query = body.substr(qm + 1); }
    auto q = parse_query(query);
    auto geti = [&](const char* k, long def) { return q.count(k) ? atol(q[k].c_str()) : def; };
    s.clear();
    if (name == "blobs") {
        // `copies` displaced spheres of ~`tris` triangles on a jittered 3-D
        // grid in [-100,100]^3, 10 materials round-robin, one 2-triangle area light,
        // constant sky. copies=1 gives the single high-poly stand-in for the dragon body.
        long copies = geti("copies", 64), tris = geti("tris", 1000), seed = geti("seed", 1);
        if (copies < 1 || tris < 8 || copies * tris > 200000000L) { err = "blobs: bad copies/tris"; return false; }
        // rings*segs*2 - 2*segs ~= tris, segs = 1.25*rings
        uint32_t rings = 2, segs = 3;
        while (2ul * rings * segs - 2ul * segs < (unsigned long)tris) { rings++; segs = rings + rings / 4; if (segs < 3) segs = 3; }
        TbMaterial m;
        uint32_t mats[10];
        const float cols[5][3] = {{0.7f, 0.2f, 0.2f}, {0.2f, 0.6f, 0.25f}, {0.25f, 0.3f, 0.7f}, {0.7f, 0.65f, 0.3f}, {0.6f, 0.6f, 0.6f}};
        for (int i = 0; i < 10; i++) {
            m = default_material({0, 0, 0});
            const float* c = cols[i % 5];
            m.albedo = {c[0], c[1], c[2]};
            char nm[32];
            switch (i % 4) {
            case 0: m.Flags |= TB_NO_SPECULAR_MATERIAL_FLAG; snprintf(nm, 32, "matte%d", i); break; // matte (:435-447)
            case 1: { float ks = 0.04f; m.IOR = (sqrtf(ks) + 1.0f) / (1.0f - sqrtf(ks)); m.SpecularCoef = ks; m.roughness = 0.1f + 0.05f * i; snprintf(nm, 32, "substrate%d", i); break; } // substrate (:408-421)
            case 2: m.albedo = {1, 1, 1}; m.IOR = 0.75f; m.roughness = 0.05f * (i - 1); m.Flags |= TB_METALLIC_MATERIAL_FLAG; snprintf(nm, 32, "metal%d", i); break; // metal (:397-406)
            default: m.albedo = {0, 0, 0}; m.IOR = 1.5f; m.Flags |= TB_SUBSURFACE_SCATTER_MATERIAL_FLAG; snprintf(nm, 32, "glass%d", i); break; // glass (:423-431)
            }
            mats[i] = add_mat(s, nm, m);
        }
        m = default_material({0, 0, 0});
        m.albedo = {0.5f, 0.5f, 0.5f};
        m.Flags |= TB_NO_SPECULAR_MATERIAL_FLAG;
        uint32_t floorMat = add_mat(s, "floor", m);
        TbFloat3 Le = {40.0f, 36.0f, 30.0f};
        m = default_material(Le);
        m.Flags |= TB_NO_SPECULAR_MATERIAL_FLAG;
        uint32_t lightMat = add_mat(s, "light", m);

        uint32_t side = 1;
        while ((long)side * side * side < copies) side++;
        float cell = 200.0f / (float)side;
        float radius = 0.36f * cell;
        uint32_t rs = (uint32_t)seed;
        long made = 0;
        for (uint32_t z = 0; z < side && made < copies; z++)
            for (uint32_t y = 0; y < side && made < copies; y++)
                for (uint32_t x = 0; x < side && made < copies; x++, made++) {
                    TbFloat3 c = {-100.0f + cell * ((float)x + 0.5f + 0.2f * (unit(rs) - 0.5f)),
                                  -100.0f + cell * ((float)y + 0.5f + 0.2f * (unit(rs) - 0.5f)),
                                  -100.0f + cell * ((float)z + 0.5f + 0.2f * (unit(rs) - 0.5f))};
                    add_blob(s, c, radius, rings, segs, (uint32_t)(seed * 7919 + made), mats[made % 10]);
                }
        add_quad(s, {-160, -101, -160}, {-160, -101, 160}, {160, -101, 160}, {160, -101, -160}, floorMat, false, {0, 0, 0});
        add_quad(s, {-40, 140, -40}, {40, 140, -40}, {40, 140, 40}, {-40, 140, 40}, lightMat, true, Le);
        // constant sky
        Image sky;
        sky.width = sky.height = 1; sky.format = 0;
        float px[4] = {0.35f, 0.45f, 0.6f, 1.0f};
        sky.data.assign((uint8_t*)px, (uint8_t*)px + 16);
        s.images.push_back(sky);
        s.envImage = 0;
        // camera: looking at the cube from +z, +y (fixed)
        TbFloat3 eye = {150.0f, 90.0f, 330.0f}, target = {0, -10.0f, 0};
        TbFloat3 view = normalize3(sub3(target, eye));
        TbFloat3 right = normalize3(cross3({0, 1, 0}, view));
        TbFloat3 up = cross3(view, right);
        s.camera.LensHeight = 2.0f;
        s.camera.FocalDistance = 1.0f / 0.36397f; // fov 40 deg: 1/tan(20deg)
        s.camera.Position = {eye.x + (s.camera.FocalDistance + 0.01f) * view.x, eye.y + (s.camera.FocalDistance + 0.01f) * view.y, eye.z + (s.camera.FocalDistance + 0.01f) * view.z};
        s.camera.LookAt = {s.camera.Position.x + view.x, s.camera.Position.y + view.y, s.camera.Position.z + view.z};
        s.camera.Right = right;
        s.camera.Up = up;
        return true;
    }
    if (name == "showcase") {
        // Every material / texture / light code path of the tracer in one small scene (parity coverage):
        // matte + checker, substrate + scale(checker, image) texture, plastic-like dielectric + RGBA8 image with
        // gamma flag, metal with specular map, mirror, glass, single-sided uber glass, SSS with artist albedo,
        // mix(metal, matte), emissive texture, hair flag; one area light, one directional light, HDR-ish sky.
        long tris = geti("tris", 400), seed = geti("seed", 1);
        uint32_t rings = 2, segs = 3;
        while (2ul * rings * segs - 2ul * segs < (unsigned long)tris) { rings++; segs = rings + rings / 4; if (segs < 3) segs = 3; }
        // images: 0 = sky (float4 4x2), 1 = RGBA8 8x8 pattern, 2 = float4 4x4 pattern
        {
            Image sky; sky.width = 4; sky.height = 2; sky.format = 0;
            float px[8][4] = {{0.3f, 0.4f, 0.7f, 1}, {0.5f, 0.5f, 0.6f, 1}, {2.5f, 2.2f, 1.8f, 1}, {0.4f, 0.45f, 0.6f, 1},
                              {0.2f, 0.2f, 0.25f, 1}, {0.25f, 0.22f, 0.2f, 1}, {0.3f, 0.25f, 0.2f, 1}, {0.2f, 0.2f, 0.22f, 1}};
            sky.data.assign((uint8_t*)px, (uint8_t*)px + sizeof(px));
            s.images.push_back(sky);
            Image a; a.width = a.height = 8; a.format = 1; a.data.resize(8 * 8 * 4);
            uint32_t rs = (uint32_t)seed * 977u + 5u;
            for (size_t i = 0; i < a.data.size(); i++) a.data[i] = (uint8_t)(64 + (lcg(rs) >> 25));
            s.images.push_back(a);
            Image b; b.width = b.height = 4; b.format = 0; b.data.resize(4 * 4 * 16);
            float* bp = (float*)b.data.data();
            for (int i = 0; i < 64; i++) bp[i] = 0.1f + 0.8f * unit(rs);
            s.images.push_back(b);
        }
        s.envImage = 0;
        s.envTransform[0] = {0.8f, 0.0f, 0.6f, 0}; s.envTransform[1] = {0.0f, 1.0f, 0.0f, 0}; s.envTransform[2] = {-0.6f, 0.0f, 0.8f, 0};
        s.envColorScale = {1.0f, 0.9f, 0.8f};
        s.flipTextureUVs = 1;
        auto checker = [&](float us, float vs, TbFloat3 c1, TbFloat3 c2) {
            TbTextureData t; memset(&t, 0, sizeof(t));
            t.TextureType = TB_CHECKER_TEXTURE_TYPE; t.UScale = us; t.VScale = vs; t.CheckerColor1 = c1; t.CheckerColor2 = c2;
            s.textures.push_back(t); return (uint32_t)s.textures.size() - 1;
        };
        auto image = [&](uint32_t img, uint32_t flags) {
            TbTextureData t; memset(&t, 0, sizeof(t));
            t.TextureType = TB_IMAGE_TEXTURE_TYPE; t.DescriptorHeapIndex = img; t.TextureFlags = flags;
            s.textures.push_back(t); return (uint32_t)s.textures.size() - 1;
        };
        uint32_t texChecker = checker(6.0f, 3.0f, {0.8f, 0.2f, 0.2f}, {0.2f, 0.2f, 0.8f});
        uint32_t texImg8 = image(1, TB_NEEDS_GAMMA_CORRECTION_TEXTURE_FLAG);
        uint32_t texImgF = image(2, 0);
        uint32_t texScale;
        {
            TbTextureData t; memset(&t, 0, sizeof(t));
            t.TextureType = TB_SCALE_TEXTURE_TYPE; t.TextureIndex1 = texChecker; t.TextureIndex2 = texImgF;
            t.ScaleColor1 = {0.5f, 0.6f, 0.7f}; t.ScaleColor2 = {0.4f, 0.3f, 0.2f};
            s.textures.push_back(t); texScale = (uint32_t)s.textures.size() - 1;
        }
        uint32_t texSpec = checker(4.0f, 4.0f, {0.0f, 0.35f, 1.0f}, {0.0f, 0.1f, 0.0f}); // g = roughness, b > 0.5 => metallic
        std::vector<uint32_t> mats;
        TbMaterial m;
        m = default_material({0, 0, 0}); m.albedo = {0.5f, 0.5f, 0.5f}; m.albedoIndex = texChecker; m.Flags |= TB_NO_SPECULAR_MATERIAL_FLAG; mats.push_back(add_mat(s, "matte_checker", m));
        m = default_material({0, 0, 0}); m.albedo = {0.9f, 0.9f, 0.9f}; m.albedoIndex = texScale; m.IOR = 1.5f; m.SpecularCoef = 0.04f; m.roughness = 0.2f; mats.push_back(add_mat(s, "substrate_scale", m));
        m = default_material({0, 0, 0}); m.albedo = {0.5f, 0.5f, 0.5f}; m.albedoIndex = texImg8; m.IOR = 1.46f; m.SpecularCoef = 0.04f; m.roughness = 0.03f; mats.push_back(add_mat(s, "plastic_image", m));
        m = default_material({0, 0, 0}); m.albedo = {0.9f, 0.8f, 0.6f}; m.specularMapIndex = texSpec; m.roughness = 0.3f; m.IOR = 0.8f; mats.push_back(add_mat(s, "specmap", m));
        m = default_material({0, 0, 0}); m.albedo = {0.9f, 0.9f, 0.9f}; m.SpecularCoef = 1.0f; m.roughness = 0.0f; m.Flags |= TB_METALLIC_MATERIAL_FLAG; mats.push_back(add_mat(s, "mirror", m));
        m = default_material({0, 0, 0}); m.albedo = {0, 0, 0}; m.IOR = 1.5f; m.roughness = 0.0f; m.Flags |= TB_SUBSURFACE_SCATTER_MATERIAL_FLAG; mats.push_back(add_mat(s, "glass", m));
        m = default_material({0, 0, 0}); m.albedo = {0, 0, 0}; m.IOR = 1.3f; m.roughness = 0.3f; m.Flags |= TB_SUBSURFACE_SCATTER_MATERIAL_FLAG; mats.push_back(add_mat(s, "rough_glass", m));
        m = default_material({0, 0, 0}); m.albedo = {0.0f, 0.0f, 0.0f}; m.IOR = 1.5f; m.absorption = {0.2f, 0.05f, 0.4f}; m.roughness = 0.1f; m.Flags |= TB_SUBSURFACE_SCATTER_MATERIAL_FLAG | TB_SINGLE_SIDED_MATERIAL_FLAG; mats.push_back(add_mat(s, "uber_thin", m));
        m = default_material({0, 0, 0}); m.albedo = {0.8f, 0.5f, 0.3f}; m.scattering = {2.0f, 1.0f, 0.5f}; m.IOR = 1.4f; m.roughness = 0.4f; m.Flags |= TB_SUBSURFACE_SCATTER_MATERIAL_FLAG | TB_NO_SPECULAR_MATERIAL_FLAG; mats.push_back(add_mat(s, "sss_albedo", m));
        {   // mix(metal, matte): sub-materials first, as MaterialTracker adds them (TracerBoy.cpp:365-373)
            TbMaterial a = default_material({0, 0, 0}); a.albedo = {1, 1, 1}; a.IOR = 0.9f; a.roughness = 0.15f; a.Flags |= TB_METALLIC_MATERIAL_FLAG;
            uint32_t i0 = add_mat(s, "mix_metal", a);
            TbMaterial b = default_material({0, 0, 0}); b.albedo = {0.1f, 0.3f, 0.6f}; b.Flags |= TB_NO_SPECULAR_MATERIAL_FLAG;
            uint32_t i1 = add_mat(s, "mix_matte", b);
            m = default_material({0, 0, 0}); m.Flags = TB_MIX_MATERIAL_FLAG; m.albedo = {(float)i0, (float)i1, 0.35f};
            mats.push_back(add_mat(s, "mix", m));
        }
        m = default_material({0, 0, 0}); m.albedo = {0.3f, 0.3f, 0.3f}; m.emissiveIndex = texImgF; m.Flags |= TB_NO_SPECULAR_MATERIAL_FLAG; mats.push_back(add_mat(s, "emissive_tex", m));
        m = default_material({0, 0, 0}); m.albedo = {0.4f, 0.25f, 0.1f}; m.roughness = 0.25f; m.Flags |= TB_HAIR_MATERIAL_FLAG; mats.push_back(add_mat(s, "hair", m));
        m = default_material({0, 0, 0}); m.albedo = {0.6f, 0.6f, 0.6f}; m.normalMapIndex = texImgF; m.IOR = 1.5f; m.SpecularCoef = 0.05f; m.roughness = 0.3f; mats.push_back(add_mat(s, "normalmapped", m));
        uint32_t side = 4;
        for (size_t k = 0; k < mats.size(); k++) {
            TbFloat3 c = {-45.0f + 30.0f * (float)(k % side), -20.0f + 28.0f * (float)(k / side), 10.0f * (float)((k * 7) % 3)};
            add_blob(s, c, 11.0f, rings, segs, (uint32_t)(seed * 131 + k), mats[k]);
        }
        m = default_material({0, 0, 0}); m.albedo = {0.5f, 0.5f, 0.5f}; m.albedoIndex = texChecker; m.IOR = 1.5f; m.SpecularCoef = 0.04f; m.roughness = 0.15f;
        uint32_t floorMat = add_mat(s, "floor", m);
        TbFloat3 Le = {30.0f, 28.0f, 24.0f};
        m = default_material(Le); m.Flags |= TB_NO_SPECULAR_MATERIAL_FLAG;
        uint32_t lightMat = add_mat(s, "light", m);
        add_quad(s, {-120, -34, -120}, {-120, -34, 120}, {120, -34, 120}, {120, -34, -120}, floorMat, false, {0, 0, 0});
        add_quad(s, {-30, 110, -30}, {30, 110, -30}, {30, 110, 30}, {-30, 110, 30}, lightMat, true, Le);
        {   // directional light (TracerBoy.cpp:1908-1916)
            TbLight l; memset(&l, 0, sizeof(l));
            l.LightType = TB_LIGHT_TYPE_DIRECTIONAL; l.LightColor = {1.5f, 1.4f, 1.2f};
            TbFloat3 d = normalize3({-0.3f, -0.8f, 0.5f}); l.Direction = d;
            s.lights.push_back(l);
        }
        TbFloat3 eye = {20.0f, 40.0f, -230.0f}, target = {0, 20.0f, 0};
        TbFloat3 view = normalize3(sub3(target, eye));
        TbFloat3 right = normalize3(cross3({0, 1, 0}, view));
        TbFloat3 up = cross3(view, right);
        s.camera.LensHeight = 2.0f;
        s.camera.FocalDistance = 1.0f / 0.36397f;
        float fd = s.camera.FocalDistance + 0.01f;
        s.camera.Position = {eye.x + fd * view.x, eye.y + fd * view.y, eye.z + fd * view.z};
        s.camera.LookAt = {s.camera.Position.x + view.x, s.camera.Position.y + view.y, s.camera.Position.z + view.z};
        s.camera.Right = right;
        s.camera.Up = up;
        return true;
    }
    if (name == "furnace") {
        // One convex matte sphere of albedo a under a constant white sky, no lights. Known answer
        // with MaxBounces = 2: every pixel that hits the sphere resolves to a (throughput =
        // albedo * cos/pi / (cos/pi), second ray escapes to the sky), every other pixel to 1.
        TbMaterial m = default_material({0, 0, 0});
        float a = q.count("albedo") ? (float)atof(q["albedo"].c_str()) : 0.5f;
        m.albedo = {a, a, a};
        m.Flags |= TB_NO_SPECULAR_MATERIAL_FLAG;
        uint32_t mat = add_mat(s, "furnace", m);
        add_blob(s, {0, 0, 0}, 5.0f, 48, 64, 0, mat, 0.0f); // plain convex sphere
        Image sky;
        sky.width = sky.height = 1; sky.format = 0;
        float px[4] = {1, 1, 1, 1};
        sky.data.assign((uint8_t*)px, (uint8_t*)px + 16);
        s.images.push_back(sky);
        s.envImage = 0;
        s.camera.LensHeight = 2.0f;
        s.camera.FocalDistance = 2.0f;
        s.camera.Position = {0, 0, -20.0f + 2.01f};
        s.camera.LookAt = {0, 0, -20.0f + 3.01f};
        s.camera.Right = {1, 0, 0};
        s.camera.Up = {0, 1, 0};
        return true;
    }
    err = "unknown synthetic scene: " + name;
    return false;
}

} // namespace tb
