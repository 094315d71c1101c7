// scene.h — flattened scene: the output of TracerBoy::LoadScene steps 2-4
// (TracerBoy.cpp:1243-1272 camera, 1356-1835 shapes/lights/records, 1861-1944
// materials/textures/env) as plain host arrays, plus its on-disk cache (.tbscene).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "tracerboy_b200.h"

namespace tb {

struct Image {
    uint32_t width = 0, height = 0;
    uint32_t format = 0; // 0 = float4, 1 = unorm8 x4, 2 = unorm8 x4 sRGB (R8G8B8A8_UNORM_SRGB: the sampler linearises every texel before filtering)
    bool unorm16 = false; // import-time only: a float4 image that came from a 16-bit UNORM file ("normalized" for the gamma flag)
    std::vector<uint8_t> data;
};

struct Scene {
    // geometry (one global BLAS: all top-level shapes, transforms baked; TracerBoy.cpp:1361-1366)
    std::vector<TbGeometryRecord> geoms;
    std::vector<TbFloat3> positions; // pooled
    std::vector<TbVertex> vertices;  // pooled, same indexing as positions
    std::vector<uint32_t> indices;   // pooled, geometry-local vertex ids
    std::vector<TbMaterial> materials;
    std::vector<std::string> materialNames;
    std::vector<TbLight> lights;
    std::vector<TbTextureData> textures;
    std::vector<Image> images;
    TbCamera camera{};
    uint32_t flipTextureUVs = 0;       // m_flipTextureUVs (TracerBoy.cpp:1208)
    int32_t envImage = -1;             // index into images, -1 = 1x1 black (TracerBoy.cpp:1918-1934)
    TbFloat4 envTransform[3] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}}; // vx,vy,vz columns (:3378-3380)
    TbFloat3 envColorScale{1, 1, 1};

    uint32_t numTriangles() const { return (uint32_t)(indices.size() / 3); }
    void clear() { *this = Scene(); }
};

// .tbscene cache (the role of the reference's .pbf cache, TracerBoy.cpp:1200-1223).
// Layout is documented in DESIGN.md; little-endian, 8-byte magic "TBSCENE1".
bool save_tbscene(const Scene& s, const std::string& path, std::string& err);
bool load_tbscene(Scene& s, const std::string& path, std::string& err);
// Structural validation of everything the kernels index without a bounds check (geometry ranges, vertex indices,
// material -> texture, texture -> image / texture, mix-material ids, environment image, image byte sizes).
bool validate_scene(const Scene& s, std::string& err);
bool validate_material(const Scene& s, const TbMaterial& m, std::string& err);

// Radiance .hdr (RGBE) -> float4 image, as DirectXTex LoadFromHDRFile gives the reference.
bool load_hdr(const std::string& path, Image& img, std::string& err);
// Any supported texture file by extension: .hdr, .png, .tga (image_decode.cpp; the formats follow DirectXTex's loaders).
bool load_image_file(const std::string& path, Image& img, bool* hasAlpha, std::string& err);

// Built-in procedural scenes: "synthetic:<name>?key=value&..." (SURVEY §8d C5 and the
// dragon / vw-van stand-ins). Integer-only generator, no libm, deterministic.
bool make_synthetic(Scene& s, const std::string& spec, std::string& err);

// Append one mesh as a new geometry. If normals == nullptr, flat per-face normals are
// written into shared vertices exactly as TracerBoy.cpp:1710-1729 does.
uint32_t append_geometry(Scene& s, const TbFloat3* pos, const TbFloat3* nrm, const TbFloat2* uv,
                         const TbFloat3* tan, uint32_t nverts, const uint32_t* idx, uint32_t nidx,
                         uint32_t material);

// TracerBoy.cpp:1527-1576: one area Light per triangle of an emissive mesh.
void append_area_lights(Scene& s, const TbFloat3* pos, const TbFloat3* nrm, const uint32_t* idx,
                        uint32_t nidx, TbFloat3 emissive);

TbMaterial default_material(TbFloat3 emissive); // CreateMaterial prologue, TracerBoy.cpp:275-283

// image_io.cpp
bool save_png_rgba8(const std::string& path, const uint8_t* rgba, uint32_t w, uint32_t h, std::string& err);
bool save_exr_f32(const std::string& path, const float* px, uint32_t w, uint32_t h, int channels, std::string& err);
bool save_pfm_rgb(const std::string& path, const float* px, uint32_t w, uint32_t h, int channels, std::string& err);

} // namespace tb
