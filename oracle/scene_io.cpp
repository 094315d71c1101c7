// scene_io.cpp — CPU ORACLE (test infrastructure): independent reader of the .tbscene
// flattened-scene cache (layout documented in DESIGN.md §"tbscene").
#include <cstdio>
#include <cstring>
#include "oracle.h"

namespace oracle {

namespace {
struct FileHeader {
    char magic[8];
    uint32_t version, flipTextureUVs;
    uint32_t numGeoms, numVerts, numIndices, numMaterials, numLights, numTextures, numImages;
    int32_t envImage;
    TbCamera camera;
    TbFloat4 envTransform[3];
    TbFloat3 envColorScale;
    uint32_t reserved[8];
};
template <class T> bool rd(FILE* f, T* p, size_t n) { return n == 0 || fread(p, sizeof(T), n, f) == n; }
} // namespace

bool load_tbscene(Scene& s, const std::string& path, std::string& err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open " + path; return false; }
    FileHeader h;
    if (!rd(f, &h, 1) || memcmp(h.magic, "TBSCENE1", 8) != 0 || h.version != 1) { fclose(f); err = "bad .tbscene header"; return false; }
    s = Scene();
    s.flipTextureUVs = h.flipTextureUVs;
    s.envImage = h.envImage;
    s.camera = h.camera;
    memcpy(s.envTransform, h.envTransform, sizeof(h.envTransform));
    s.envColorScale = h.envColorScale;
    s.geoms.resize(h.numGeoms); s.positions.resize(h.numVerts); s.vertices.resize(h.numVerts);
    s.indices.resize(h.numIndices); s.materials.resize(h.numMaterials); s.lights.resize(h.numLights);
    s.textures.resize(h.numTextures); s.images.resize(h.numImages);
    bool ok = rd(f, s.geoms.data(), s.geoms.size()) && rd(f, s.positions.data(), s.positions.size()) &&
              rd(f, s.vertices.data(), s.vertices.size()) && rd(f, s.indices.data(), s.indices.size()) &&
              rd(f, s.materials.data(), s.materials.size()) && rd(f, s.lights.data(), s.lights.size()) &&
              rd(f, s.textures.data(), s.textures.size());
    if (ok && h.numMaterials) ok = fseek(f, 64L * h.numMaterials, SEEK_CUR) == 0; // material names
    for (size_t i = 0; ok && i < s.images.size(); i++) {
        uint32_t ih[4];
        ok = rd(f, ih, 4);
        if (!ok) break;
        s.images[i].width = ih[0]; s.images[i].height = ih[1]; s.images[i].format = ih[2];
        s.images[i].data.resize(ih[3]);
        ok = rd(f, s.images[i].data.data(), ih[3]);
    }
    fclose(f);
    if (!ok) err = "short read " + path;
    return ok;
}

} // namespace oracle
