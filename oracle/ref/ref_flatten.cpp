// ref_flatten.cpp — CPU ORACLE (test infrastructure): the scene flatten's material and light rules compiled from the mount:
// ConvertFloat3, ChannelAverage, ConvertSpecularToIOR, GetAreaLightColor, reciprocol and the whole CreateMaterial of
// TracerBoy/TracerBoy.cpp (:92-95, 117-125, 253-271, 273-505), struct MaterialTracker of TracerBoy.h (:130-156), the
// per-triangle area-light block of LoadScene (:1531-1576) and struct Light / struct Material + the flag constants of
// SharedShaderStructs.h — pre-passed into oracle/_ref/flatten_gen.inc by prepass.run_flatten. Linked against the
// reference's vendored pbrt-parser (third-party, compiled in place for the importer: build/pbrt/*.o).
// Restated here: the walk over world->shapes with the six lines of LoadScene that call CreateMaterial (:1578-1592), and a
// TextureAllocator stand-in that hands out indices in the order TextureAllocator::CreateTexture does (a scale texture's
// children first, then itself; TracerBoy.cpp:177-251) and reports "no alpha".
// tests/test_cpu_host.py compares the product importer's .tbscene materials / lights / per-shape material indices with this.
#include <climits>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>
#include "pbrtParser/Scene.h"

typedef unsigned int UINT;
typedef unsigned int uint;
struct float2 { float x, y; };
struct float3 { float x, y, z; };
#define VERIFY(x) ((void)0)
#define HANDLE_FAILURE() ((void)0)

struct TextureRecord { int type; int gamma; };
struct TextureAllocator {
    std::vector<TextureRecord> created;
    UINT CreateTexture(pbrt::Texture::SP& tex, bool bGammaCorrect = false, bool* bHasAlpha = nullptr) {
        if (!tex) return UINT_MAX;
        if (bHasAlpha) *bHasAlpha = false;
        auto scl = std::dynamic_pointer_cast<pbrt::ScaleTexture>(tex);
        if (scl) { CreateTexture(scl->tex1, bGammaCorrect, nullptr); CreateTexture(scl->tex2, bGammaCorrect, nullptr); }
        created.push_back({scl ? 2 : (std::dynamic_pointer_cast<pbrt::CheckerTexture>(tex) ? 1 : 0), bGammaCorrect ? 1 : 0});
        return (UINT)created.size() - 1;
    }
};

#include "../_ref/flatten_vertex_struct_gen.inc"
#include "../_ref/flatten_gen.inc"

static_assert(sizeof(Material) == 84 && sizeof(Light) == 104 && sizeof(Vertex) == 32, "SharedShaderStructs.h layout");
typedef uint32_t UINT32;

// The geometry of LoadScene's loop iteration `shapeIndex` (top-level shapes first, then the object instances)
// exactly as it is written into the upload buffers: which shape and transform the iteration takes (:1358-1376), whether
// the transform is baked (:1623-1624), the per-vertex loop (positions, normals, uvs, tangents; :1638-1661) and the index
// loop with its flat-normal rule for meshes without normals (:1704-1730), all compiled from the mount.
// `insertInstancesIntoBLAS` is LoadScene's local switch of that name (:1355; the reference build has it false).
extern "C" __attribute__((visibility("default")))
int ref_flatten_geometry(const char* pbrtPath, int shapeIndex, float* positions3, Vertex* vertices, uint32_t* indices, int capVerts, int capIndices, int* counts,
                         int insertInstancesIntoBLAS) {
    using pbrt::math::normalize; using pbrt::math::xfmNormal;
    pbrt::Scene::SP pScene;
    try { pScene = pbrt::importPBRT(pbrtPath); } catch (...) { return -1; }
    if (!pScene || !pScene->world || shapeIndex < 0 || shapeIndex >= (int)(pScene->world->shapes.size() + pScene->world->instances.size())) return -1;
    const UINT i = (UINT)shapeIndex;
    const bool bInsertInstancesIntoBLAS = insertInstancesIntoBLAS != 0;
#include "../_ref/flatten_select_gen.inc"
    pbrt::TriangleMesh::SP pTriangleMesh = std::dynamic_pointer_cast<pbrt::TriangleMesh>(pGeometry);
    if (!pTriangleMesh) return -2;
    if ((int)pTriangleMesh->vertex.size() > capVerts || (int)pTriangleMesh->index.size() * 3 > capIndices) return -1;
#include "../_ref/flatten_bake_gen.inc"
    bool bNormalsProvided = pTriangleMesh->normal.size();
    Vertex* pVertexBufferData = vertices;
    float3* pPositionBufferData = (float3*)positions3;
    UINT32* pIndexBufferData = indices;
#include "../_ref/flatten_vertices_gen.inc"
#include "../_ref/flatten_indices_gen.inc"
    counts[0] = (int)pTriangleMesh->vertex.size(); counts[1] = (int)pTriangleMesh->index.size() * 3;
    return 0;
}

// out: materials in MaterialTracker order, one material index per world->shapes entry that is a triangle mesh (-1 otherwise),
// the area lights in creation order. Returns 0, or -1 when a capacity is too small / the scene cannot be read.
extern "C" __attribute__((visibility("default")))
int ref_flatten(const char* pbrtPath, Material* mats, int capMats, int* shapeMaterial, int capShapes, Light* lights, int capLights, int* counts,
                int insertInstancesIntoBLAS) {
    pbrt::Scene::SP pScene;
    try { pScene = pbrt::importPBRT(pbrtPath); } catch (...) { return -1; }
    if (!pScene || !pScene->world) return -1;
    MaterialTracker m_MaterialTracker;
    TextureAllocator textureAllocator;
    std::vector<Light> lightList;
    int nShapes = 0;
    const bool bInsertInstancesIntoBLAS = insertInstancesIntoBLAS != 0;
    const UINT totalSceneInstances = pScene->world->shapes.size() + (bInsertInstancesIntoBLAS ? pScene->world->instances.size() : 0);
    for (UINT i = 0; i < totalSceneInstances; i++) {
        if (nShapes >= capShapes) return -1;
#include "../_ref/flatten_select_gen.inc"
        pbrt::TriangleMesh::SP pTriangleMesh = std::dynamic_pointer_cast<pbrt::TriangleMesh>(pGeometry);
        if (!pTriangleMesh) { shapeMaterial[nShapes++] = -1; continue; }
        pbrt::vec3f emissive(0.0f);
        if (pTriangleMesh->areaLight) {
            emissive = GetAreaLightColor(pTriangleMesh->areaLight);
            UINT numTriangles = pTriangleMesh->index.size();
            for (UINT i = 0; i < numTriangles; i++) {
#include "../_ref/flatten_light_gen.inc"
            }
        }
        UINT materialIndex = 0;
        if (m_MaterialTracker.Exists(pTriangleMesh->material.get())) materialIndex = m_MaterialTracker.GetMaterial(pTriangleMesh->material.get());
        else materialIndex = m_MaterialTracker.AddMaterial(pTriangleMesh->material.get(), CreateMaterial(
                 pTriangleMesh->material,
                 pTriangleMesh->textures.find("alpha") != pTriangleMesh->textures.end() ? &pTriangleMesh->textures["alpha"] : nullptr,
                 emissive, m_MaterialTracker, textureAllocator));
        shapeMaterial[nShapes++] = (int)materialIndex;
    }
    if ((int)m_MaterialTracker.MaterialList.size() > capMats || (int)lightList.size() > capLights) return -1;
    memcpy(mats, m_MaterialTracker.MaterialList.data(), sizeof(Material) * m_MaterialTracker.MaterialList.size());
    memcpy(lights, lightList.data(), sizeof(Light) * lightList.size());
    counts[0] = (int)m_MaterialTracker.MaterialList.size(); counts[1] = nShapes; counts[2] = (int)lightList.size(); counts[3] = (int)textureAllocator.created.size();
    return 0;
}
