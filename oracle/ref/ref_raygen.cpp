// ref_raygen.cpp — CPU ORACLE (test infrastructure): pure pieces of the reference's RayGenCommon.h glue compiled from the
// mount as host C++ (oracle/_ref/raygen_gen.inc): `struct Light` with GetPosition (SharedShaderStructs.h:92-114),
// SampleEnvironmentMap (RayGenCommon.h:20-44), Halton / Halton2 / Halton23 (:48-69), GetRandomBarycentric,
// RestirReservoir, GetLightTargetPDF, GetOneLightSample (:124-261) and hash13 (:662-667), plus ColorToLuma from
// Tonemap.h. Restated here: the constant buffers, the light list and the environment texture as plain objects (the
// bilinear fetch itself is the oracle's pinned sampler), `float4x3` and `mul(v, M)` as SURVEY §8c trap 19 reads them,
// and rand() (kernel.glsl:39-40) with its seed / Time exposed.
#define HLSL 1
#define RC_MATERIAL 1
#include "hlsl_compat.h"
#include <climits>
#include <cstring>
#include "../glue.h"

namespace refcore {

struct float4x3 { float4 c0, c1, c2; }; // three float4 columns (ConfigConstants.EnvironmentMapTransform)
inline float3 mul(float3 v, const float4x3& m) { return float3(dot(v, m.c0.xyz()), dot(v, m.c1.xyz()), dot(v, m.c2.xyz())); }
inline float atan2(float y, float x) { return tbm::atan2_(y, x); }
struct PerFrame { uint LightCount, EnableNextEventEstimation, EnableSamplingImportanceResampling, GlobalFrameCount, IsRealTime; float DebugValue, DebugValue2, Time; uint EnableNormalMaps; };
struct Config { float4x3 EnvironmentMapTransform; float3 EnvironmentMapColorScale; uint FlipTextureUVs; };
static thread_local PerFrame perFrameConstants;
static thread_local Config configConstants;
struct SamplerShim {} static BilinearSampler;
struct TextureShim {
    const oracle::Image* image = nullptr;
    float4 SampleLevel(SamplerShim, float2 uv, float) const {
        tbm::f4 t = oracle::sample_bilinear_wrap(*image, uv.t());
        return float4(t.x, t.y, t.z, t.w);
    }
};
static thread_local TextureShim EnvironmentMap;
static thread_local float g_seed;
inline float rand() { float s = g_seed; g_seed = s + 1.0f; return frac(sin(s + perFrameConstants.Time) * 43758.5453123f); } // kernel.glsl:39-40
inline float ColorToLuma(float3 color) { return dot(color, float3(0.212671, 0.715160, 0.072169)); }                  // Tonemap.h:12-15

#include "../_ref/raygen_light_gen.inc"
static thread_local const Light* LightList;
#include "../_ref/raygen_gen.inc"

// ---- material and texture fetch (prepass.run_material -> oracle/_ref/raygen_material_gen.inc): struct Material / TextureData
// and the texture / material flag values, GammaToLinear, GetMaterial_NonRecursive ... GetTextureData_Recursive
// (SharedRaytracing.h:55-137), GetDetailNormal and GetMaterialInternal (RayGenCommon.h:273-341). Shims: the structured
// buffers as pointers into the oracle's scene, the image table (the bilinear fetch itself is the oracle's pinned sampler,
// as for the environment map).
struct Material;
struct TextureData;
template <typename T> struct BufferShim { const T* p = nullptr; const T& operator[](uint i) const { return p[i]; } };
struct ImageTableShim {
    const oracle::Image* images = nullptr;
    TextureShim operator[](uint i) const { TextureShim t; t.image = images + i; return t; }
};
inline uint NonUniformResourceIndex(uint i) { return i; }
static thread_local ImageTableShim ImageTextures;
static thread_local BufferShim<Material> MaterialBuffer;       // holds pointers only: the structs are defined by the text below
static thread_local BufferShim<TextureData> TextureDataBuffer;
#include "../_ref/raygen_material_gen.inc"
} // namespace refcore

extern "C" __attribute__((visibility("default")))
void ref_light_sample(const TbLight* lights, unsigned int n, unsigned int nee, unsigned int sir, float debug1, float debug2, float time,
                      const float* pos, float* seed, float* out12) {
    using namespace refcore;
    static_assert(sizeof(Light) == sizeof(TbLight), "Light layout");
    perFrameConstants.LightCount = n; perFrameConstants.EnableNextEventEstimation = nee;
    perFrameConstants.EnableSamplingImportanceResampling = sir; perFrameConstants.DebugValue = debug1; perFrameConstants.DebugValue2 = debug2;
    perFrameConstants.Time = time;
    LightList = (const Light*)lights;
    g_seed = *seed;
    float3 dir, col, nrm; float pdf, att;
    GetOneLightSample(float3(pos[0], pos[1], pos[2]), dir, col, pdf, nrm, att);
    *seed = g_seed;
    float o[12] = {dir.x, dir.y, dir.z, col.x, col.y, col.z, nrm.x, nrm.y, nrm.z, pdf, att, 0.0f};
    memcpy(out12, o, sizeof(o));
}

extern "C" __attribute__((visibility("default")))
void ref_env(const float* rgba, unsigned int w, unsigned int h, const float* transform12, const float* scale3, const float* v, float* out3) {
    using namespace refcore;
    oracle::Image im;
    im.width = w; im.height = h; im.format = 0;
    im.data.assign((const uint8_t*)rgba, (const uint8_t*)rgba + 16ull * w * h);
    EnvironmentMap.image = &im;
    configConstants.EnvironmentMapTransform = {float4(transform12[0], transform12[1], transform12[2], transform12[3]),
                                               float4(transform12[4], transform12[5], transform12[6], transform12[7]),
                                               float4(transform12[8], transform12[9], transform12[10], transform12[11])};
    configConstants.EnvironmentMapColorScale = float3(scale3[0], scale3[1], scale3[2]);
    float3 c = SampleEnvironmentMap(float3(v[0], v[1], v[2]));
    out3[0] = c.x; out3[1] = c.y; out3[2] = c.z;
}

static void bind_scene(const oracle::Scene& sc, unsigned int enableNormalMaps, float time, float seed) {
    using namespace refcore;
    static_assert(sizeof(Material) == sizeof(TbMaterial) && sizeof(Material) == 84, "Material layout (SharedShaderStructs.h:141-161)");
    static_assert(sizeof(TextureData) == sizeof(TbTextureData) && sizeof(TextureData) == 80, "TextureData layout (:169-190)");
    MaterialBuffer.p = (const Material*)sc.materials.data();
    TextureDataBuffer.p = (const TextureData*)sc.textures.data();
    ImageTextures.images = sc.images.data();
    configConstants.FlipTextureUVs = sc.flipTextureUVs;
    perFrameConstants.EnableNormalMaps = enableNormalMaps;
    perFrameConstants.Time = time;
    g_seed = seed;
}

// GetMaterialInternal on the oracle's scene (`scene` = oracle_scene_ptr). out21 = struct Material, 84 bytes.
extern "C" __attribute__((visibility("default")))
void ref_material(const void* scene, float time, int materialId, const float* uv, int backside, float* seed, void* out21) {
    using namespace refcore;
    bind_scene(*(const oracle::Scene*)scene, 0, time, *seed);
    Material m = GetMaterialInternal(materialId, 0u, float3(0.0f), float2(uv[0], uv[1]), backside != 0);
    *seed = g_seed;
    memcpy(out21, &m, sizeof(m));
}

// GetDetailNormal for the (untextured-fetch) material record `materialId`.
extern "C" __attribute__((visibility("default")))
void ref_detail_normal(const void* scene, unsigned int enableNormalMaps, int materialId, const float* normal, const float* tangent,
                       const float* uv, float* out3) {
    using namespace refcore;
    const oracle::Scene& sc = *(const oracle::Scene*)scene;
    bind_scene(sc, enableNormalMaps, 0.0f, 0.0f);
    float3 n = GetDetailNormal(MaterialBuffer[materialId], float3(normal[0], normal[1], normal[2]), float3(tangent[0], tangent[1], tangent[2]),
                               float2(uv[0], uv[1]));
    out3[0] = n.x; out3[1] = n.y; out3[2] = n.z;
}

// GetTextureData for texture record `textureIndex` (UINT_MAX = invalid).
extern "C" __attribute__((visibility("default")))
void ref_texture(const void* scene, unsigned int textureIndex, const float* uv, float* out4) {
    using namespace refcore;
    bind_scene(*(const oracle::Scene*)scene, 0, 0.0f, 0.0f);
    float4 t = GetTextureData(textureIndex, float2(uv[0], uv[1]));
    out4[0] = t.x; out4[1] = t.y; out4[2] = t.z; out4[3] = t.w;
}

extern "C" __attribute__((visibility("default"))) float ref_hash13(float x, float y, float z) { return refcore::hash13(refcore::float3(x, y, z)); }
extern "C" __attribute__((visibility("default"))) float ref_halton(int b, int i) { return refcore::Halton(b, i); }
