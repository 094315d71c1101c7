// ref_glue.cpp — CPU ORACLE (test infrastructure): compiles the reference's own
// TracerBoy/kernel.glsl (pre-passed into oracle/_ref/kernel_gen.inc by prepass.py) as host C++
// and plugs it into the oracle's render loop in place of the hand-restated PathTrace/Trace.
//
// What comes from the reference text: everything kernel.glsl defines outside its IS_SHADER_TOY
// blocks — rand(), the BRDF/sampling helpers, GetMaterial, Trace, PathTrace.
// What is restated (oracle/glue.h, shared with core.cpp): the RayGenCommon.h / SharedHitGroup.h /
// SharedRaytracing.h glue the shader text calls (resource access cannot be compiled here), and
// the BVH traversal.
#include "hlsl_compat.h"
#include "../glue.h"

namespace refcore {

// ---- SharedShaderStructs.h as seen by the shader (HLSL side)
#define DEFAULT_MATERIAL_FLAG 0x0
#define METALLIC_MATERIAL_FLAG 0x1
#define SUBSURFACE_SCATTER_MATERIAL_FLAG 0x2
#define NO_SPECULAR_MATERIAL_FLAG 0x4
#define MIX_MATERIAL_FLAG 0x8
#define LIGHT_MATERIAL_FLAG 0x10
#define NO_ALPHA_MATERIAL_FLAG 0x20
#define HAIR_MATERIAL_FLAG 0x40
#define SINGLE_SIDED_MATERIAL_FLAG 0x80
#define OUTPUT_TYPE_HEATMAP 9
#define FILTER_TYPE_BOX 0
#define FILTER_TYPE_TRIANGLE 1
#define FILTER_TYPE_GAUSSIAN 2
#define INVALID_MATERIAL_ID -1   // kernel.glsl:614 (inside a block this build does not need otherwise)
#define GLOBAL static thread_local

struct Material { // SharedShaderStructs.h:141-161
    float3 albedo; uint albedoIndex, alphaIndex, normalMapIndex, emissiveIndex, specularMapIndex;
    float IOR; float3 absorption; float roughness; float3 scattering; float3 emissive; int Flags; float SpecularCoef;
};
struct Ray { float3 origin; float3 direction; }; // RayGenCommon.h:343-347
struct PerFrame { // the PerFrameConstants members the shader text reads
    uint MaxBounces, EnableNextEventEstimation, OutputMode, FilterType;
    float2 FixedPixelOffset; float FilterWidth, DOFFocusDistance, DOFApertureWidth, FireflyClampValue;
};
struct BlueNoiseData { float2 PrimaryJitter, SecondaryRayDirection, AreaLightJitter, DOFJitter; }; // RayGenCommon.h:70-76

static thread_local oracle::Ctx* g_ctx = nullptr;
static thread_local PerFrame perFrameConstants;

static Material from_oracle(const oracle::Material& m) {
    Material r;
    r.albedo = float3(m.albedo); r.albedoIndex = m.albedoIndex; r.alphaIndex = m.alphaIndex; r.normalMapIndex = m.normalMapIndex;
    r.emissiveIndex = m.emissiveIndex; r.specularMapIndex = m.specularMapIndex; r.IOR = m.IOR; r.absorption = float3(m.absorption);
    r.roughness = m.roughness; r.scattering = float3(m.scattering); r.emissive = float3(m.emissive); r.Flags = m.Flags; r.SpecularCoef = m.SpecularCoef;
    return r;
}
static oracle::Material to_oracle(const Material& m) {
    oracle::Material r;
    r.albedo = m.albedo.t(); r.albedoIndex = m.albedoIndex; r.alphaIndex = m.alphaIndex; r.normalMapIndex = m.normalMapIndex;
    r.emissiveIndex = m.emissiveIndex; r.specularMapIndex = m.specularMapIndex; r.IOR = m.IOR; r.absorption = m.absorption.t();
    r.roughness = m.roughness; r.scattering = m.scattering.t(); r.emissive = m.emissive.t(); r.Flags = m.Flags; r.SpecularCoef = m.SpecularCoef;
    return r;
}

// ---- RayGenCommon.h accessors (:9-19, 85-103) over the oracle's per-pixel context
float GetTime() { return g_ctx->rp.time; }
float3 GetResolution() { return float3((float)g_ctx->W, (float)g_ctx->H, 1.0f); }
float GetRotationFactor() { return 0.5f; }
bool ShouldInvalidateHistory() { return false; }
float3 GetCameraPosition() { return float3(oracle::F3(g_ctx->rp.camera.Position)); }
float3 GetCameraLookAt() { return float3(oracle::F3(g_ctx->rp.camera.LookAt)); }
float3 GetCameraUp() { return float3(oracle::F3(g_ctx->rp.camera.Up)); }
float3 GetCameraRight() { return float3(oracle::F3(g_ctx->rp.camera.Right)); }
float GetCameraLensHeight() { return g_ctx->rp.camera.LensHeight; }
float GetCameraFocalDistance() { return g_ctx->rp.camera.FocalDistance; }
bool IsTargettingRealTime() { return g_ctx->rp.settings.RenderMode == TB_RENDER_REALTIME; }
float4 GetLastFrameData() { return float4(0.0f); }
float4 GetAccumulatedColor(float2) { return float4(0.0f); } // only feeds the ShaderToy accumulation path

BlueNoiseData GetBlueNoise() { // :104-122 (restated in glue; draws from the shader's rand() stream through seedp)
    oracle::BlueNoiseData d = oracle::get_blue_noise(*g_ctx);
    BlueNoiseData r;
    r.PrimaryJitter = float2(d.PrimaryJitter); r.SecondaryRayDirection = float2(d.SecondaryRayDirection);
    r.AreaLightJitter = float2(d.AreaLightJitter); r.DOFJitter = float2(d.DOFJitter);
    return r;
}
float2 IntersectWithMaxDistance(Ray ray, float maxT, float3& normal, float3& tangent, float2& uv, uint& PrimitiveID) { // :365-487
    oracle::Ray r = {ray.origin.t(), ray.direction.t()};
    oracle::HitResult h = oracle::intersect(*g_ctx, r, maxT);
    normal = float3(h.normal); tangent = float3(h.tangent); uv = float2(h.uv); PrimitiveID = 0;
    return float2(h.t, (float)h.material);
}
Material GetMaterialInternal(int MaterialID, uint, float3, float2 uv, bool IsBacksideOfGeometry) { // :298-341
    return from_oracle(oracle::get_material_internal(*g_ctx, MaterialID, uv.t(), IsBacksideOfGeometry));
}
float3 GetDetailNormal(Material mat, float3 normal, float3 tangent, float2 uv) { // :273-295
    return float3(oracle::get_detail_normal(*g_ctx, to_oracle(mat), normal.t(), tangent.t(), uv.t()));
}
float3 SampleEnvironmentMap(float3 v) { return float3(oracle::sample_environment_map(g_ctx->sc, v.t())); } // :21-44
void GetOneLightSample(float3 P, float3& LightDirection, float3& LightColor, float& PDFValue, float3& LightNormal, float& LightAttenuation) { // :170-261
    tbm::f3 d, c, n;
    oracle::get_one_light_sample(*g_ctx, P.t(), d, c, PDFValue, n, LightAttenuation);
    LightDirection = float3(d); LightColor = float3(c); LightNormal = float3(n);
}
// AOV writers (:524-654)
void OutputPrimaryAlbedo(float3 albedo, float) { g_ctx->aovAlbedo = tbm::mk4(albedo.t(), 1.0f); }
void OutputPrimaryEmissive(float3 e) { g_ctx->aovEmissive = tbm::mk4(e.t(), 1.0f); g_ctx->wroteEmissive = true; }
void OutputPrimaryNormal(float3 n) { g_ctx->aovNormal = tbm::mk4(n.t(), 1.0f); }
void OutputPrimaryWorldPosition(float3 p, float d) { g_ctx->worldPosition += p.t(); g_ctx->distanceToNeighbor += d; }
void OutputDistanceToFirstHit(float Distance) {
    g_ctx->aovDepth = tbm::saturate(Distance / g_ctx->rp.settings.MaxZ); g_ctx->wroteDepth = true;
    if (g_ctx->selected()) { g_ctx->statDistance = Distance; g_ctx->wroteStats = true; }
}
void OutputMaterial(int id) { if (g_ctx->selected()) { g_ctx->statMaterial = id; g_ctx->wroteStats = true; } }
void OutputVisualizationRay(float3, float3, float, float) {} // debug overlay, out of scope

#include "../_ref/kernel_gen.inc"

} // namespace refcore

#include <atomic>
static std::atomic<unsigned long long> g_calls{0};

// Drop-in for oracle::path_trace: runs the reference's PathTrace on the oracle's pixel context.
static tbm::f4 ref_path_trace(oracle::Ctx& c, tbm::f2 pixelCoord) {
    using namespace refcore;
    g_calls.fetch_add(1, std::memory_order_relaxed);
    g_ctx = &c;
    const TbOutputSettings& S = c.rp.settings;
    perFrameConstants.MaxBounces = (uint)S.MaxBounces;
    perFrameConstants.EnableNextEventEstimation = S.EnableNextEventEstimation;
    perFrameConstants.OutputMode = S.OutputType;
    perFrameConstants.FilterType = S.FilterType;
    perFrameConstants.FixedPixelOffset = float2(-1.0f, -1.0f); // TracerBoy.cpp:2838
    perFrameConstants.FilterWidth = S.FilterWidth;
    perFrameConstants.DOFFocusDistance = S.DOFFocalDistance;
    perFrameConstants.DOFApertureWidth = S.ApertureWidth;
    perFrameConstants.FireflyClampValue = S.FireflyClampValue;
    seed = c.seed;          // kernel.glsl:39, the shader's own counter
    c.seedp = &seed;        // glue draws (mix pick, light sample, blue noise off) advance the same counter
    float4 r = PathTrace(float2(pixelCoord.x, pixelCoord.y));
    c.seed = seed;
    c.seedp = &c.seed;
    return tbm::mk4(r.x, r.y, r.z, r.w);
}

// number of pixels the reference core has traced (tests assert the override really ran)
extern "C" __attribute__((visibility("default"))) unsigned long long ref_core_calls() { return g_calls.load(); }

extern "C" __attribute__((visibility("default"))) void ref_core_enable(int on) {
    oracle::set_path_trace_override(on ? ref_path_trace : nullptr);
}
