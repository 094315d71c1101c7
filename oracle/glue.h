// glue.h — CPU ORACLE (test infrastructure): the per-pixel context and the restated
// RayGenCommon.h / SharedHitGroup.h / SharedRaytracing.h glue functions, exported so that
// oracle/_ref (the reference's kernel.glsl compiled as host C++) can call the very same glue
// the hand-restated core (core.cpp) uses. Everything here is in namespace oracle.
#pragma once
#include "../tracerboy_b200/csrc/common/tb_vec.h"
#include "oracle.h"

#pragma GCC visibility push(default)
namespace oracle {
using namespace tbm;

const float EPSILON = 0.000001f;        // kernel.glsl:1 (redefinition wins, SURVEY §8c trap 1)
const float PI = 3.1415926535f;         // kernel.glsl:2
const float LARGE_NUMBER = 1e20f;
const float AIR_IOR = 1.0f;
const float MIN_ROUGHNESS = 0.04f;
const float MIN_ROUGHNESS_SQUARED = MIN_ROUGHNESS * MIN_ROUGHNESS;
const float MIN_T = 0.001f;             // RayGenCommon.h:364

inline f3 F3(const TbFloat3& v) { return mk3(v.x, v.y, v.z); }

struct Material { // SharedShaderStructs.h:141-161 in registers
    f3 albedo; uint32_t albedoIndex, alphaIndex, normalMapIndex, emissiveIndex, specularMapIndex;
    float IOR; f3 absorption; float roughness; f3 scattering; f3 emissive; int Flags; float SpecularCoef;
};
inline Material load_material(const TbMaterial& m) {
    Material r;
    r.albedo = F3(m.albedo); r.albedoIndex = m.albedoIndex; r.alphaIndex = m.alphaIndex;
    r.normalMapIndex = m.normalMapIndex; r.emissiveIndex = m.emissiveIndex; r.specularMapIndex = m.specularMapIndex;
    r.IOR = m.IOR; r.absorption = F3(m.absorption); r.roughness = m.roughness; r.scattering = F3(m.scattering);
    r.emissive = F3(m.emissive); r.Flags = m.Flags; r.SpecularCoef = m.SpecularCoef;
    return r;
}

struct Ray { f3 origin, direction; };

struct Ctx {
    const Scene& sc;
    const RenderParams& rp;
    uint32_t W, H, px, py;
    float seed;
    float* seedp; // the rand() counter in use: &seed, or the reference core's own global when oracle/_ref drives the path
    // per-pixel side outputs
    f3 worldPosition; float distanceToNeighbor;
    f4 aovAlbedo, aovNormal, aovEmissive; bool wroteEmissive;
    float aovDepth; bool wroteDepth;
    uint32_t primGeom, primPrim; bool firstIntersect;
    uint32_t tris, boxes, rays;
    float statDistance; int statMaterial; bool wroteStats;
    Ctx(const Scene& s, const RenderParams& r) : sc(s), rp(r), seedp(&seed) {}

    float rand() { float s = *seedp; *seedp = s + 1.0f; return frac(sin_(s + rp.time) * 43758.5453123f); } // kernel.glsl:39-40
    bool selected() const { return (int)px == rp.selectedX && (int)py == rp.selectedY; }
};

struct BlueNoiseData { f2 PrimaryJitter, SecondaryRayDirection, AreaLightJitter, DOFJitter; };
struct HitResult { float t; int material; f3 normal, tangent; f2 uv; };

// glue (definitions in core.cpp; file:line of the reference in the comments there)
BlueNoiseData get_blue_noise(Ctx& c);                                   // RayGenCommon.h:104-122
HitResult intersect(Ctx& c, const Ray& ray, float maxT = 999999.0f);    // IntersectWithMaxDistance, :365-414
Material get_material_internal(Ctx& c, int id, f2 uv, bool backside);   // :298-341
f3 get_detail_normal(Ctx& c, const Material& mat, f3 normal, f3 tangent, f2 uv); // :273-295
void get_one_light_sample(Ctx& c, f3 pos, f3& LightDirection, f3& LightColor, float& PDFValue, f3& LightNormal, float& LightAttenuation); // :170-261
f3 sample_environment_map(const Scene& sc, f3 v);                       // :21-44
f4 get_texture_data(const Scene& sc, uint32_t index, f2 uv);            // SharedRaytracing.h:67-137
f4 sample_bilinear_wrap(const Image& im, f2 uv);                        // SampleLevel(BilinearSampler, uv, 0), pinned weights
float hash13(f3 p3);                                                    // :662-667
f4 path_trace(Ctx& c, f2 pixelCoord);                                   // the hand-restated PathTrace (core.cpp)

// oracle/_ref hook: when set, render_frame calls this instead of the restated path_trace()
typedef f4 (*PathTraceFn)(Ctx& c, f2 pixelCoord);
void set_path_trace_override(PathTraceFn fn);

} // namespace oracle
#pragma GCC visibility pop
