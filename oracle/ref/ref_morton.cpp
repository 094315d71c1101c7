// ref_morton.cpp — CPU ORACLE (test infrastructure): the reference's own GetMortonCodesFromUnitCoord / CalculateMortonCode
// (D3D12RaytracingFallback/src/CalculateMortonCodesBindings.h:116-149, the non-SCALED_MORTON_CODES branch), pre-passed from
// the mount into oracle/_ref/morton_gen.inc and compiled as host C++. Restated here: the scene-AABB resource access
// (GetSceneAABB reads a RWByteAddressBuffer) and `pow(2, numBits)`, which the shader compiler folds to exactly 1024.
#include "hlsl_compat.h"

namespace refcore {

struct AABB { float3 min, max; };
static thread_local AABB g_sceneAABB;
inline AABB GetSceneAABB() { return g_sceneAABB; }
inline unsigned int pow(int base, unsigned int e) { unsigned int r = 1; while (e--) r *= (unsigned int)base; return r; } // exact
inline float3 max(float3 a, double s) { return max(a, float3((float)s)); }

#include "../_ref/morton_gen.inc"

} // namespace refcore

extern "C" __attribute__((visibility("default")))
unsigned int ref_morton(const float* centroid, const float* smin, const float* smax) {
    using namespace refcore;
    g_sceneAABB.min = float3(smin[0], smin[1], smin[2]);
    g_sceneAABB.max = float3(smax[0], smax[1], smax[2]);
    return CalculateMortonCode(float3(centroid[0], centroid[1], centroid[2]));
}
