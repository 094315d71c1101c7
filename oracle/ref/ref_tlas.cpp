// ref_tlas.cpp — CPU ORACLE (test infrastructure): the arithmetic of the top-level instance load compiled from the mount
// — AABBtoBoundingBox, BoundingBoxToAABB, Determinant, InverseAffineTransform and TransformAABB of
// D3D12RaytracingFallback/src/RayTracingHelper.hlsli (:229-243, 287-344), which TopLevelLoadAABBs.hlsli:58-100 and
// TopLevelComputeAABBs.hlsl apply to every instance — pre-passed into oracle/_ref/tlas_gen.inc by prepass.run_tlas.
// Restated here: the matrix type (AffineMatrix = float3x4 with [row][column] access), mul(float3x4, float4) as one
// left-to-right dot product per row (the pin of the primitive load's transform), rcp(x) = 1 / x, the float4 constructors.
#define RC_TRAVERSE 1
#define RC_LOAD 1
#include "hlsl_compat.h"
#include <cfloat>
#include <cstring>

namespace refcore {

struct Row4 { float v[4]; float& operator[](int i) { return v[i]; } float operator[](int i) const { return v[i]; } };
struct AffineMatrix { Row4 r[3]; Row4& operator[](int i) { return r[i]; } const Row4& operator[](int i) const { return r[i]; } };
struct AABB { float3 min, max; };                 // RayTracingHlslCompat.h:40-45
struct BoundingBox { float3 center, halfDim; };   // :47-52
inline float rcp(float x) { return 1.0f / x; }
inline float4 make4(float3 a, float w) { return float4(a.x, a.y, a.z, w); }
inline float4 make4(float2 a, float z, float w) { return float4(a.x, a.y, z, w); }
inline float4 make4(float x, float2 b, float w) { return float4(x, b.x, b.y, w); }
inline float4 make4(float x, float y, float z, float w) { return float4(x, y, z, w); }
inline float3 mul(const AffineMatrix& m, float4 v) {
    return float3(((m[0][0] * v.x + m[0][1] * v.y) + m[0][2] * v.z) + m[0][3] * v.w,
                  ((m[1][0] * v.x + m[1][1] * v.y) + m[1][2] * v.z) + m[1][3] * v.w,
                  ((m[2][0] * v.x + m[2][1] * v.y) + m[2][2] * v.z) + m[2][3] * v.w);
}

#include "../_ref/tlas_gen.inc"

} // namespace refcore

extern "C" __attribute__((visibility("default")))
void ref_inverse_affine(const float* m12, float* out12) {
    using namespace refcore;
    AffineMatrix a;
    memcpy(&a, m12, 48);
    AffineMatrix r = InverseAffineTransform(a);
    memcpy(out12, &r, 48);
}
// the world box of an instance exactly as TopLevelLoadAABBs.hlsli:73-88 forms it: BoundingBoxToAABB of the bottom-level
// root (centre / half), TransformAABB through ObjectToWorld, AABBtoBoundingBox; out = centre xyz, half xyz
extern "C" __attribute__((visibility("default")))
void ref_instance_box(const float* center3, const float* half3, const float* objectToWorld12, float* out6) {
    using namespace refcore;
    BoundingBox b;
    b.center = float3(center3[0], center3[1], center3[2]); b.halfDim = float3(half3[0], half3[1], half3[2]);
    AffineMatrix m;
    memcpy(&m, objectToWorld12, 48);
    BoundingBox w = AABBtoBoundingBox(TransformAABB(BoundingBoxToAABB(b), m));
    out6[0] = w.center.x; out6[1] = w.center.y; out6[2] = w.center.z; out6[3] = w.halfDim.x; out6[4] = w.halfDim.y; out6[5] = w.halfDim.z;
}
