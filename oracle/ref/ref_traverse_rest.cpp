// ref_traverse_rest.cpp — CPU ORACLE (test infrastructure): the reference's own GetIndexOfBiggestChannel, swap, RayData,
// GetRayData (TraverseFunction.hlsli:431-452, 463-495) and Swizzle, IsPositive, RayTriangleIntersect (:223-313), pre-passed
// from the mount into oracle/_ref/traverse_rest_gen.inc and compiled as host C++ without contraction (`precise`).
#define RC_TRAVERSE 1
#include "hlsl_compat.h"
#include <cstring>

namespace refcore {

struct int3 { int x, y, z; int3() : x(0), y(0), z(0) {} int3(int a, int b, int c) : x(a), y(b), z(c) {} };
inline float3 rcp(float3 v) { return float3(tbm::rcp(v.x), tbm::rcp(v.y), tbm::rcp(v.z)); }
inline float rcp(float x) { return tbm::rcp(x); }
#define precise
static const uint D3D12_RAYTRACING_INSTANCE_FLAG_TRIANGLE_CULL_DISABLE = 0x1;          // RayTracingHlslCompat.h:208-209
static const uint D3D12_RAYTRACING_INSTANCE_FLAG_TRIANGLE_FRONT_COUNTERCLOCKWISE = 0x2;
#define RAY_FLAG_CULL_BACK_FACING_TRIANGLES 0x10                                          // DXR RAY_FLAG values
#define RAY_FLAG_CULL_FRONT_FACING_TRIANGLES 0x20

#include "../_ref/traverse_rest_gen.inc"

// out-of-line entry for ref_traverse_loop.cpp (the shader's function is `inline` and leaves no symbol of its own)
void RayTriangleIntersect_precise(float& hitT, uint rayFlags, uint instanceFlags, float2& bary, float3 rayOrigin, float3 rayDirection,
                                  int3 swizzledIndicies, float3 shear, float3 v0, float3 v1, float3 v2) {
    RayTriangleIntersect(hitT, rayFlags, instanceFlags, bary, rayOrigin, rayDirection, swizzledIndicies, shear, v0, v1, v2);
}

} // namespace refcore

extern "C" __attribute__((visibility("default")))
void ref_ray_data(const float* org, const float* dir, float* inv, float* oinv, float* shear, int* swz) {
    using namespace refcore;
    RayData d = GetRayData(float3(org[0], org[1], org[2]), float3(dir[0], dir[1], dir[2]));
    inv[0] = d.InverseDirection.x; inv[1] = d.InverseDirection.y; inv[2] = d.InverseDirection.z;
    oinv[0] = d.OriginTimesRayInverseDirection.x; oinv[1] = d.OriginTimesRayInverseDirection.y; oinv[2] = d.OriginTimesRayInverseDirection.z;
    shear[0] = d.Shear.x; shear[1] = d.Shear.y; shear[2] = d.Shear.z;
    swz[0] = d.SwizzledIndices.x; swz[1] = d.SwizzledIndices.y; swz[2] = d.SwizzledIndices.z;
}

// RAY_FLAG_NONE, instance flags 0: TracerBoy's configuration (two-sided). Returns 1 when the function reached its end
// (hit accepted: hitT and bary written), 0 when it returned early.
extern "C" __attribute__((visibility("default")))
int ref_ray_tri(float* hitT, const float* org, const int* swz, const float* shear, const float* v9, float* bary) {
    using namespace refcore;
    const uint32_t sentinelBits = 0x7fc0beefu;
    float sentinel;
    memcpy(&sentinel, &sentinelBits, 4);
    float2 b(sentinel, sentinel);
    RayTriangleIntersect(*hitT, 0u, 0u, b, float3(org[0], org[1], org[2]), float3(0.0f), int3(swz[0], swz[1], swz[2]),
                         float3(shear[0], shear[1], shear[2]), float3(v9[0], v9[1], v9[2]), float3(v9[3], v9[4], v9[5]), float3(v9[6], v9[7], v9[8]));
    uint32_t bits;
    memcpy(&bits, &b.x, 4);
    if (bits == sentinelBits) return 0;
    bary[0] = b.x; bary[1] = b.y;
    return 1;
}
