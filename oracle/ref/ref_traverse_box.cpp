// ref_traverse_box.cpp — CPU ORACLE (test infrastructure): the reference's own RayBoxTest (D3D12RaytracingFallback/src/
// TraverseFunction.hlsli:201-221), pre-passed from the mount into oracle/_ref/traverse_box_gen.inc and compiled as host
// C++ with floating-point contraction ON: HLSL leaves a*b+-c contraction to the driver and the pinned semantics of this
// repository make each of the slab test's nine products one fused multiply-add.
#include "hlsl_compat.h"

namespace refcore {
#include "../_ref/traverse_box_gen.inc"
// out-of-line entry for ref_traverse_loop.cpp (the shader's function is `inline` and leaves no symbol of its own)
bool RayBoxTest_fused(float& resultT, float closestT, float3 rayOriginTimesRayInverseDirection, float3 rayInverseDirection, float3 boxCenter, float3 boxHalfDim) {
    return RayBoxTest(resultT, closestT, rayOriginTimesRayInverseDirection, rayInverseDirection, boxCenter, boxHalfDim);
}
} // namespace refcore

extern "C" __attribute__((visibility("default")))
int ref_ray_box(float closestT, const float* oinv, const float* inv, const float* c, const float* h, float* resultT) {
    using namespace refcore;
    return RayBoxTest(*resultT, closestT, float3(oinv[0], oinv[1], oinv[2]), float3(inv[0], inv[1], inv[2]),
                      float3(c[0], c[1], c[2]), float3(h[0], h[1], h[2])) ? 1 : 0;
}
