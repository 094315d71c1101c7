// ref_boxes.cpp — CPU ORACLE (test infrastructure): the reference's own box constructors of the BVH node writer —
// CreateFlag (D3D12RaytracingFallback/src/RayTracingHelper.hlsli:97-103), AABBtoBoundingBox / BoundingBoxToAABB (:229-243),
// GetBoxDataFromTriangle, GetMinCorner, GetMaxCorner, GetBoxFromChildBoxes (:251-285) — pre-passed from the mount into
// oracle/_ref/boxes_gen.inc and compiled as host C++. Restated: the structs AABB / BoundingBox (:57-61,
// RayTracingHlslCompat.h:40-45) and the two constants AABB_Min_Padding, IsLeafFlag (:19, 24).
#include "hlsl_compat.h"

namespace refcore {

struct uint2 { uint x, y; };
struct AABB { float3 min, max; };
struct BoundingBox { float3 center, halfDim; };
#define AABB_Min_Padding 0.001
static const int IsLeafFlag = 0x80000000;

#include "../_ref/boxes_gen.inc"

} // namespace refcore

extern "C" __attribute__((visibility("default")))
unsigned int ref_leaf_box(const float* v, int triangleIndex, float* c3, float* h3) {
    using namespace refcore;
    uint2 flag;
    BoundingBox b = GetBoxDataFromTriangle(float3(v[0], v[1], v[2]), float3(v[3], v[4], v[5]), float3(v[6], v[7], v[8]), triangleIndex, flag);
    c3[0] = b.center.x; c3[1] = b.center.y; c3[2] = b.center.z; h3[0] = b.halfDim.x; h3[1] = b.halfDim.y; h3[2] = b.halfDim.z;
    return flag.x;
}
extern "C" __attribute__((visibility("default")))
void ref_parent_box(const float* ac, const float* ah, const float* bc, const float* bh, float* c3, float* h3) {
    using namespace refcore;
    BoundingBox a{float3(ac[0], ac[1], ac[2]), float3(ah[0], ah[1], ah[2])}, b{float3(bc[0], bc[1], bc[2]), float3(bh[0], bh[1], bh[2])};
    uint2 flag;
    BoundingBox p = GetBoxFromChildBoxes(a, 1, b, 2, flag);
    c3[0] = p.center.x; c3[1] = p.center.y; c3[2] = p.center.z; h3[0] = p.halfDim.x; h3[1] = p.halfDim.y; h3[2] = p.halfDim.z;
}
