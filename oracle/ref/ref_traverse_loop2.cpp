// ref_traverse_loop2.cpp — CPU ORACLE (test infrastructure): the reference's ray-query loop compiled from the mount with
// FAST_PATH 0, i.e. the TWO-LEVEL walk of TraverseFunction.hlsli:537-785 (instance leaves :603-638, return to the top
// level :770-774), from the same pre-passed text as ref_traverse_loop.cpp (oracle/_ref/traverse_loop_gen.inc) plus the
// instance-desc readers of RayTracingHlslCompat.h:217-337 (oracle/_ref/instance_desc_gen.inc, prepass.run_instance_desc).
// Restated here: the resources (byte-address buffers behind emulated GPU pointers: pointer word 0 selects the top level
// or bottom-level structure i), float3x4 with mul(float3x4, float4) as one left-to-right dot product per row, integer
// vector types, layout constants, DXR built-in constants. The three pure functions are the objects of
// ref_traverse_box.cpp / ref_traverse_rest.cpp. tests/test_cpu_tlas.py requires oracle::trace_ray_tlas to match this build.
#define RC_TRAVERSE 1
#include "hlsl_compat.h"
#include <cfloat>
#include <cstring>
#include <vector>
#include "tracerboy_b200.h"

namespace refcore {

struct int3 { int x, y, z; int3() : x(0), y(0), z(0) {} int3(int a, int b, int c) : x(a), y(b), z(c) {} };
struct uint2 { uint x, y; uint2() : x(0), y(0) {} uint2(uint a, uint b) : x(a), y(b) {} };
struct uint3 { uint x, y, z; uint3() : x(0), y(0), z(0) {} uint3(uint a, uint b, uint c) : x(a), y(b), z(c) {} };
struct GpuVA { uint lo, hi; GpuVA(uint a = 0, uint b = 0) : lo(a), hi(b) {} GpuVA(uint2 v) : lo(v.x), hi(v.y) {} };
struct uint4 { uint x, y, z, w; uint2 zw() const { return uint2(z, w); } };
struct int4 { int x, y, z, w; int4() : x(0), y(0), z(0), w(0) {} int4(const uint4& u) : x((int)u.x), y((int)u.y), z((int)u.z), w((int)u.w) {} };
struct float4e : float4 {
    float4e() : float4() {}
    float4e(float a, float b, float c, float d) : float4(a, b, c, d) {}
    float4e(float3 v, float w_) : float4(v.x, v.y, v.z, w_) {}
    float4e(const float4& v) : float4(v) {}
    float2 zw() const { return float2(z, w); }
};
inline float asfloat(int i) { float f; memcpy(&f, &i, 4); return f; }
inline float asfloat(uint u) { float f; memcpy(&f, &u, 4); return f; }
inline float4e asfloat(uint4 u) { return float4e(asfloat(u.x), asfloat(u.y), asfloat(u.z), asfloat(u.w)); }
inline uint asuint(float f) { uint u; memcpy(&u, &f, 4); return u; }
struct float3x4 {
    float4e r[3];
    float4e& operator[](int i) { return r[i]; }
    const float4e& operator[](int i) const { return r[i]; }
};
inline float3 mul(const float3x4& m, float4e v) { // one dot product per row, left to right
    return float3(((m[0].x * v.x + m[0].y * v.y) + m[0].z * v.z) + m[0].w * v.w,
                  ((m[1].x * v.x + m[1].y * v.y) + m[1].z * v.z) + m[1].w * v.w,
                  ((m[2].x * v.x + m[2].y * v.y) + m[2].z * v.z) + m[2].w * v.w);
}
#define row_major
#define groupshared static thread_local
#define float4 float4e
#define HLSL 1
#define AffineMatrix float3x4

struct RWByteAddressBuffer {
    uint8_t* bytes = nullptr;
    uint Load(uint o) const { uint v; memcpy(&v, bytes + o, 4); return v; }
    uint3 Load3(uint o) const { uint3 v; memcpy(&v, bytes + o, 12); return v; }
    uint4 Load4(uint o) const { uint4 v; memcpy(&v, bytes + o, 16); return v; }
    void Store(uint o, uint v) { memcpy(bytes + o, &v, 4); }
    void Store4(uint o, uint4 v) { memcpy(bytes + o, &v, 16); }
};
struct RWByteAddressBufferPointer { RWByteAddressBuffer buffer; uint offsetInBytes; };
static thread_local RWByteAddressBuffer g_tlas;
static thread_local const uint8_t* const* g_blas = nullptr;
#define TOP_LEVEL_VA 0xffff0000u
static thread_local GpuVA TopLevelAccelerationStructureGpuVA(TOP_LEVEL_VA, 0);
inline RWByteAddressBufferPointer CreateRWByteAddressBufferPointerFromGpuVA(GpuVA va) {
    if (va.lo == TOP_LEVEL_VA) return RWByteAddressBufferPointer{g_tlas, 0u};
    RWByteAddressBuffer b;
    b.bytes = const_cast<uint8_t*>(g_blas[va.lo]);
    return RWByteAddressBufferPointer{b, 0u};
}
static thread_local uint GI = 0;

#define TRAVERSAL_MAX_STACK_DEPTH 16
#define SizeOfFloat 4
#define SizeOfPrimitive 40
#define OffsetToPrimitiveData 4
struct PrimitiveMetaData { uint GeometryContributionToHitGroupIndex; uint PrimitiveIndex; uint GeometryFlags; };
#define SizeOfPrimitiveMetaData (4 * 3)
#define SizeOfAABBNode (4 * 8)
#define SizeOfBVHOffsets (4 * 4)
#define SizeOfBVHMetadata 116
#define Store4StrideInBytes 16
#define D3D12_RAYTRACING_GEOMETRY_FLAG_OPAQUE 0x1
#define RAY_FLAG_NONE 0x00
#define RAY_FLAG_FORCE_OPAQUE 0x01
#define RAY_FLAG_FORCE_NON_OPAQUE 0x02
#define RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH 0x04
#define RAY_FLAG_SKIP_CLOSEST_HIT_SHADER 0x08
#define RAY_FLAG_CULL_BACK_FACING_TRIANGLES 0x10
#define RAY_FLAG_CULL_FRONT_FACING_TRIANGLES 0x20
#define RAY_FLAG_CULL_OPAQUE 0x40
#define RAY_FLAG_CULL_NON_OPAQUE 0x80
#define HIT_KIND_TRIANGLE_FRONT_FACE 0xFE
// the two-level configuration (the fallback layer's default: RayGenCommon.h:355-362 sets FAST_PATH 1 for TracerBoy)
#define FAST_PATH 0
#define DISABLE_ANYHIT
#define DISABLE_PROCEDURAL_GEOMETRY

bool RayBoxTest_fused(float& resultT, float closestT, float3 rayOriginTimesRayInverseDirection, float3 rayInverseDirection, float3 boxCenter, float3 boxHalfDim);
void RayTriangleIntersect_precise(float& hitT, uint rayFlags, uint instanceFlags, float2& bary, float3 rayOrigin, float3 rayDirection,
                                  int3 swizzledIndicies, float3 shear, float3 v0, float3 v1, float3 v2);
#define RayBoxTest RayBoxTest_fused
#define RayTriangleIntersect RayTriangleIntersect_precise
inline float3 float3_from(float a, float2 b) { return float3(a, b.x, b.y); }

#include "../_ref/instance_desc_gen.inc"
#include "../_ref/traverse_loop_gen.inc"

} // namespace refcore

// TbHit records of n two-level queries: tlas = reference-layout top-level bytes whose instance descs name their
// bottom-level structure by index into blas[]
extern "C" __attribute__((visibility("default")))
int ref_trace_rays_tlas(const uint8_t* tlas, const uint8_t* const* blas, const TbRay* rays, uint64_t n, TbHit* hits) {
    using namespace refcore;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        g_tlas.bytes = const_cast<uint8_t*>(tlas);
        g_blas = blas;
        GI = 0;
        SoftwareRayDesc rd;
        rd.Origin = float3(rays[i].Origin[0], rays[i].Origin[1], rays[i].Origin[2]); rd.TMin = rays[i].TMin;
        rd.Direction = float3(rays[i].Direction[0], rays[i].Direction[1], rays[i].Direction[2]); rd.TMax = rays[i].TMax;
        SoftwareRayQuery q;
        q.TraceRayInline(RAY_FLAG_NONE, ~0u, rd, 0);
        q.Proceed();
        TbHit& h = hits[i];
        memset(&h, 0, sizeof(h));
        if (q.CommittedStatus() == COMMITTED_TRIANGLE_HIT) {
            float2 b = q.CommittedTriangleBarycentrics();
            h.t = q.CommittedRayT(); h.b1 = b.x; h.b2 = b.y;
            h.PrimitiveIndex = q.CommittedPrimitiveIndex(); h.GeometryIndex = q.CommittedGeometryIndex(); h.InstanceIndex = q.CommittedInstanceIndex();
        } else {
            h.t = -1.0f; h.PrimitiveIndex = h.GeometryIndex = 0xffffffffu;
        }
        h.TrianglesTested = q.TrianglesTested; h.BoxesTested = q.BoxesTested;
    }
    return 0;
}
