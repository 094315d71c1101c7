// ref_traverse_loop.cpp — CPU ORACLE (test infrastructure): the reference's own ray-query LOOP compiled from the mount —
// D3D12RaytracingFallback/src/TraverseFunction.hlsli: SoftwareRayQuery (TraceRayInline, Proceed, the Committed*
// accessors), the stack helpers, IsOpaque, Cull, TestLeafNodeIntersections and Traverse (:17-199, 314-429, 497-790), and
// the node / primitive readers of RayTracingHelper.hlsli (:24-227) — pre-passed into oracle/_ref/traverse_loop_gen.inc by
// prepass.run_traverse_loop. The three pure functions the loop calls (RayBoxTest with contraction on, the `precise`
// RayTriangleIntersect, GetRayData) are the objects ref_traverse_box.cpp / ref_traverse_rest.cpp already compile from the
// same file, linked in. Compiled with FAST_PATH=1, DISABLE_ANYHIT, DISABLE_PROCEDURAL_GEOMETRY, TracerBoy's
// configuration (RayGenCommon.h:355-362). Restated here: the resources (RWByteAddressBuffer loads on the acceleration
// structure bytes, the emulated GPU pointer), integer vector types, layout constants of RayTracingHlslCompat.h and the
// DXR built-in constants. The traversal stack is the shader's own groupshared array (TRAVERSAL_MAX_STACK_DEPTH 16 x 64
// threads, indexed by GI * 16 + top without a bound check): run as thread 0 of the group it has 1024 entries before
// it would leave the array, which no test ray reaches (the shader on the GPU would run into its neighbour's stack after
// 16; the oracle and the CUDA kernels keep 96 — DESIGN.md "unbounded traversal stack").
// tests/test_cpu_oracle.py requires oracle/traverse.cpp to match this build: hit, barycentrics, ids and both counters.
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
struct uint4 { uint x, y, z, w; };
struct int4 { int x, y, z, w; int4() : x(0), y(0), z(0), w(0) {} int4(const uint4& u) : x((int)u.x), y((int)u.y), z((int)u.z), w((int)u.w) {} };
struct float4e : float4 { // float4 with the two extra swizzles BVHReadTriangle uses
    float4e(float a, float b, float c, float d) : float4(a, b, c, d) {}
    float2 zw() const { return float2(z, w); }
};
inline float asfloat(int i) { float f; memcpy(&f, &i, 4); return f; }
inline float asfloat(uint u) { float f; memcpy(&f, &u, 4); return f; }
inline float4e asfloat(uint4 u) { return float4e(asfloat(u.x), asfloat(u.y), asfloat(u.z), asfloat(u.w)); }
inline uint asuint(float f) { uint u; memcpy(&u, &f, 4); return u; }
inline float3 make3(float a, float2 b) { return float3(a, b.x, b.y); }
struct float3x4 {};
#define row_major
#define groupshared static thread_local
#define float4 float4e
#define HLSL 1

// resources: the acceleration structure bytes behind an emulated pointer (EmulatedPointer.hlsli)
struct RWByteAddressBuffer {
    uint8_t* bytes = nullptr;
    uint Load(uint o) const { uint v; memcpy(&v, bytes + o, 4); return v; }
    uint3 Load3(uint o) const { uint3 v; memcpy(&v, bytes + o, 12); return v; }
    uint4 Load4(uint o) const { uint4 v; memcpy(&v, bytes + o, 16); return v; }
    void Store(uint o, uint v) { memcpy(bytes + o, &v, 4); }
    void Store4(uint o, uint4 v) { memcpy(bytes + o, &v, 16); }
};
struct RWByteAddressBufferPointer { RWByteAddressBuffer buffer; uint offsetInBytes; };
struct GpuVA { uint lo, hi; GpuVA(uint a = 0, uint b = 0) : lo(a), hi(b) {} };
static thread_local RWByteAddressBuffer g_bvh;
static thread_local GpuVA TopLevelAccelerationStructureGpuVA;
inline RWByteAddressBufferPointer CreateRWByteAddressBufferPointerFromGpuVA(GpuVA) { return RWByteAddressBufferPointer{g_bvh, 0u}; }
static thread_local uint GI = 0; // SoftwareRayTraceCS.hlsl:4

// RayTracingHlslCompat.h:15, 29, 175-176, 182-188, 385, 398
#define TRAVERSAL_MAX_STACK_DEPTH 16
#define SizeOfFloat 4
#define SizeOfPrimitive 40
#define OffsetToPrimitiveData 4
struct PrimitiveMetaData { uint GeometryContributionToHitGroupIndex; uint PrimitiveIndex; uint GeometryFlags; };
#define SizeOfPrimitiveMetaData (4 * 3)
#define SizeOfAABBNode (4 * 8)
#define SizeOfBVHOffsets (4 * 4)
#define SizeOfBVHMetadata 116
#define D3D12_RAYTRACING_GEOMETRY_FLAG_OPAQUE 0x1
// DXR built-ins
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
// TracerBoy's configuration of the query, RayGenCommon.h:355-362
#define FAST_PATH 1
#define DISABLE_ANYHIT
#define DISABLE_PROCEDURAL_GEOMETRY

// the pure functions compiled on their own (ref_traverse_box.cpp: contraction on; ref_traverse_rest.cpp: `precise`)
bool RayBoxTest_fused(float& resultT, float closestT, float3 rayOriginTimesRayInverseDirection, float3 rayInverseDirection, float3 boxCenter, float3 boxHalfDim);
void RayTriangleIntersect_precise(float& hitT, uint rayFlags, uint instanceFlags, float2& bary, float3 rayOrigin, float3 rayDirection,
                                  int3 swizzledIndicies, float3 shear, float3 v0, float3 v1, float3 v2);
#define RayBoxTest RayBoxTest_fused
#define RayTriangleIntersect RayTriangleIntersect_precise
// float3(a.w, b.xy) and float3(b.zw, c) of BVHReadTriangle: the second form exists in hlsl_compat.h, the first does not
inline float3 float3_from(float a, float2 b) { return float3(a, b.x, b.y); }

#include "../_ref/traverse_loop_gen.inc"

// ---- between the query and the path tracer: SharedHitGroup.h geometry fetch + RayGenCommon.h IntersectWithMaxDistance
// (prepass.run_intersect -> oracle/_ref/intersect_gen.inc). Shims: StructuredBuffer / Buffer element access, the
// built-in RayDesc (same members as SoftwareRayDesc), kernel.glsl's Ray, OutputRayStats (RayGenCommon.h:537-543 writes
// the heat map AOV; here it records the two counters).
template <typename T> struct StructuredBuffer { const T* p = nullptr; const T& operator[](uint i) const { return p[i]; } };
template <typename T> struct Buffer { const T* p = nullptr; T operator[](uint i) const { return p[i]; } };
inline uint NonUniformResourceIndex(uint i) { return i; }
typedef SoftwareRayDesc RayDesc;
struct Ray { float3 origin; float3 direction; };
static thread_local uint g_statTris, g_statBoxes;
inline void OutputRayStats(uint TrianglesTested, uint BoxesTested) { g_statTris = TrianglesTested; g_statBoxes = BoxesTested; }
#define USE_INLINE_RAYTRACING 1
#define USE_SW_RAYTRACING 1
#undef float4
#include "../_ref/intersect_gen.inc"

} // namespace refcore

// Same result record as tb_trace_rays / oracle_trace_rays (TbHit), on a reference-layout BVH (tb_get_bvh bytes).
extern "C" __attribute__((visibility("default")))
int ref_trace_rays(const uint8_t* bvh, const TbRay* rays, uint64_t n, TbHit* hits) {
    using namespace refcore;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        g_bvh.bytes = const_cast<uint8_t*>(bvh);
        GI = 0;
        SoftwareRayDesc rd;
        rd.Origin = float3(rays[i].Origin[0], rays[i].Origin[1], rays[i].Origin[2]); rd.TMin = rays[i].TMin;
        rd.Direction = float3(rays[i].Direction[0], rays[i].Direction[1], rays[i].Direction[2]); rd.TMax = rays[i].TMax;
        SoftwareRayQuery q;
        q.TraceRayInline(RAY_FLAG_NONE, ~0u, rd, 0); // RayGenCommon.h:372-379
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

// IntersectWithMaxDistance on caller-provided scene arrays: geoms become the 72-byte hit-group shader records the way
// TracerBoy.cpp:1896-1944 fills them (one pooled vertex buffer of 8-float vertices, one pooled index buffer; offsets in
// bytes). out: 12 floats per ray = t, material, normal.xyz, tangent.xyz, uv.xy, TrianglesTested, BoxesTested.
extern "C" __attribute__((visibility("default")))
int ref_intersect(const uint8_t* bvh, const TbGeometryRecord* geoms, uint32_t numGeoms, const uint32_t* indices, const TbVertex* vertices,
                  const TbRay* rays, uint64_t n, float* out12) {
    using namespace refcore;
    std::vector<HitGroupShaderRecord> table(numGeoms);
    for (uint32_t g = 0; g < numGeoms; g++) {
        memset(&table[g], 0, sizeof(HitGroupShaderRecord));
        table[g].MaterialIndex = geoms[g].MaterialIndex;
        table[g].VertexBufferIndex = 0; table[g].VertexBufferOffset = geoms[g].VertexFirst * 8u * 4u;
        table[g].IndexBufferIndex = 0; table[g].IndexBufferOffset = geoms[g].IndexFirst * 4u;
        table[g].GeometryIndex = geoms[g].GeometryIndex;
    }
    static_assert(sizeof(HitGroupShaderRecord) == 72, "HitGroupShaderRecord is 72 bytes (SharedHitGroup.h:13-23)");
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        g_bvh.bytes = const_cast<uint8_t*>(bvh);
        GI = 0;
        ShaderTable.p = table.data(); IndexBuffers[0].p = indices; VertexBuffers[0].p = (const float*)vertices;
        Ray ray;
        ray.origin = float3(rays[i].Origin[0], rays[i].Origin[1], rays[i].Origin[2]);
        ray.direction = float3(rays[i].Direction[0], rays[i].Direction[1], rays[i].Direction[2]);
        float3 normal, tangent; float2 uv; uint primitiveID = 123;
        float2 r = IntersectWithMaxDistance(ray, rays[i].TMax, normal, tangent, uv, primitiveID);
        float* o = out12 + 12 * i;
        o[0] = r.x; o[1] = r.y; o[2] = normal.x; o[3] = normal.y; o[4] = normal.z; o[5] = tangent.x; o[6] = tangent.y; o[7] = tangent.z;
        o[8] = uv.x; o[9] = uv.y; o[10] = (float)g_statTris; o[11] = (float)g_statBoxes;
        if (primitiveID != 0) o[0] = -12345.0f; // PrimitiveID = 0 always (RayGenCommon.h:405)
    }
    return 0;
}
