// ref_treelet_pass.cpp — CPU ORACLE (test infrastructure): ONE WHOLE treelet-reorder pass of the reference compiled from
// the mount — ClearBuffers.hlsl main(), FindTreelets.hlsl (bottom-up box build, per-node triangle counts, the base treelet
// roots: the first node on each leaf-to-root path with at least MinTrianglesPerTreelet triangles) and ALL of
// TreeletReorder.hlsl (FormTreelet, FindOptimalPartitions, ReformTree, TraverseToParent, main() with its 33-iteration
// climb) — pre-passed into oracle/_ref/treelet_pass_gen.inc by prepass.run_treelet_pass. ClearBuffers and FindTreelets
// have no barriers (interlocked adds only): their threads run one after another. TreeletReorder is a 32-thread group
// shader with groupshared state and group barriers: a group is 32 host threads with a real barrier, and the groups
// (one per base treelet root) run one after another, which is a legal schedule because no group ever waits for another
// one (the first to arrive at a parent leaves, the second continues). Restated: the resources and constants, struct
// Triangle / Primitive (RayTracingHlslCompat.h:90-136), integer vector types.
#define RC_TRAVERSE 1
#include "hlsl_compat.h"
#include <pthread.h>
#include <cfloat>
#include <cstring>
#include <thread>
#include <vector>

namespace refcore {

struct uint2 { uint x, y; uint2() : x(0), y(0) {} uint2(uint a, uint b) : x(a), y(b) {} uint2(float2 f) : x((uint)f.x), y((uint)f.y) {} };
struct uint3 {
    uint x, y, z;
    uint3() : x(0), y(0), z(0) {}
    uint3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
    uint3(uint a, uint2 b) : x(a), y(b.x), z(b.y) {}
    uint3(uint2 a, uint c) : x(a.x), y(a.y), z(c) {}
};
struct uint4 {
    uint x, y, z, w;
    uint3 xyz() const { return uint3(x, y, z); }
    uint2 xy() const { return uint2(x, y); }
    uint2 zw() const { return uint2(z, w); }
};
struct int4 { int x, y, z, w; int4() : x(0), y(0), z(0), w(0) {} int4(const uint4& u) : x((int)u.x), y((int)u.y), z((int)u.z), w((int)u.w) {} };
inline float asfloat(int i) { float f; memcpy(&f, &i, 4); return f; }
inline float asfloat(uint u) { float f; memcpy(&f, &u, 4); return f; }
inline float3 asfloat(uint3 u) { return float3(asfloat(u.x), asfloat(u.y), asfloat(u.z)); }
struct float4e : float4 { float4e(float a, float b, float c, float d) : float4(a, b, c, d) {} float2 zw() const { return float2(z, w); } };
inline float4e asfloat(uint4 u) { return float4e(asfloat(u.x), asfloat(u.y), asfloat(u.z), asfloat(u.w)); }
inline uint asuint(float f) { uint u; memcpy(&u, &f, 4); return u; }
#define float4 float4e
inline uint countbits(uint v) { return (uint)__builtin_popcount(v); }
inline uint firstbitlow(uint v) { return v ? (uint)__builtin_ctz(v) : 0xffffffffu; }
inline uint min(uint a, uint b) { return a < b ? a : b; }
inline uint max(uint a, uint b) { return a > b ? a : b; }
inline uint max(uint a, int b) { return max(a, (uint)b); }

struct RWByteAddressBuffer {
    uint8_t* bytes = nullptr;
    uint Load(uint o) const { uint v; memcpy(&v, bytes + o, 4); return v; }
    uint3 Load3(uint o) const { uint3 v; memcpy(&v, bytes + o, 12); return v; }
    uint4 Load4(uint o) const { uint4 v; memcpy(&v, bytes + o, 16); return v; }
    void Store(uint o, uint v) { memcpy(bytes + o, &v, 4); }
    void Store4(uint o, uint4 v) { memcpy(bytes + o, &v, 16); }
    void InterlockedAdd(uint o, uint v, uint& original) { original = Load(o); Store(o, original + v); }
};
struct RWByteAddressBufferPointer { RWByteAddressBuffer buffer; uint offsetInBytes; };
struct AABB { float3 min, max; };                                             // RayTracingHlslCompat.h:40-45
struct HierarchyNode { uint ParentIndex, LeftChildIndex, RightChildIndex; }; // :33-38
struct Triangle { float3 v0, v1, v2; };                                       // :90-95
struct Primitive { uint PrimitiveType; uint4 data0; uint4 data1; uint data2; }; // :122-136 (HLSL side), 40 bytes
struct PrimitiveMetaData { uint GeometryContributionToHitGroupIndex; uint PrimitiveIndex; uint GeometryFlags; };
struct { uint NumberOfElements; uint MinTrianglesPerTreelet; } static Constants; // TreeletReorderBindings.h:17-21
static HierarchyNode* hierarchyBuffer;
static AABB* AABBBuffer;
static Primitive* InputBuffer;
static uint* BaseTreeletsIndexBuffer;
static RWByteAddressBuffer NumTrianglesBuffer, BaseTreeletsCountBuffer;
#define TRIANGLE_TYPE 0x1
#define SizeOfFloat 4
#define SizeOfUINT32 4
#define SizeOfPrimitive 40
#define OffsetToPrimitiveData 4
#define SizeOfPrimitiveMetaData (4 * 3)
#define SizeOfBVHMetadata 116
#define SizeOfAABBNode (4 * 8)
#define SizeOfBVHOffsets (4 * 4)
inline uint GetNumInternalNodes(uint numLeaves) { return numLeaves - 1; }
inline AABB GetProceduralPrimitiveAABB(Primitive) { return AABB(); } // procedural primitives do not occur on this path
static pthread_barrier_t g_barrier;
inline void GroupMemoryBarrierWithGroupSync() { pthread_barrier_wait(&g_barrier); }
inline void DeviceMemoryBarrierWithGroupSync() { pthread_barrier_wait(&g_barrier); }
inline void DeviceMemoryBarrier() {}
#define groupshared static
static uint NullItem;                    // BitonicSortCommon.hlsli:18-21 (cbuffer): 0xffffffff = ascending (BitonicSort.cpp:87)
#define ElementsSummedPerThread 8        // CalculateSceneAABBBindings.h:20
inline float3 min(float3 a, float3 b);   // hlsl_compat.h
inline float3 max(float3 a, float3 b);

#include "../_ref/treelet_pass_gen.inc"

} // namespace refcore

// H3: HierarchyNode per node, rewritten in place; prims40: the sorted primitives; returns the number of base treelet roots.
extern "C" __attribute__((visibility("default")))
int ref_treelet_pass(uint32_t* H3, const void* prims40, uint32_t n, uint32_t minTris) {
    using namespace refcore;
    static_assert(sizeof(Primitive) == 40, "Primitive is 40 bytes");
    const uint32_t total = 2 * n - 1;
    std::vector<AABB> boxes(total);
    std::vector<uint32_t> numTris(n), baseIndex(n / 7 + 8), baseCount(1, 0);
    Constants.NumberOfElements = n; Constants.MinTrianglesPerTreelet = minTris;
    hierarchyBuffer = (HierarchyNode*)H3; AABBBuffer = boxes.data(); InputBuffer = (Primitive*)prims40;
    BaseTreeletsIndexBuffer = baseIndex.data();
    NumTrianglesBuffer.bytes = (uint8_t*)numTris.data(); BaseTreeletsCountBuffer.bytes = (uint8_t*)baseCount.data();
    for (uint32_t t = 0; t < n; t++) clear_main(uint3(t, 0, 0));
    for (uint32_t t = 0; t < n; t++) find_main(uint3(t, 0, 0));
    const uint32_t groups = baseCount[0];
    pthread_barrier_init(&g_barrier, nullptr, NumThreadsInGroup);
    std::vector<std::thread> pool;
    for (uint32_t t = 0; t < NumThreadsInGroup; t++)
        pool.emplace_back([t, groups] {
            for (uint32_t g = 0; g < groups; g++) {
                reorder_main(uint3(g, 0, 0), uint3(t, 0, 0));
                pthread_barrier_wait(&g_barrier); // one group at a time
            }
        });
    for (auto& th : pool) th.join();
    pthread_barrier_destroy(&g_barrier);
    return (int)groups;
}

// ---- the builder's front (same translation unit for its shims)
// ShouldSwap of the bitonic network, ascending as GpuBVH2Builder.cpp:316-322 asks for it
extern "C" __attribute__((visibility("default")))
int ref_should_swap(uint32_t A, uint32_t B, uint32_t indexA, uint32_t indexB) {
    refcore::NullItem = 0xffffffffu;
    return refcore::ShouldSwap(A, B, indexA, indexB) ? 1 : 0;
}
extern "C" __attribute__((visibility("default")))
void ref_centroid(const void* prim40, float* out3) {
    using namespace refcore;
    InputBuffer = (Primitive*)prim40;
    float3 c = GetCentroid(0);
    out3[0] = c.x; out3[1] = c.y; out3[2] = c.z;
}
// CalculateSceneAABB per thread (8 primitives each, CalculateSceneAABBFromPrimitives.hlsl), then the min / max of the
// per-thread boxes, which is what the further reduction passes of SceneAABBCalculator.cpp:36-84 compute
extern "C" __attribute__((visibility("default")))
void ref_scene_box(const void* prims40, uint32_t n, float* out6) {
    using namespace refcore;
    InputBuffer = (Primitive*)prims40;
    Constants.NumberOfElements = n;
    AABB total;
    total.min = float3(FLT_MAX, FLT_MAX, FLT_MAX); total.max = float3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    for (uint32_t base = 0; base < n; base += ElementsSummedPerThread) {
        AABB t = CalculateSceneAABB(base);
        total.min = min(total.min, t.min); total.max = max(total.max, t.max);
    }
    out6[0] = total.min.x; out6[1] = total.min.y; out6[2] = total.min.z; out6[3] = total.max.x; out6[4] = total.max.y; out6[5] = total.max.z;
}
