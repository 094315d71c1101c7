// ref_refit.cpp — CPU ORACLE (test infrastructure): the builder's last stage compiled from the mount —
// BottomLevelPrepareForComputeAABBs.hlsl main() (header offsets, thread -> node map), ComputeLeafAABB
// (BottomLevelComputeAABBs.hlsl), the whole ComputeAABBs.hlsli (leaf / internal node encoding, the bottom-up climb in which
// the second thread to arrive at a node continues, the smaller-subtree-left rule) and the helpers of
// RayTracingHelper.hlsli it uses — pre-passed into oracle/_ref/refit_*.inc by prepass.run_refit. The kernel has no
// barriers, only interlocked adds, so running its threads one after another is one of its legal schedules; the caller
// chooses ascending or descending thread order, the two schedules that resolve an equal-count node (deviation D1: the
// reference's result there depends on arrival order) in opposite ways. Restated: the resources (byte-address buffers, the
// structured hierarchy buffer), the constants, integer vector types and layout constants of RayTracingHlslCompat.h.
#define RC_TRAVERSE 1
#include "hlsl_compat.h"
#include <cstring>
#include <vector>

namespace refcore {

struct uint2 { uint x, y; uint2() : x(0), y(0) {} uint2(uint a, uint b) : x(a), y(b) {} uint2(float2 f) : x((uint)f.x), y((uint)f.y) {} };
struct uint3 { uint x, y, z; uint3() : x(0), y(0), z(0) {} uint3(uint a, uint b, uint c) : x(a), y(b), z(c) {} };
struct uint4 { uint x, y, z, w; };
struct int4 { int x, y, z, w; int4() : x(0), y(0), z(0), w(0) {} int4(const uint4& u) : x((int)u.x), y((int)u.y), z((int)u.z), w((int)u.w) {} };
inline float asfloat(int i) { float f; memcpy(&f, &i, 4); return f; }
inline float asfloat(uint u) { float f; memcpy(&f, &u, 4); return f; }
inline float3 asfloat(uint3 u) { return float3(asfloat(u.x), asfloat(u.y), asfloat(u.z)); }
struct float4e : float4 { float4e(float a, float b, float c, float d) : float4(a, b, c, d) {} float2 zw() const { return float2(z, w); } };
inline float4e asfloat(uint4 u) { return float4e(asfloat(u.x), asfloat(u.y), asfloat(u.z), asfloat(u.w)); }
inline uint asuint(float f) { uint u; memcpy(&u, &f, 4); return u; }
#define float4 float4e

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
struct HierarchyNode { uint ParentIndex, LeftChildIndex, RightChildIndex; }; // RayTracingHlslCompat.h:33-38
struct AABB { float3 min, max; };                                             // :40-45
struct PrimitiveMetaData { uint GeometryContributionToHitGroupIndex; uint PrimitiveIndex; uint GeometryFlags; };
struct InputConstants { uint NumberOfElements; uint UpdateFlags; };           // ConstructAABBBindings.h:18-22
#define PREPARE_UPDATE_FLAG 0x1
#define PERFORM_UPDATE_FLAG 0x2
#define ShouldPrepareUpdate (Constants.UpdateFlags & PREPARE_UPDATE_FLAG)
#define ShouldPerformUpdate (Constants.UpdateFlags & PERFORM_UPDATE_FLAG)
static RWByteAddressBuffer outputBVH, scratchMemory, childNodesProcessedCounter;
static HierarchyNode* hierarchyBuffer;
static uint* aabbParentBuffer;
static InputConstants Constants;
// RayTracingHlslCompat.h:29-31, 98, 120, 175-176, 188, 194, 385, 398, 404-407, 427-430, 440-444
#define SizeOfFloat 4
#define SizeOfUINT32 4
#define TRIANGLE_TYPE 0x1
#define SizeOfPrimitive 40
#define OffsetToPrimitiveData 4
#define SizeOfPrimitiveMetaData (4 * 3)
#define SizeOfBVHMetadata 116
#define SizeOfAABBNode (4 * 8)
#define SizeOfBVHOffsets (4 * 4)
inline uint GetNumInternalNodes(uint numLeaves) { return numLeaves - 1; }
inline uint GetOffsetFromPrimitivesToPrimitiveMetaData(uint numPrimitives) { return SizeOfPrimitive * numPrimitives; }
inline uint GetOffsetToPrimitives(uint numTriangles) { uint numAABBs = numTriangles + GetNumInternalNodes(numTriangles); return SizeOfBVHOffsets + SizeOfAABBNode * numAABBs; }

#include "../_ref/refit_compute_gen.inc"
#include "../_ref/refit_prepare_gen.inc"

} // namespace refcore

// bvh: the reference-layout buffer with sorted primitives and metadata in place (header and nodes are written here);
// H3: HierarchyNode per node; descending != 0 runs the ComputeAABBs threads n-1 .. 0 instead of 0 .. n-1.
extern "C" __attribute__((visibility("default")))
int ref_refit(uint8_t* bvh, const uint32_t* H3, uint32_t n, int descending) {
    using namespace refcore;
    std::vector<uint32_t> scratch(n), counters(n ? n : 1);
    std::vector<HierarchyNode> H((const HierarchyNode*)H3, (const HierarchyNode*)H3 + (2 * (size_t)n - 1));
    outputBVH.bytes = bvh; scratchMemory.bytes = (uint8_t*)scratch.data(); childNodesProcessedCounter.bytes = (uint8_t*)counters.data();
    hierarchyBuffer = H.data(); aabbParentBuffer = nullptr;
    Constants.NumberOfElements = n; Constants.UpdateFlags = 0;
    for (uint32_t t = 0; t < n; t++) prepare_main(uint3(t, 0, 0));
    for (uint32_t k = 0; k < n; k++) compute_main(uint3(descending ? n - 1 - k : k, 0, 0));
    return 0;
}

// The same kernel with PERFORM_UPDATE (GpuBVH2Builder.cpp:165-234): bvh is a finished acceleration structure whose
// sorted primitives have been replaced by the moved ones; child indices come from the stored node flags and parents
// from `parents` (the aabbParentBuffer a PREPARE_UPDATE build leaves behind: parents[child] = node).
extern "C" __attribute__((visibility("default")))
int ref_refit_update(uint8_t* bvh, const uint32_t* parents, uint32_t n, int descending) {
    using namespace refcore;
    std::vector<uint32_t> scratch(n), counters(n ? n : 1);
    std::vector<uint32_t> par(parents, parents + (2 * (size_t)n - 1));
    outputBVH.bytes = bvh; scratchMemory.bytes = (uint8_t*)scratch.data(); childNodesProcessedCounter.bytes = (uint8_t*)counters.data();
    hierarchyBuffer = nullptr; aabbParentBuffer = par.data();
    Constants.NumberOfElements = n; Constants.UpdateFlags = PERFORM_UPDATE_FLAG;
    for (uint32_t t = 0; t < n; t++) prepare_main(uint3(t, 0, 0));
    for (uint32_t k = 0; k < n; k++) compute_main(uint3(descending ? n - 1 - k : k, 0, 0));
    return 0;
}
