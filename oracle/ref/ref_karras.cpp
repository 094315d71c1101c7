// ref_karras.cpp — CPU ORACLE (test infrastructure): the reference's own Karras-2012 hierarchy construction
// (D3D12RaytracingFallback/src/BuildBVHSplits.hlsli:18-131: CountLeadingZeroes … GenerateHierarchy), pre-passed from the
// mount into oracle/_ref/karras_gen.inc and compiled as host C++. Restated here: the resource bindings (`mortonCodes`,
// `hierarchyBuffer`, `Constants`) as plain arrays and the dispatch loop of main() (:133-142).
#include "hlsl_compat.h"
#include <algorithm>
#include <cstring>

namespace refcore {

struct uint2 { uint x, y; uint2() : x(0), y(0) {} uint2(uint a, uint b) : x(a), y(b) {} };
struct HierarchyNode { uint ParentIndex, LeftChildIndex, RightChildIndex; }; // RayTracingHlslCompat.h:33-38
struct { uint NumberOfElements; } static thread_local Constants;
static thread_local const uint* mortonCodes;
static thread_local HierarchyNode* hierarchyBuffer;
inline int firstbithigh(uint v) { return v ? 31 - __builtin_clz(v) : -1; } // HLSL: -1 when no bit is set
inline int clamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline uint min(uint a, int b) { return std::min(a, (uint)b); }   // min(idx, j): HLSL promotes int to uint
inline uint max(uint a, int b) { return std::max(a, (uint)b); }

#include "../_ref/karras_gen.inc"

} // namespace refcore

extern "C" __attribute__((visibility("default")))
void ref_karras(const unsigned int* codes, unsigned int n, unsigned int* out3) {
    using namespace refcore;
    Constants.NumberOfElements = n;
    mortonCodes = codes;
    hierarchyBuffer = (HierarchyNode*)out3;
    for (unsigned int i = 0; i < 2 * n - 1; i++) hierarchyBuffer[i] = HierarchyNode{0xffffffffu, 0, 0};
    for (unsigned int idx = 0; idx + 1 < n; idx++) GenerateHierarchy(idx); // main(): one thread per internal node
}
