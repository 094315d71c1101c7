// ref_treelet.cpp — CPU ORACLE (test infrastructure): the reference's own treelet optimisation — the tables and helpers of
// D3D12RaytracingFallback/src/TreeletReorderBindings.h:34, 53-112 and CalculateCost, FormTreelet, FindOptimalPartitions,
// ReformTree of TreeletReorder.hlsl:20-236 — pre-passed from the mount into oracle/_ref/treelet_gen.inc and compiled as
// host C++. The shader runs one 32-thread group per treelet with groupshared arrays and group barriers; here the group is
// 32 host threads and GroupMemoryBarrierWithGroupSync is a pthread barrier. Restated: the resource bindings as plain
// arrays, AABB (RayTracingHlslCompat.h:40-45), GetNumInternalNodes (:404-407), FLT_MAX (ShaderUtil.hlsli:22), and the
// per-group driver (main(): one FormTreelet / FindOptimalPartitions / ReformTree round for a given root; the climb to
// the parent, TraverseToParent, is the part deviation D2 replaces and is not compiled).
#include "hlsl_compat.h"
#include <pthread.h>
#include <cfloat>
#include <cstring>
#include <thread>
#include <vector>

namespace refcore {

struct AABB { float3 min, max; };
struct HierarchyNode { uint ParentIndex, LeftChildIndex, RightChildIndex; };
struct { uint NumberOfElements; uint MinTrianglesPerTreelet; } static Constants;
static HierarchyNode* hierarchyBuffer;
static AABB* AABBBuffer;
inline uint GetNumInternalNodes(uint numLeaves) { return numLeaves - 1; }
inline uint countbits(uint v) { return (uint)__builtin_popcount(v); }
inline uint firstbitlow(uint v) { return v ? (uint)__builtin_ctz(v) : 0xffffffffu; }
inline uint min(uint a, uint b) { return a < b ? a : b; }
inline uint max(uint a, uint b) { return a > b ? a : b; }
inline uint max(uint a, int b) { return max(a, (uint)b); }
static pthread_barrier_t g_barrier;
inline void GroupMemoryBarrierWithGroupSync() { pthread_barrier_wait(&g_barrier); }
#define groupshared static
#define unroll

#include "../_ref/treelet_gen.inc"

} // namespace refcore

// One optimisation round of the treelet rooted at `root` on a caller-provided hierarchy (3 words per node) and boxes
// (min, max: 6 floats per node), executed by a 32-thread group.
extern "C" __attribute__((visibility("default")))
void ref_treelet(unsigned int* H3, float* aabb6, unsigned int n, unsigned int root) {
    using namespace refcore;
    const unsigned int total = 2 * n - 1;
    std::vector<AABB> boxes(total);
    for (unsigned int i = 0; i < total; i++) {
        boxes[i].min = float3(aabb6[6 * i], aabb6[6 * i + 1], aabb6[6 * i + 2]);
        boxes[i].max = float3(aabb6[6 * i + 3], aabb6[6 * i + 4], aabb6[6 * i + 5]);
    }
    Constants.NumberOfElements = n;
    Constants.MinTrianglesPerTreelet = 7;
    hierarchyBuffer = (HierarchyNode*)H3;
    AABBBuffer = boxes.data();
    nodeIndex = root;
    pthread_barrier_init(&g_barrier, nullptr, NumThreadsInGroup);
    std::vector<std::thread> group;
    for (unsigned int t = 0; t < NumThreadsInGroup; t++)
        group.emplace_back([t] { FormTreelet(t); FindOptimalPartitions(t); ReformTree(t); });
    for (auto& th : group) th.join();
    pthread_barrier_destroy(&g_barrier);
    for (unsigned int i = 0; i < total; i++) {
        aabb6[6 * i] = boxes[i].min.x; aabb6[6 * i + 1] = boxes[i].min.y; aabb6[6 * i + 2] = boxes[i].min.z;
        aabb6[6 * i + 3] = boxes[i].max.x; aabb6[6 * i + 4] = boxes[i].max.y; aabb6[6 * i + 5] = boxes[i].max.z;
    }
}
