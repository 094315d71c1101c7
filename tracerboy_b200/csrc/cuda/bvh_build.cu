// bvh_build.cu — GPU LBVH builder for sm_100a.
//
// Role: the D3D12 Raytracing Fallback Layer's GpuBvh2Builder::BuildBVH
// (D3D12RaytracingFallback/src/GpuBVH2Builder.cpp:167-356) and the ~20 HLSL kernels it
// dispatches. Output is byte-identical to the reference layout
// (RayTracingHlslCompat.h:344-398) for a deterministic child-order rule; a second,
// traversal-friendly copy (PairNode/WideTri, device_types.h) is derived from it.
//
// Pipeline (one stream, no host sync inside):
//   load_prims      BottomLevelLoadTriangles.hlsli:88-126 + scene AABB (CalculateSceneAABB*.hlsl)
//   morton          CalculateMortonCodesBindings.h:117-162 (30 bit, y,x,z interleave)
//   radix sort      hand-written stable LSD radix sort (4 x 8 bit) in place of the O(N log^2 N) bitonic
//                   sort (BitonicSort.cpp:67-143); stability reproduces the reference's
//                   tie-break-by-index compare (BitonicSortCommon.hlsli:37-47)
//   rearrange       RearrangeTriangles.hlsl:29-36
//   karras          BuildBVHSplits.hlsli:34-141
//   treelet x3      ClearBuffers.hlsl / FindTreelets.hlsl / TreeletReorder.hlsl (1 warp / treelet)
//   refit           ComputeAABBs.hlsli:69-172 (centre/half boxes, smaller subtree left)
//   widen           reference nodes -> PairNode/WideTri
//
// Determinism: every cross-thread hand-off is "second arrival continues and recomputes
// from both children", so arrival order never selects data. Two documented deviations
// from the reference, shared with the oracle: equal-size siblings keep Karras order
// (swap iff leftCount > rightCount), and treelet climbing is not capped at 33 levels.
#include <cfloat>
#include <cstdio>
#include "../common/tb_vec.h"
#include "device_types.h"
#include "launch.h"

using namespace tbm;

namespace tbd {

namespace {

struct Prim { uint32_t type; float v[9]; };
struct Meta { uint32_t geom, prim, flags; };
struct Box { f3 mn, mx; };

__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// ---------------------------------------------------------------- load + AABB
// BottomLevelLoadTriangles.hlsli:14-126 in its three index-format variants (none / 16 bit / 32 bit), with the optional
// 3x4 transform (TransformVertex :83-86: mul(float3x4, float4(v, 1)), one dot product per row, left to right) and a
// vertex stride, straight from the caller's device buffers (D3D12_RAYTRACING_GEOMETRY_TRIANGLES_DESC).
__device__ __forceinline__ void load_triangle(const BuildGeometry& G, uint32_t t, Prim& p) {
    p.type = 1;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        uint32_t vi = 3 * t + k;
        if (G.indexFormat == 4) vi = ((const uint32_t*)G.indices)[vi];
        else if (G.indexFormat == 2) vi = ((const uint16_t*)G.indices)[vi];
        const float* pv = (const float*)(G.positions + (size_t)vi * G.strideBytes);
        float x = pv[0], y = pv[1], z = pv[2];
        if (G.transform) {
            const float* m = G.transform;
            float tx = ((m[0] * x + m[1] * y) + m[2] * z) + m[3];
            float ty = ((m[4] * x + m[5] * y) + m[6] * z) + m[7];
            float tz = ((m[8] * x + m[9] * y) + m[10] * z) + m[11];
            x = tx; y = ty; z = tz;
        }
        p.v[3 * k] = x; p.v[3 * k + 1] = y; p.v[3 * k + 2] = z;
    }
}

__global__ void k_load_prims(const BuildGeometry* __restrict__ geoms, const uint32_t* __restrict__ triPrefix,
                             uint32_t numGeoms, uint32_t n, Prim* __restrict__ prims,
                             Meta* __restrict__ meta, uint32_t* __restrict__ sceneBox /*6 ordered uints*/) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    f3 mn = mk3(FLT_MAX), mx = mk3(-FLT_MAX);
    if (i < n) {
        // geometry lookup: last g with triPrefix[g] <= i
        uint32_t lo = 0, hi = numGeoms;
        while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (triPrefix[mid] <= i) lo = mid; else hi = mid; }
        const BuildGeometry G = geoms[lo];
        uint32_t t = i - triPrefix[lo];
        Prim p;
        load_triangle(G, t, p);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            mn = min3(mn, mk3(p.v[3 * k], p.v[3 * k + 1], p.v[3 * k + 2]));
            mx = max3(mx, mk3(p.v[3 * k], p.v[3 * k + 1], p.v[3 * k + 2]));
        }
        uint32_t* dst = (uint32_t*)(prims + i);
        const uint32_t* src = (const uint32_t*)&p;
#pragma unroll
        for (int k = 0; k < 10; k++) dst[k] = src[k];
        Meta m = {lo, t, G.flags};
        meta[i] = m;
    }
    // warp reduce, block reduce, then 6 atomics per block (min/max are order independent => deterministic). One set of
    // atomics per warp serialised 3.9 M same-address operations at 20 M triangles (2.8 ms for a 0.3 ms copy).
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn.x = fminf(mn.x, __shfl_xor_sync(0xffffffffu, mn.x, o)); mn.y = fminf(mn.y, __shfl_xor_sync(0xffffffffu, mn.y, o)); mn.z = fminf(mn.z, __shfl_xor_sync(0xffffffffu, mn.z, o));
        mx.x = fmaxf(mx.x, __shfl_xor_sync(0xffffffffu, mx.x, o)); mx.y = fmaxf(mx.y, __shfl_xor_sync(0xffffffffu, mx.y, o)); mx.z = fmaxf(mx.z, __shfl_xor_sync(0xffffffffu, mx.z, o));
    }
    __shared__ float s_red[32][6];
    const uint32_t warp = threadIdx.x >> 5, warps = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) {
        const bool valid = mn.x <= mx.x; // a warp without triangles (or with NaN only) contributes nothing
        s_red[warp][0] = valid ? mn.x : FLT_MAX; s_red[warp][1] = valid ? mn.y : FLT_MAX; s_red[warp][2] = valid ? mn.z : FLT_MAX;
        s_red[warp][3] = valid ? mx.x : -FLT_MAX; s_red[warp][4] = valid ? mx.y : -FLT_MAX; s_red[warp][5] = valid ? mx.z : -FLT_MAX;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        const bool isMin = threadIdx.x < 3;
        float r = s_red[0][threadIdx.x];
        for (uint32_t k = 1; k < warps; k++) r = isMin ? fminf(r, s_red[k][threadIdx.x]) : fmaxf(r, s_red[k][threadIdx.x]);
        float bmn = s_red[0][0], bmx = s_red[0][3];
        for (uint32_t k = 1; k < warps; k++) { bmn = fminf(bmn, s_red[k][0]); bmx = fmaxf(bmx, s_red[k][3]); }
        if (bmn <= bmx) {
            if (isMin) atomicMin(&sceneBox[threadIdx.x], float_to_ordered(r));
            else atomicMax(&sceneBox[threadIdx.x], float_to_ordered(r));
        }
    }
}

__global__ void k_morton(const Prim* __restrict__ prims, uint32_t n, const uint32_t* __restrict__ sceneBox,
                         uint32_t* __restrict__ codes, uint32_t* __restrict__ order) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    f3 smin = mk3(ordered_to_float(sceneBox[0]), ordered_to_float(sceneBox[1]), ordered_to_float(sceneBox[2]));
    f3 smax = mk3(ordered_to_float(sceneBox[3]), ordered_to_float(sceneBox[4]), ordered_to_float(sceneBox[5]));
    const float* v = prims[i].v;
    f3 c = ((mk3(v[0], v[1], v[2]) + mk3(v[3], v[4], v[5])) + mk3(v[6], v[7], v[8])) / 3.0f;
    f3 dim = max3(smax - smin, mk3(0.00001f));
    f3 unit = (c - smin) / dim;
    f3 adj = min3(max3(unit * 1024.0f, mk3(0.0f)), mk3(1023.0f));
    uint32_t coords[3] = {(uint32_t)adj.y, (uint32_t)adj.x, (uint32_t)adj.z};
    uint32_t code = 0;
#pragma unroll
    for (uint32_t bit = 0; bit < 10; bit++)
#pragma unroll
        for (uint32_t axis = 0; axis < 3; axis++)
            if ((1u << bit) & coords[axis]) code |= 1u << (bit * 3 + axis);
    codes[i] = code;
    order[i] = i;
}

__global__ void k_rearrange(const Prim* __restrict__ prims, const Meta* __restrict__ meta,
                            const uint32_t* __restrict__ order, uint32_t n, Prim* __restrict__ outPrims,
                            Meta* __restrict__ outMeta) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s = order[i];
    const uint32_t* src = (const uint32_t*)(prims + s);
    uint32_t* dst = (uint32_t*)(outPrims + i);
#pragma unroll
    for (int k = 0; k < 10; k++) dst[k] = src[k];
    outMeta[i] = meta[s];
}

// --------------------------------------------------------------- radix sort
// Stable LSD radix sort of (morton code, primitive index) pairs, 8 bits per pass, 4 passes over the
// 30-bit codes. Replaces the reference's bitonic sort (BitonicSort.cpp:67-143, O(N log^2 N) passes over
// global memory); stability gives the same order as its tie-break-by-index compare
// (BitonicSortCommon.hlsli:37-47). Three kernels per pass: per-block digit histograms, one exclusive
// scan over (digit, block), and an order-preserving scatter that ranks items with warp match/ballot.
#define RS_THREADS 256
#define RS_ITEMS 16
#define RS_TILE (RS_THREADS * RS_ITEMS)

__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const uint32_t* __restrict__ keys, uint32_t n, int shift, uint32_t numBlocks,
                                                           uint32_t* __restrict__ hist /* [256][numBlocks] */) {
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll
    for (int c = 0; c < RS_ITEMS; c++) {
        uint32_t i = base + c * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&sh[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * numBlocks + blockIdx.x] = sh[threadIdx.x];
}

// Exclusive scan of hist[256][numBlocks] in (digit, block) order, two small kernels: each block scans one
// digit row (coalesced, running carry) and records the row total; then 256 row totals are scanned.
__global__ void __launch_bounds__(256) k_radix_scan_rows(uint32_t* __restrict__ hist, uint32_t numBlocks, uint32_t* __restrict__ rowTotal) {
    __shared__ uint32_t warpSum[8];
    __shared__ uint32_t carry;
    uint32_t* row = hist + (size_t)blockIdx.x * numBlocks;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < numBlocks; base += 256) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < numBlocks ? row[i] : 0u;
        uint32_t incl = v;
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += t; }
        if (lane == 31) warpSum[warp] = incl;
        __syncthreads();
        uint32_t before = 0;
        for (uint32_t w = 0; w < warp; w++) before += warpSum[w];
        uint32_t c = carry;
        if (i < numBlocks) row[i] = c + before + incl - v;
        __syncthreads();
        if (threadIdx.x == 255) carry = c + before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) rowTotal[blockIdx.x] = carry;
}
__global__ void __launch_bounds__(256) k_radix_scan_digits(uint32_t* __restrict__ rowTotal) {
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = rowTotal[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t run = 0; for (int d = 0; d < 256; d++) { uint32_t t = sh[d]; sh[d] = run; run += t; } }
    __syncthreads();
    rowTotal[threadIdx.x] = sh[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) k_radix_scatter(const uint32_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn, uint32_t n,
                                                              int shift, uint32_t numBlocks, const uint32_t* __restrict__ offsets /* row-scanned hist */,
                                                              const uint32_t* __restrict__ digitStart, uint32_t* __restrict__ keysOut, uint32_t* __restrict__ valsOut) {
    __shared__ uint32_t digitBase[256];                 // global position of the next item of each digit from this block
    __shared__ uint32_t warpCount[RS_THREADS / 32][256]; // per chunk: items of each digit in each warp, then exclusive prefix over warps
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    digitBase[threadIdx.x] = digitStart[threadIdx.x] + offsets[(size_t)threadIdx.x * numBlocks + blockIdx.x];
    const uint32_t base = blockIdx.x * RS_TILE;
    for (int c = 0; c < RS_ITEMS; c++) {
        for (int w = 0; w < RS_THREADS / 32; w++) warpCount[w][threadIdx.x] = 0;
        __syncthreads();
        uint32_t i = base + c * RS_THREADS + threadIdx.x;
        bool valid = i < n;
        uint32_t key = valid ? keysIn[i] : 0u, val = valid ? valsIn[i] : 0u;
        uint32_t digit = valid ? ((key >> shift) & 255u) : 256u; // invalid items form their own group
        uint32_t peers = __match_any_sync(0xffffffffu, digit);
        uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) warpCount[warp][digit] = __popc(peers);
        __syncthreads();
        { // thread d: exclusive prefix over the warps for digit d, then advance the block's base
            uint32_t run = 0;
            for (int w = 0; w < RS_THREADS / 32; w++) { uint32_t t = warpCount[w][threadIdx.x]; warpCount[w][threadIdx.x] = run; run += t; }
            __syncthreads();
            if (valid) {
                uint32_t pos = digitBase[digit] + warpCount[warp][digit] + rank;
                keysOut[pos] = key;
                valsOut[pos] = val;
            }
            __syncthreads();
            digitBase[threadIdx.x] += run;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------- Karras
__device__ __forceinline__ int lcp(const uint32_t* __restrict__ codes, uint32_t n, uint32_t a, uint32_t b) {
    if (a >= n || b >= n) return -1;
    uint32_t ca = codes[a], cb = codes[b];
    if (ca != cb) return __clz(ca ^ cb);
    return __clz(a ^ b) + 31;
}
__global__ void k_karras(const uint32_t* __restrict__ codes, uint32_t n, HNode* __restrict__ H) {
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n - 1) return;
    int d = lcp(codes, n, idx, idx + 1) - lcp(codes, n, idx, idx - 1);
    d = d < -1 ? -1 : (d > 1 ? 1 : d);
    int minPrefix = lcp(codes, n, idx, idx - d);
    int maxLength = 2;
    while (lcp(codes, n, idx, idx + (uint32_t)(maxLength * d)) > minPrefix) maxLength *= 4;
    int length = 0;
    for (int t = maxLength / 2; t > 0; t /= 2)
        if (lcp(codes, n, idx, idx + (uint32_t)((length + t) * d)) > minPrefix) length += t;
    uint32_t j = idx + (uint32_t)(length * d);
    uint32_t first = min(idx, j), last = max(idx, j);
    int common = lcp(codes, n, first, last);
    int sp = (int)first, step = (int)(last - first);
    do {
        step = (step + 1) >> 1;
        int ns = sp + step;
        if ((uint32_t)ns < last && lcp(codes, n, first, (uint32_t)ns) > common) sp = ns;
    } while (step > 1);
    uint32_t split = (uint32_t)sp, nInternal = n - 1;
    uint32_t a = (split == first) ? nInternal + split : split;
    uint32_t b = (split + 1 == last) ? nInternal + split + 1 : split + 1;
    H[idx].left = a;
    H[idx].right = b;
    H[a].parent = idx;
    H[b].parent = idx;
    if (idx == 0) H[0].parent = 0xffffffffu;
}

// ------------------------------------------------------------------ treelets
__device__ __forceinline__ float surface_area(const Box& b) {
    f3 d = b.mx - b.mn;
    return 2.0f * ((d.x * d.y + d.x * d.z) + d.y * d.z);
}
__device__ __forceinline__ Box combine(const Box& a, const Box& b) { Box r; r.mn = min3(a.mn, b.mn); r.mx = max3(a.mx, b.mx); return r; }
__device__ __forceinline__ void leaf_box(const Prim* prims, uint32_t i, f3& c, f3& h) {
    const float* v = prims[i].v;
    f3 v0 = mk3(v[0], v[1], v[2]), v1 = mk3(v[3], v[4], v[5]), v2 = mk3(v[6], v[7], v[8]);
    f3 mn = min3(min3(v0, v1), v2), mx = max3(max3(v0, v1), v2);
    mn = min3(mn, mx - 0.001f);
    c = (mn + mx) * 0.5f;
    h = mx - c;
}
// L2-coherent accessors for data handed between thread blocks (L1 is not coherent). A scratch box is
// 8 floats (min.xyz, pad, max.xyz, pad) so that it moves as two 128-bit transactions instead of six scalar ones.
__device__ __forceinline__ Box ld_box(const float* aabb, uint32_t i) {
    const float4* p = (const float4*)(aabb + 8 * (size_t)i);
    const float4 lo = __ldcg(p), hi = __ldcg(p + 1);
    Box b;
    b.mn = mk3(lo.x, lo.y, lo.z);
    b.mx = mk3(hi.x, hi.y, hi.z);
    return b;
}
__device__ __forceinline__ void st_box(float* aabb, uint32_t i, const Box& b) {
    float4* p = (float4*)(aabb + 8 * (size_t)i);
    __stcg(p, make_float4(b.mn.x, b.mn.y, b.mn.z, 0.0f));
    __stcg(p + 1, make_float4(b.mx.x, b.mx.y, b.mx.z, 0.0f));
}
__device__ __forceinline__ uint32_t ld_u(const uint32_t* p) { return __ldcg(p); }

__global__ void k_treelet_clear(uint32_t* numTris, uint32_t nInternal, uint32_t* baseCount) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *baseCount = 0;
    if (i < nInternal) numTris[i] = 0;
}

// FindTreelets.hlsl:31-88
__global__ void k_find_treelets(const Prim* __restrict__ prims, uint32_t n, uint32_t minTris, HNode* H, float* aabb,
                                uint32_t* numTris, uint32_t* baseCount, uint32_t* baseRoots) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n) return;
    const uint32_t nInternal = n - 1, total = 2 * n - 1;
    uint32_t node = total - tid - 1;
    uint32_t count = 1;
    bool isLeaf = true;
    while (true) {
        Box b;
        if (isLeaf) {
            f3 c, h;
            leaf_box(prims, node - nInternal, c, h);
            b.mn = c - h; b.mx = c + h;
        } else {
            b = combine(ld_box(aabb, ld_u(&H[node].left)), ld_box(aabb, ld_u(&H[node].right)));
        }
        st_box(aabb, node, b);
        __threadfence();
        if (count >= minTris) {
            uint32_t slot = atomicAdd(baseCount, 1u);
            baseRoots[slot] = node;
            return;
        }
        uint32_t parent = ld_u(&H[node].parent);
        uint32_t other = atomicAdd(&numTris[parent], count);
        if (other == 0) return;
        __threadfence();
        node = parent;
        count += other;
        isLeaf = false;
    }
}

// masks 0..127 ordered by popcount (sizes 2..7 used), built at compile time: the tables are part of the module image,
// so every device that loads the module has them (a run-time cudaMemcpyToSymbol only writes the current device's copy).
struct MaskTables { uint8_t bySize[128]; uint8_t sizeStart[9]; };
constexpr MaskTables make_mask_tables() {
    MaskTables t{};
    uint32_t k = 0;
    for (uint32_t s = 0; s <= 7; s++) {
        t.sizeStart[s] = (uint8_t)k;
        for (uint32_t m = 0; m < 128; m++) {
            uint32_t bits = 0;
            for (uint32_t b = 0; b < 7; b++) bits += (m >> b) & 1u;
            if (bits == s) t.bySize[k++] = (uint8_t)m;
        }
    }
    t.sizeStart[8] = (uint8_t)k;
    return t;
}
__constant__ const MaskTables c_maskTables = make_mask_tables();
#define c_masksBySize c_maskTables.bySize
#define c_sizeStart c_maskTables.sizeStart

// TreeletReorder.hlsl:38-312 — one OCTET (8 lanes) per base treelet root, climbing to the BVH root; the four
// octets of a warp run in lock step, each on its own treelet.
//
// A treelet has 7 leaves and 6 internal nodes, so 8 lanes hold it; a whole warp per treelet (the first version)
// left three quarters of every instruction idle and, more importantly, only one treelet per warp in flight on a
// pass that is a chain of dependent L2 / DRAM round trips (every word handed between warps goes through ld.cg /
// st.cg). The kernel is organised around few round trips per treelet and many treelets in flight per SM
// (profiles/: the first version spent ~30 round trips per treelet and 83 % of the 20 M-triangle build):
//  * every lane that holds a treelet leaf keeps that node's box AND its two children in registers, loaded in
//    one round trip when the lane receives the node; expanding the largest leaf is then shuffles only;
//  * the record of the root's parent (needed for the climb) is requested before FormTreelet starts;
//  * FindOptimalPartitions: subsets are spread over the 8 lanes; the one 7-leaf subset, whose 63 partitions
//    would be a serial tail, is searched by all 8 lanes with a lexicographic (cost, sequence index) argmin,
//    which is the reference's first-lowest-wins rule;
//  * ReformTree: lane 0 derives the topology from the partition table (registers + shared memory only, same
//    allocation order as the reference's stack walk), lanes 0..5 then write one internal node each; a new
//    node's box is the union of the treelet leaves in its subset, which is bit-identical to the reference's
//    bottom-up recombination because min/max are exact and order-independent;
//  * the climb carries the parent's record, box and triangle count into the next iteration in registers.
// Treelets are independent unless one is an ancestor of the other, and an ancestor only starts after both of
// its children have arrived (atomic counter), so any schedule produces the reference's result.
#define TL_OCTETS 16 // per 128-thread block
#ifndef TL_MIN_BLOCKS
#define TL_MIN_BLOCKS 8 // 9 / 10 / 12 blocks (56 / 48 / 40 registers, 40-156 B of spills) build no faster: profiles/r2_treelet_occupancy_sweep.log
#endif
__global__ void __launch_bounds__(128, TL_MIN_BLOCKS) k_treelet_reorder(uint32_t n, HNode* H, float* aabb, uint32_t* numTris,
                                                         const uint32_t* baseCount, const uint32_t* baseRoots) {
    const uint32_t lane = threadIdx.x & 31, ol = lane & 7u;
    const uint32_t oct = threadIdx.x >> 3; // octet within the block
    const uint32_t nInternal = n - 1, total = 2 * n - 1;
    const uint32_t FULL = 0xffffffffu;
    // Cost and partition tables of the four octets of a warp are interleaved ([subset][octet]): the octets run in lock step
    // on the same subset and partition indices, so side-by-side tables ([octet][subset]) put every access of the four on
    // the same bank (3.2 G bank conflicts in the 20 M-triangle pass, profiles/r1c_treelet_octets_20m_ncu_full_raw.csv).
    __shared__ float s_cost[TL_OCTETS / 4][128][4];
    __shared__ uint8_t s_part[TL_OCTETS / 4][128][4];
    __shared__ float s_box[TL_OCTETS][7][6];
    __shared__ uint32_t s_new[TL_OCTETS][6][3]; // per new internal node: leaf subset, left child code, right child code
    float (*cost4)[4] = s_cost[oct >> 2];
    uint8_t (*part4)[4] = s_part[oct >> 2];
    const uint32_t o4 = oct & 3u;
    // Replayed offline over the exact access sequence of the partition search: 4.77 wavefronts per load for side-by-side
    // tables, 1.96 interleaved. Swizzling the subset index as well (m ^ m >> 4: 1.22) was measured SLOWER (37.6 -> 40.8 ms at
    // 20.8 M triangles, profiles/r2_bvh_build_ms_interleaved_cost_tables.log): with the conflicts between octets gone the
    // kernel is bound by issue slots, and the swizzle costs two more instructions per table access.
#define cost(m) cost4[m][o4]
#define part(m) part4[m][o4]
    const uint32_t numRoots = *baseCount;
    const uint32_t stride = gridDim.x * TL_OCTETS;
    uint32_t w = blockIdx.x * TL_OCTETS + oct; // next base root of this octet
    bool have = false;                          // this octet is working on a treelet
    uint32_t root = 0, rl = 0, rr = 0, rparent = 0, mine = 0;
    Box rb; rb.mn = mk3(0.0f); rb.mx = mk3(0.0f);
    while (true) {
        if (!have && w < numRoots) {
            // the root's record; the lanes of the octet read the same words (one broadcast transaction each)
            root = baseRoots[w];
            w += stride;
            rl = ld_u(&H[root].left); rr = ld_u(&H[root].right); rparent = ld_u(&H[root].parent);
            mine = ld_u(&numTris[root]);
            rb = ld_box(aabb, root);
            have = true;
        }
        if (!__any_sync(FULL, have)) break;
        // The parent's record is only needed by the climb: requested now, consumed after the treelet is done.
        // Nobody rewrites it before this octet (or its sibling's) has climbed there.
        uint32_t pl = 0, pr = 0, pparent = 0;
        if (have && root != 0) { pl = ld_u(&H[rparent].left); pr = ld_u(&H[rparent].right); pparent = ld_u(&H[rparent].parent); }
        // ---- FormTreelet: lanes 0..6 hold the treelet leaves (id, box, children), lanes 0..5 the internal nodes
        uint32_t leaf = 0xffffffffu, internal = 0xffffffffu, cl = 0xffffffffu, cr = 0xffffffffu;
        Box lb; lb.mn = mk3(FLT_MAX); lb.mx = mk3(-FLT_MAX);
        if (ol == 0) { leaf = rl; internal = root; }
        if (ol == 1) leaf = rr;
        if (have && ol < 2) {
            lb = ld_box(aabb, leaf);
            if (leaf < nInternal) { cl = ld_u(&H[leaf].left); cr = ld_u(&H[leaf].right); }
        }
        // No expandable leaf (every candidate has a zero-area or NaN box): the reference then follows whatever the
        // node array holds for a leaf; here the treelet is left as it is.
        bool degenerate = false;
        for (uint32_t size = 2; size < 7; size++) {
            float sa = 0.0f;
            if (ol < size && leaf < nInternal) sa = surface_area(lb);
            // argmax, first index wins on ties, must be strictly > 0
            float best = sa; uint32_t bestLane = ol;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                float os = __shfl_xor_sync(FULL, best, o, 8);
                uint32_t bl = __shfl_xor_sync(FULL, bestLane, o, 8);
                if (os > best || (os == best && bl < bestLane)) { best = os; bestLane = bl; }
            }
            bestLane = __shfl_sync(FULL, bestLane, 0, 8);
            const uint32_t pick = __shfl_sync(FULL, leaf, bestLane, 8);
            const uint32_t pcl = __shfl_sync(FULL, cl, bestLane, 8), pcr = __shfl_sync(FULL, cr, bestLane, 8);
            if (pick >= nInternal) degenerate = true;
            if (have && !degenerate) {
                bool fresh = false;
                if (ol == bestLane) { leaf = pcl; fresh = true; }
                if (ol == size) { leaf = pcr; fresh = true; }
                if (ol == size - 1) internal = pick;
                if (fresh) {
                    cl = cr = 0xffffffffu;
                    if (leaf < total) lb = ld_box(aabb, leaf);
                    if (size < 6 && leaf < nInternal) { cl = ld_u(&H[leaf].left); cr = ld_u(&H[leaf].right); }
                }
            }
        }
        const bool work = have && !degenerate;
        if (ol < 7) {
            float* sb = s_box[oct][ol];
            sb[0] = lb.mn.x; sb[1] = lb.mn.y; sb[2] = lb.mn.z; sb[3] = lb.mx.x; sb[4] = lb.mx.y; sb[5] = lb.mx.z;
        }
        __syncwarp();
        // ---- FindOptimalPartitions
        const float rootSA = surface_area(rb);
        {
            // Surface area of every leaf subset: lane ol takes the 16 subsets whose leaves 0..2 are the bits of ol (the
            // eight lanes of an octet then store to eight different banks). Leaves 5, 6 and the union of the lane's low
            // leaves are held in registers, so a subset costs a few register min/max pairs instead of 6 shared-memory
            // loads per member (this loop was the larger half of the kernel's shared-memory traffic, which bounds it).
            // min / max are exact, so the order in which a subset's boxes are united does not matter.
            auto box_at = [&](uint32_t i) {
                const float* sb = s_box[oct][i];
                Box t; t.mn = mk3(sb[0], sb[1], sb[2]); t.mx = mk3(sb[3], sb[4], sb[5]);
                return t;
            };
            Box base; base.mn = mk3(FLT_MAX); base.mx = mk3(-FLT_MAX);
#pragma unroll
            for (uint32_t i = 0; i < 3; i++)
                if ((ol >> i) & 1u) base = combine(base, box_at(i));
            const Box L5 = box_at(5), L6 = box_at(6);
#pragma unroll
            for (uint32_t mid = 0; mid < 4; mid++) { // leaves 3 and 4 (re-read per group: registers are the scarce resource)
                Box g = base;
                if (mid & 1u) g = combine(g, box_at(3));
                if (mid & 2u) g = combine(g, box_at(4));
#pragma unroll
                for (uint32_t top = 0; top < 4; top++) {
                    Box b = g;
                    if (top & 1u) b = combine(b, L5);
                    if (top & 2u) b = combine(b, L6);
                    const uint32_t mask = ol + mid * 8 + top * 32;
                    cost(mask) = mask == 0 ? 0.0f : surface_area(b);
                }
            }
        }
        __syncwarp();
        if (ol < 7) cost(1u << ol) = 1.0f * surface_area(lb) / rootSA;
        __syncwarp();
        for (uint32_t sz = 2; sz <= 6; sz++) {
            const uint32_t count = (1u << (sz - 1)) - 1u; // partitions of a subset of sz leaves (its lowest leaf stays right)
            for (uint32_t k = c_sizeStart[sz] + ol; k < c_sizeStart[sz + 1]; k += 8) {
                const uint32_t mask = c_masksBySize[k];
                float lowest = FLT_MAX;
                uint32_t bestP = 0;
                const uint32_t delta = (mask - 1) & mask;
                uint32_t p = (0u - delta) & mask;
                // the reference's do/while over p, with a known trip count so that the loads of several
                // iterations are in flight together (the loop is shared-memory latency bound)
#pragma unroll 4
                for (uint32_t j = 0; j < count; j++) {
                    const float c = cost(p) + cost(mask ^ p);
                    if (c < lowest) { lowest = c; bestP = p; }
                    p = (p - delta) & mask;
                }
                cost(mask) = 1.0f * cost(mask) + lowest;
                part(mask) = (uint8_t)bestP;
            }
            __syncwarp();
        }
        {
            // all seven leaves: the reference's sequence p = 2, 4, ..., 126 (delta = 126), the first lowest wins.
            // Lane ol takes the sequence positions j = ol, ol + 8, ...; p_j = 2 (j + 1).
            float lowest = FLT_MAX;
            uint32_t bestJ = 0xffffffffu;
            for (uint32_t j = ol; j < 63; j += 8) {
                const uint32_t p = 2u * (j + 1u);
                float c = cost(p) + cost(127u ^ p);
                if (c < lowest) { lowest = c; bestJ = j; }
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                float oc = __shfl_xor_sync(FULL, lowest, o, 8);
                uint32_t oj = __shfl_xor_sync(FULL, bestJ, o, 8);
                if (oc < lowest || (oc == lowest && oj < bestJ)) { lowest = oc; bestJ = oj; }
            }
            // bestJ == 0xffffffff: no partition was below FLT_MAX (NaN costs); the serial loop then leaves bestP = 0
            if (ol == 0) part(127) = bestJ == 0xffffffffu ? (uint8_t)0 : (uint8_t)(2u * (bestJ + 1u));
        }
        __syncwarp();
        // ---- ReformTree. Lane 0 walks the partition table exactly like the reference's stack loop (pop an entry,
        // allocate the left then the right composite child, push in that order). A stack entry is
        // (internal slot << 7 | subset), 10 bits, kept in one 64-bit register. A child code is an internal slot
        // (0..5) or 8 + treelet leaf index.
        if (ol == 0) {
            unsigned long long stk = 127ull; // slot 0, all seven leaves
            uint32_t sp = 1, allocated = 1;
            while (sp > 0) {
                --sp;
                const uint32_t e = (uint32_t)(stk >> (10 * sp)) & 1023u;
                stk &= ~(1023ull << (10 * sp));
                const uint32_t em = e & 127u, es = e >> 7;
                uint32_t lm = part(em);
                if (lm == 0 || (lm & ~em) != 0 || lm == em) lm = em & (0u - em); // table not meaningful (NaN costs, idle octet): stay in bounds
                const uint32_t rm = em ^ lm;
                uint32_t lcode, rcode;
                if (__popc(lm) > 1) { lcode = allocated++; stk |= (unsigned long long)((lcode << 7) | lm) << (10 * sp); sp++; }
                else lcode = 8u + (uint32_t)(__ffs(lm) - 1);
                if (__popc(rm) > 1) { rcode = allocated++; stk |= (unsigned long long)((rcode << 7) | rm) << (10 * sp); sp++; }
                else rcode = 8u + (uint32_t)(__ffs(rm) - 1);
                s_new[oct][es][0] = em; s_new[oct][es][1] = lcode; s_new[oct][es][2] = rcode;
            }
        }
        __syncwarp();
        {
            uint32_t em = 0, lcode = 0, rcode = 0;
            if (ol < 6) { em = s_new[oct][ol][0]; lcode = s_new[oct][ol][1]; rcode = s_new[oct][ol][2]; }
            // node ids of the children: internal slot k lives in lane k's `internal`, treelet leaf i in lane i's `leaf`
            const uint32_t lInt = __shfl_sync(FULL, internal, lcode & 7u, 8), lLeaf = __shfl_sync(FULL, leaf, lcode & 7u, 8);
            const uint32_t rInt = __shfl_sync(FULL, internal, rcode & 7u, 8), rLeaf = __shfl_sync(FULL, leaf, rcode & 7u, 8);
            if (work && ol < 6) {
                const uint32_t ln = lcode >= 8u ? lLeaf : lInt, rn = rcode >= 8u ? rLeaf : rInt;
                __stcg(&H[internal].left, ln);
                __stcg(&H[internal].right, rn);
                __stcg(&H[ln].parent, internal);
                __stcg(&H[rn].parent, internal);
                Box b; b.mn = mk3(FLT_MAX); b.mx = mk3(-FLT_MAX);
#pragma unroll
                for (uint32_t i = 0; i < 7; i++)
                    if ((1u << i) & em) {
                        const float* sb = s_box[oct][i];
                        Box t; t.mn = mk3(sb[0], sb[1], sb[2]); t.mx = mk3(sb[3], sb[4], sb[5]);
                        b = combine(b, t);
                    }
                st_box(aabb, internal, b);
                __threadfence();
            }
        }
        __syncwarp();
        // ---- TraverseToParent: the second octet to arrive at the parent goes on with it
        uint32_t other = 0;
        const bool climbing = have && root != 0;
        if (climbing && ol == 0) {
            __threadfence();
            other = atomicAdd(&numTris[rparent], mine);
        }
        other = __shfl_sync(FULL, other, 0, 8);
        if (climbing && other != 0) {
            __threadfence();
            const uint32_t sibling = (pl == root) ? pr : pl;
            Box b = combine(rb, ld_box(aabb, sibling)); // every lane computes it, lane 0 publishes it
            if (ol == 0) st_box(aabb, rparent, b);
            root = rparent; rl = pl; rr = pr; rparent = pparent; mine += other; rb = b;
        } else {
            have = false;
        }
        __syncwarp();
    }
}

#undef cost
#undef part

// The bottom-up climb shared by the bottom-level and top-level node writers (ComputeAABBs.hlsli:69-172): starts at a leaf
// whose 32-byte node is already written; the second arrival at a parent writes the parent's node (smaller subtree
// left, D1) and goes on. `nodes` = the structure's node array (8 floats per node).
__device__ __forceinline__ void refit_climb(uint32_t node, const HNode* H, float* nodes, unsigned long long* counter, uint32_t* depthOut) {
    uint32_t count = 1, height = 0;
    while (node != 0) {
        uint32_t parent = ld_u(&H[node].parent);
        __threadfence();
        const unsigned long long arrived = atomicAdd(&counter[parent], ((unsigned long long)height << 32) | count);
        const uint32_t other = (uint32_t)arrived, otherHeight = (uint32_t)(arrived >> 32);
        if (other == 0) return;
        __threadfence();
        uint32_t l = ld_u(&H[parent].left), r = ld_u(&H[parent].right);
        uint32_t lc = (l == node) ? count : other, rc = (l == node) ? other : count;
        if (lc > rc) { uint32_t t = l; l = r; r = t; } // smaller subtree left; ties keep Karras order
        const float4* A = (const float4*)(nodes + 8 * (size_t)l);
        const float4* B = (const float4*)(nodes + 8 * (size_t)r);
        const float4 a0 = __ldcg(A), a1 = __ldcg(A + 1), b0 = __ldcg(B), b1 = __ldcg(B + 1);
        f3 ac = mk3(a0.x, a0.y, a0.z), ah = mk3(a1.x, a1.y, a1.z);
        f3 bc = mk3(b0.x, b0.y, b0.z), bh = mk3(b1.x, b1.y, b1.z);
        f3 mn = min3(ac - ah, bc - bh), mx = max3(ac + ah, bc + bh);
        f3 c = (mn + mx) * 0.5f;
        f3 h = mx - c;
        float4* nd = (float4*)(nodes + 8 * (size_t)parent);
        __stcg(nd, make_float4(c.x, c.y, c.z, __uint_as_float(l & 0x3fffffffu)));
        __stcg(nd + 1, make_float4(h.x, h.y, h.z, __uint_as_float(r)));
        node = parent;
        count += other;
        height = (height > otherHeight ? height : otherHeight) + 1;
        if (node == 0) *depthOut = height;
    }
}

// --------------------------------------------------------------------- refit
// ComputeAABBs.hlsli:69-172 (+ PrepareForComputeAABBs header)
// The climb also carries the subtree height (count in the low word of the 64-bit arrival counter, height in the
// high word: the value the first arrival leaves is exactly what the second one reads back), so the depth of the
// finished tree is known when the root is written. The traversal keeps at most one waiting far child per level,
// so `depth` bounds the stack it needs: the host rejects a tree deeper than TB_STACK_DEPTH instead of the traversal
// dropping children silently.
__global__ void k_refit(uint32_t n, const HNode* H, uint8_t* bvh, unsigned long long* counter, uint32_t* depthOut) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n) return;
    const uint32_t nInternal = n - 1, total = 2 * n - 1;
    // 64-bit addressing throughout; the header's four words are the reference's 32-bit byte offsets
    // (RayTracingHlslCompat.h:344-398), which the host guarantees to fit (tb_max_triangles: 116 N - 16 < 4 GiB)
    const size_t offPrims = 16 + 32 * (size_t)total;
    if (tid == 0) {
        uint32_t* hd = (uint32_t*)bvh;
        hd[0] = 16; hd[1] = (uint32_t)offPrims; hd[2] = (uint32_t)(offPrims + 40 * (size_t)n); hd[3] = (uint32_t)(offPrims + 52 * (size_t)n);
        if (n == 1) *depthOut = 0;
    }
    float* nodes = (float*)(bvh + 16);
    const Prim* prims = (const Prim*)(bvh + offPrims);
    uint32_t node = total - tid - 1;
    {
        f3 c, h;
        leaf_box(prims, node - nInternal, c, h);
        float4* nd = (float4*)(nodes + 8 * (size_t)node); // 16-byte aligned: the node array starts 16 bytes into the buffer
        __stcg(nd, make_float4(c.x, c.y, c.z, __uint_as_float((node - nInternal) | 0x80000000u)));
        __stcg(nd + 1, make_float4(h.x, h.y, h.z, __uint_as_float(1u)));
    }
    refit_climb(node, H, nodes, counter, depthOut);
}

// -------------------------------------------------------------------- update
// BuildRaytracingAccelerationStructure with PERFORM_UPDATE (GpuBVH2Builder.cpp:165-234): the hierarchy stays, the moved
// triangles go straight into their sorted slots and every box is refitted bottom-up (ComputeAABBs.hlsli:39-67 reads the
// children from the stored node flags and the parents from the aabbParentBuffer). The reference finds a triangle's slot
// through a sort cache written at build time; a slot's own metadata (geometry, primitive index) names the same triangle,
// so no cache is kept: one thread per SLOT reloads its triangle. Parents are derived from the stored child indices.
__global__ void k_update_prims(const BuildGeometry* __restrict__ geoms, uint32_t n, uint8_t* bvh) {
    uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    const size_t offPrims = 16 + 32 * (2 * (size_t)n - 1);
    Prim* prims = (Prim*)(bvh + offPrims);
    const Meta m = ((const Meta*)(bvh + offPrims + 40 * (size_t)n))[slot];
    Prim p;
    load_triangle(geoms[m.geom], m.prim, p);
    uint32_t* dst = (uint32_t*)(prims + slot);
    const uint32_t* src = (const uint32_t*)&p;
#pragma unroll
    for (int k = 0; k < 10; k++) dst[k] = src[k];
}
__global__ void k_update_parents(uint32_t n, const uint8_t* __restrict__ bvh, uint32_t* __restrict__ parent) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const RefNode* nodes = (const RefNode*)(bvh + 16);
    parent[nodes[i].flags & 0x3fffffffu] = i;
    parent[nodes[i].right] = i;
    if (i == 0) parent[0] = 0xffffffffu;
}
__global__ void k_update_refit(uint32_t n, uint8_t* bvh, const uint32_t* __restrict__ parent, uint32_t* arrive) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n) return;
    const uint32_t nInternal = n - 1, total = 2 * n - 1;
    float* nodes = (float*)(bvh + 16);
    const Prim* prims = (const Prim*)(bvh + 16 + 32 * (size_t)total);
    uint32_t node = total - tid - 1;
    {
        f3 c, h;
        leaf_box(prims, node - nInternal, c, h);
        float4* nd = (float4*)(nodes + 8 * (size_t)node);
        __stcg(nd, make_float4(c.x, c.y, c.z, __uint_as_float((node - nInternal) | 0x80000000u)));
        __stcg(nd + 1, make_float4(h.x, h.y, h.z, __uint_as_float(1u)));
    }
    while (node != 0) {
        const uint32_t up = parent[node];
        __threadfence();
        if (atomicAdd(&arrive[up], 1u) == 0) return; // the second arrival continues
        __threadfence();
        float4* nd = (float4*)(nodes + 8 * (size_t)up);
        const float4 p0 = __ldcg(nd), p1 = __ldcg(nd + 1);        // keeps its child references (w lanes)
        const uint32_t l = __float_as_uint(p0.w) & 0x3fffffffu, r = __float_as_uint(p1.w);
        const float4* A = (const float4*)(nodes + 8 * (size_t)l);
        const float4* B = (const float4*)(nodes + 8 * (size_t)r);
        const float4 a0 = __ldcg(A), a1 = __ldcg(A + 1), b0 = __ldcg(B), b1 = __ldcg(B + 1);
        f3 ac = mk3(a0.x, a0.y, a0.z), ah = mk3(a1.x, a1.y, a1.z);
        f3 bc = mk3(b0.x, b0.y, b0.z), bh = mk3(b1.x, b1.y, b1.z);
        f3 mn = min3(ac - ah, bc - bh), mx = max3(ac + ah, bc + bh);
        f3 c = (mn + mx) * 0.5f;
        f3 h = mx - c;
        __stcg(nd, make_float4(c.x, c.y, c.z, p0.w));
        __stcg(nd + 1, make_float4(h.x, h.y, h.z, p1.w));
        node = up;
    }
}

// ---------------------------------------------------------------------- widen
__global__ void k_widen(uint32_t n, const uint8_t* __restrict__ bvh, PairNode* __restrict__ pairs, WideTri* __restrict__ tris) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nInternal = n - 1, total = 2 * n - 1;
    const RefNode* nodes = (const RefNode*)(bvh + 16);
    if (i < nInternal) {
        RefNode nd = nodes[i];
        uint32_t l = nd.flags & 0x3fffffffu, r = nd.right;
        RefNode L = nodes[l], R = nodes[r];
        uint32_t lref = (L.flags & 0x80000000u) ? (0x80000000u | (L.flags & 0x3fffffffu)) : l;
        uint32_t rref = (R.flags & 0x80000000u) ? (0x80000000u | (R.flags & 0x3fffffffu)) : r;
        PairNode p;
        p.lc = make_float4(L.c[0], L.c[1], L.c[2], __uint_as_float(lref));
        p.lh = make_float4(L.h[0], L.h[1], L.h[2], __uint_as_float(rref));
        p.rc = make_float4(R.c[0], R.c[1], R.c[2], 0.0f);
        p.rh = make_float4(R.h[0], R.h[1], R.h[2], 0.0f);
        pairs[i] = p;
    }
    if (i < n) {
        const Prim* prims = (const Prim*)(bvh + 16 + 32 * (size_t)total);
        const Meta* meta = (const Meta*)(bvh + 16 + 32 * (size_t)total + 40 * (size_t)n);
        const float* v = prims[i].v;
        Meta m = meta[i];
        WideTri t;
        t.v0 = make_float4(v[0], v[1], v[2], __uint_as_float(m.geom));
        t.v1 = make_float4(v[3], v[4], v[5], __uint_as_float(m.prim));
        t.v2 = make_float4(v[6], v[7], v[8], 0.0f);
        tris[i] = t;
    }
}


// ------------------------------------------------------------------ top level
// BuildRaytracingAccelerationStructure for TYPE_TOP_LEVEL (GpuBVH2Builder.cpp:116-146, SceneType::BottomLevelBVHs): the
// same pipeline over instances instead of triangles, no treelet pass.
//   k_tlas_load    TopLevelLoadAABBs.hlsli:58-100: world box = TransformAABB of the bottom-level root box, the desc's
//                  transform replaced by InverseAffineTransform, ObjectToWorld kept (RayTracingHelper.hlsli:287-344);
//                  scene box
//   k_tlas_morton  Morton code of the box centre; then the radix sort by (code, index) and k_karras, shared with the
//                  bottom level
//   k_tlas_emit    sorted BVHMetadata records (116 B), leaf nodes, the query's per-instance records
//   k_tlas_refit   TopLevelComputeAABBs.hlsl + ComputeAABBs.hlsli: smaller subtree left (D1), bottom-up
struct Mat34 { float m[3][4]; };
__device__ __forceinline__ f3 mul_point(const Mat34& a, f3 p, float w) { // mul(float3x4, float4): one dot product per row, left to right
    return mk3(((a.m[0][0] * p.x + a.m[0][1] * p.y) + a.m[0][2] * p.z) + a.m[0][3] * w,
               ((a.m[1][0] * p.x + a.m[1][1] * p.y) + a.m[1][2] * p.z) + a.m[1][3] * w,
               ((a.m[2][0] * p.x + a.m[2][1] * p.y) + a.m[2][2] * p.z) + a.m[2][3] * w);
}
__device__ __forceinline__ float determinant(const Mat34& t) { // RayTracingHelper.hlsli:287-295
    return ((((t.m[0][0] * t.m[1][1] * t.m[2][2] - t.m[0][0] * t.m[2][1] * t.m[1][2]) - t.m[1][0] * t.m[0][1] * t.m[2][2]) +
             t.m[1][0] * t.m[2][1] * t.m[0][2]) + t.m[2][0] * t.m[0][1] * t.m[1][2]) - t.m[2][0] * t.m[1][1] * t.m[0][2];
}
__device__ Mat34 inverse_affine(const Mat34& a) { // :297-316, term by term (the constant fourth row 0 0 0 1 written out as the shader has it)
    const float (*t)[4] = a.m;
    const float invDet = 1.0f / determinant(a);
    Mat34 r;
    r.m[0][0] = invDet * ((t[1][1] * (t[2][2] * 1.0f - 0.0f * t[2][3]) + t[2][1] * (0.0f * t[1][3] - t[1][2] * 1.0f)) + 0.0f * (t[1][2] * t[2][3] - t[2][2] * t[1][3]));
    r.m[1][0] = invDet * ((t[1][2] * (t[2][0] * 1.0f - 0.0f * t[2][3]) + t[2][2] * (0.0f * t[1][3] - t[1][0] * 1.0f)) + 0.0f * (t[1][0] * t[2][3] - t[2][0] * t[1][3]));
    r.m[2][0] = invDet * ((t[1][3] * (t[2][0] * 0.0f - 0.0f * t[2][1]) + t[2][3] * (0.0f * t[1][1] - t[1][0] * 0.0f)) + 1.0f * (t[1][0] * t[2][1] - t[2][0] * t[1][1]));
    r.m[0][1] = invDet * ((t[2][1] * (t[0][2] * 1.0f - 0.0f * t[0][3]) + 0.0f * (t[2][2] * t[0][3] - t[0][2] * t[2][3])) + t[0][1] * (0.0f * t[2][3] - t[2][2] * 1.0f));
    r.m[1][1] = invDet * ((t[2][2] * (t[0][0] * 1.0f - 0.0f * t[0][3]) + 0.0f * (t[2][0] * t[0][3] - t[0][0] * t[2][3])) + t[0][2] * (0.0f * t[2][3] - t[2][0] * 1.0f));
    r.m[2][1] = invDet * ((t[2][3] * (t[0][0] * 0.0f - 0.0f * t[0][1]) + 1.0f * (t[2][0] * t[0][1] - t[0][0] * t[2][1])) + t[0][3] * (0.0f * t[2][1] - t[2][0] * 0.0f));
    r.m[0][2] = invDet * ((0.0f * (t[0][2] * t[1][3] - t[1][2] * t[0][3]) + t[0][1] * (t[1][2] * 1.0f - 0.0f * t[1][3])) + t[1][1] * (0.0f * t[0][3] - t[0][2] * 1.0f));
    r.m[1][2] = invDet * ((0.0f * (t[0][0] * t[1][3] - t[1][0] * t[0][3]) + t[0][2] * (t[1][0] * 1.0f - 0.0f * t[1][3])) + t[1][2] * (0.0f * t[0][3] - t[0][0] * 1.0f));
    r.m[2][2] = invDet * ((1.0f * (t[0][0] * t[1][1] - t[1][0] * t[0][1]) + t[0][3] * (t[1][0] * 0.0f - 0.0f * t[1][1])) + t[1][3] * (0.0f * t[0][1] - t[0][0] * 0.0f));
    r.m[0][3] = invDet * ((t[0][1] * (t[2][2] * t[1][3] - t[1][2] * t[2][3]) + t[1][1] * (t[0][2] * t[2][3] - t[2][2] * t[0][3])) + t[2][1] * (t[1][2] * t[0][3] - t[0][2] * t[1][3]));
    r.m[1][3] = invDet * ((t[0][2] * (t[2][0] * t[1][3] - t[1][0] * t[2][3]) + t[1][2] * (t[0][0] * t[2][3] - t[2][0] * t[0][3])) + t[2][2] * (t[1][0] * t[0][3] - t[0][0] * t[1][3]));
    r.m[2][3] = invDet * ((t[0][3] * (t[2][0] * t[1][1] - t[1][0] * t[2][1]) + t[1][3] * (t[0][0] * t[2][1] - t[2][0] * t[0][1])) + t[2][3] * (t[1][0] * t[0][1] - t[0][0] * t[1][1]));
    return r;
}
#define TLAS_META_WORDS 29 // BVHMetadata: 116 bytes
// what k_tlas_load leaves per instance (caller order): the box, then the metadata record
struct TlasLoaded { float4 c, h; uint32_t meta[TLAS_META_WORDS]; uint32_t pad[3]; }; // 160 bytes
__global__ void k_tlas_load(const TbInstanceDesc* __restrict__ inst, const TlasBlasInfo* __restrict__ blas, uint32_t n, TlasLoaded* __restrict__ out,
                            uint32_t* __restrict__ sceneBox) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const TbInstanceDesc d = inst[i];
    Mat34 o2w;
    for (int k = 0; k < 12; k++) o2w.m[k >> 2][k & 3] = d.Transform[k];
    const Mat34 w2o = inverse_affine(o2w);
    // BoundingBoxToAABB, TransformAABB over the eight corners, AABBtoBoundingBox
    const TlasBlasInfo b = blas[i];
    const f3 rc = mk3(b.c[0], b.c[1], b.c[2]), rh = mk3(b.h[0], b.h[1], b.h[2]);
    const f3 bmn = rc - rh, bmx = rc + rh;
    f3 wmn = mk3(FLT_MAX), wmx = mk3(-FLT_MAX);
    for (int k = 0; k < 8; k++) {
        const f3 v = mul_point(o2w, mk3((k & 4) ? bmx.x : bmn.x, (k & 2) ? bmx.y : bmn.y, (k & 1) ? bmx.z : bmn.z), 1.0f);
        wmn = min3(wmn, v); wmx = max3(wmx, v);
    }
    const f3 c = (wmn + wmx) * 0.5f, h = wmx - c;
    TlasLoaded& o = out[i];
    o.c = make_float4(c.x, c.y, c.z, 0.0f);
    o.h = make_float4(h.x, h.y, h.z, 0.0f);
    for (int k = 0; k < 12; k++) { o.meta[k] = __float_as_uint(w2o.m[k >> 2][k & 3]); o.meta[16 + k] = __float_as_uint(o2w.m[k >> 2][k & 3]); }
    o.meta[12] = d.InstanceIDAndMask; o.meta[13] = d.InstanceContributionToHitGroupIndexAndFlags;
    o.meta[14] = (uint32_t)d.AccelerationStructure; o.meta[15] = (uint32_t)(d.AccelerationStructure >> 32);
    o.meta[28] = i;
    // the scene box spans the boxes as the leaves store them (centre -/+ half extent); min / max are order independent
    const f3 lo = c - h, hi = c + h;
    atomicMin(&sceneBox[0], float_to_ordered(lo.x)); atomicMin(&sceneBox[1], float_to_ordered(lo.y)); atomicMin(&sceneBox[2], float_to_ordered(lo.z));
    atomicMax(&sceneBox[3], float_to_ordered(hi.x)); atomicMax(&sceneBox[4], float_to_ordered(hi.y)); atomicMax(&sceneBox[5], float_to_ordered(hi.z));
}

__global__ void k_tlas_morton(const TlasLoaded* __restrict__ loaded, uint32_t n, const uint32_t* __restrict__ sceneBox, uint32_t* __restrict__ codes,
                              uint32_t* __restrict__ order) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const f3 smin = mk3(ordered_to_float(sceneBox[0]), ordered_to_float(sceneBox[1]), ordered_to_float(sceneBox[2]));
    const f3 smax = mk3(ordered_to_float(sceneBox[3]), ordered_to_float(sceneBox[4]), ordered_to_float(sceneBox[5]));
    const float4 c4 = loaded[i].c;
    const f3 dim = max3(smax - smin, mk3(0.00001f)); // CalculateMortonCodesBindings.h:117-162
    const f3 unit = (mk3(c4.x, c4.y, c4.z) - smin) / dim;
    const f3 adj = min3(max3(unit * 1024.0f, mk3(0.0f)), mk3(1023.0f));
    const uint32_t coords[3] = {(uint32_t)adj.y, (uint32_t)adj.x, (uint32_t)adj.z};
    uint32_t code = 0;
#pragma unroll
    for (uint32_t bit = 0; bit < 10; bit++)
#pragma unroll
        for (uint32_t axis = 0; axis < 3; axis++)
            if ((1u << bit) & coords[axis]) code |= 1u << (bit * 3 + axis);
    codes[i] = code;
    order[i] = i;
}

// dst: header (16 B) | 32-byte nodes (2n - 1) | 116-byte metadata (n, sorted); records: one TlasInstanceRecord per sorted leaf
__global__ void k_tlas_emit(const TlasLoaded* __restrict__ loaded, const TlasBlasInfo* __restrict__ blas, const uint32_t* __restrict__ order, uint32_t n,
                            uint8_t* __restrict__ dst, TlasInstanceRecord* __restrict__ records) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t nInternal = n - 1, total = 2 * n - 1;
    const uint32_t offBoxes = 16, offMeta = offBoxes + 32 * total;
    if (i == 0) { // OffsetToLeafNodeMetaDataOffset = 4 (RayTracingHelper.hlsli:46); word 2 is not used by a top level
        uint32_t* hd = (uint32_t*)dst;
        hd[0] = offBoxes; hd[1] = offMeta; hd[2] = 0; hd[3] = offMeta + 116 * n;
    }
    const uint32_t src = order[i];
    const TlasLoaded& L = loaded[src];
    uint32_t* m = (uint32_t*)(dst + offMeta + 116 * (size_t)i);
    for (int k = 0; k < TLAS_META_WORDS; k++) m[k] = L.meta[k];
    float4* nd = (float4*)(dst + offBoxes + 32 * (size_t)(nInternal + i));
    nd[0] = make_float4(L.c.x, L.c.y, L.c.z, __uint_as_float(i | 0x80000000u));
    nd[1] = make_float4(L.h.x, L.h.y, L.h.z, __uint_as_float(1u));
    const TlasBlasInfo b = blas[src];
    TlasInstanceRecord r;
    for (int k = 0; k < 12; k++) r.worldToObject[k] = __uint_as_float(L.meta[k]);
    r.pairs = b.pairs; r.tris = b.tris; r.rootRef = b.rootRef;
    r.instanceIndex = L.meta[28];
    r.mask = L.meta[12] >> 24;
    r.pad = 0;
    records[i] = r;
}

// refit_climb from every leaf node (written by k_tlas_emit)
__global__ void k_tlas_refit(uint32_t n, const HNode* H, uint8_t* dst, unsigned long long* counter, uint32_t* depthOut) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n) return;
    if (tid == 0 && n == 1) *depthOut = 0;
    refit_climb(2 * n - 2 - tid, H, (float*)(dst + 16), counter, depthOut);
}

} // namespace

uint64_t bvh_ref_bytes(uint32_t n) { return 16ull + 32ull * (2ull * n - 1) + 40ull * n + 12ull * n; }

// The builder's temporaries, carved out of ONE scratch allocation (the caller's, as in
// BuildRaytracingAccelerationStructure's ScratchAccelerationStructureData, or the library's own).
namespace {
struct ScratchLayout {
    size_t prims, meta, codes, order, codesAlt, orderAlt, sceneBox, numTris, baseCount, baseRoots, H, aabb, radixHist, arrive, depth, end;
    uint32_t sortBlocks;
};
ScratchLayout scratch_layout(uint32_t n) {
    ScratchLayout L;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t at = off; off += (bytes + 255) & ~(size_t)255; return at; };
    const size_t total = 2 * (size_t)n - 1;
    L.sortBlocks = (n + RS_TILE - 1) / RS_TILE;
    L.prims = take(sizeof(Prim) * (size_t)n);
    L.meta = take(sizeof(Meta) * (size_t)n);
    L.codes = take(4 * (size_t)n); L.order = take(4 * (size_t)n);
    L.codesAlt = take(4 * (size_t)n); L.orderAlt = take(4 * (size_t)n);
    L.sceneBox = take(6 * 4);
    L.numTris = take(4 * (size_t)n);
    L.baseCount = take(4);
    L.baseRoots = take(4 * ((size_t)n / 7 + 1));
    L.H = take(sizeof(HNode) * total);
    L.aabb = take(32 * total);
    L.radixHist = take(4 * 256 * ((size_t)L.sortBlocks + 1));
    L.arrive = take(8 * (size_t)n);
    L.depth = take(4);
    L.end = off;
    return L;
}
} // namespace

uint64_t bvh_scratch_bytes(uint32_t n) { return n ? scratch_layout(n).end : 0; }
// update: parent of every node + one arrival counter per internal node
uint64_t bvh_update_scratch_bytes(uint32_t n) { return n ? 4 * (2 * (uint64_t)n - 1) + 256 + 4 * (uint64_t)n : 0; }

cudaError_t update_bvh(const BuildGeometry* d_geoms, uint32_t n, DeviceBvh& bvh, void* scratch, cudaStream_t stream, LaunchCounter& lc) {
    const uint32_t T = 256;
    auto grid = [&](uint32_t c) { return (c + T - 1) / T; };
    cudaError_t err;
    uint32_t* parent = (uint32_t*)scratch;
    uint32_t* arrive = (uint32_t*)((uint8_t*)scratch + ((4 * (2 * (size_t)n - 1) + 255) & ~(size_t)255));
    k_update_prims<<<grid(n), T, 0, stream>>>(d_geoms, n, bvh.ref); lc.count++;
    if (n > 1) { k_update_parents<<<grid(n - 1), T, 0, stream>>>(n, bvh.ref, parent); lc.count++; }
    if ((err = cudaMemsetAsync(arrive, 0, 4 * (size_t)n, stream)) != cudaSuccess) return err;
    k_update_refit<<<grid(n), T, 0, stream>>>(n, bvh.ref, parent, arrive); lc.count++;
    k_widen<<<grid(n), T, 0, stream>>>(n, bvh.ref, bvh.pairs, bvh.tris); lc.count++;
    if ((err = cudaMemcpyAsync(&bvh.root, bvh.ref + 16, sizeof(RefNode), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return err;
    if ((err = cudaGetLastError()) != cudaSuccess) return err;
    return cudaStreamSynchronize(stream);
}

// Builds the reference layout into out.ref and the traversal layout into out.pairs / out.tris. `scratch` holds at
// least bvh_scratch_bytes(n) bytes (256-byte aligned). No allocation happens in here; the stream is synchronised
// once at the end (the root box and the tree depth come back to the host).
cudaError_t build_bvh(const BuildGeometry* d_geoms, const uint32_t* d_triPrefix, uint32_t numGeoms, uint32_t n, int treeletPasses,
                      DeviceBvh& out, void* scratch, cudaStream_t stream, LaunchCounter& lc) {
    const uint32_t total = 2 * n - 1, nInternal = n - 1;
    const uint32_t T = 256;
    auto grid = [&](uint32_t c) { return (c + T - 1) / T; };
    cudaError_t err;
#define CK(x) do { err = (x); if (err != cudaSuccess) return err; } while (0)
    const ScratchLayout L = scratch_layout(n);
    uint8_t* base = (uint8_t*)scratch;
    Prim* prims = (Prim*)(base + L.prims); Meta* meta = (Meta*)(base + L.meta);
    uint32_t *codes = (uint32_t*)(base + L.codes), *order = (uint32_t*)(base + L.order);
    uint32_t *codesAlt = (uint32_t*)(base + L.codesAlt), *orderAlt = (uint32_t*)(base + L.orderAlt);
    uint32_t *sceneBox = (uint32_t*)(base + L.sceneBox), *numTris = (uint32_t*)(base + L.numTris);
    uint32_t *baseCount = (uint32_t*)(base + L.baseCount), *baseRoots = (uint32_t*)(base + L.baseRoots);
    HNode* H = (HNode*)(base + L.H); float* aabb = (float*)(base + L.aabb);
    const uint32_t sortBlocks = L.sortBlocks;
    uint32_t* radixHist = (uint32_t*)(base + L.radixHist);
    uint32_t* radixDigitStart = radixHist + 256 * (size_t)sortBlocks;
    unsigned long long* arrive = (unsigned long long*)(base + L.arrive);
    uint32_t* depthDev = (uint32_t*)(base + L.depth);

    uint32_t boxInit[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    CK(cudaMemcpyAsync(sceneBox, boxInit, sizeof(boxInit), cudaMemcpyHostToDevice, stream));
    k_load_prims<<<grid(n), T, 0, stream>>>(d_geoms, d_triPrefix, numGeoms, n, prims, meta, sceneBox); lc.count++;
    k_morton<<<grid(n), T, 0, stream>>>(prims, n, sceneBox, codes, order); lc.count++;
    uint32_t *kIn = codes, *vIn = order, *kOut = codesAlt, *vOut = orderAlt;
    for (int shift = 0; shift < 30; shift += 8) {
        k_radix_hist<<<sortBlocks, RS_THREADS, 0, stream>>>(kIn, n, shift, sortBlocks, radixHist); lc.count++;
        k_radix_scan_rows<<<256, 256, 0, stream>>>(radixHist, sortBlocks, radixDigitStart); lc.count++;
        k_radix_scan_digits<<<1, 256, 0, stream>>>(radixDigitStart); lc.count++;
        k_radix_scatter<<<sortBlocks, RS_THREADS, 0, stream>>>(kIn, vIn, n, shift, sortBlocks, radixHist, radixDigitStart, kOut, vOut); lc.count++;
        uint32_t* t = kIn; kIn = kOut; kOut = t; t = vIn; vIn = vOut; vOut = t;
    }
    const uint32_t* sortedCodes = kIn; const uint32_t* sortedOrder = vIn;
    uint8_t* bvh = out.ref;
    Prim* sortedPrims = (Prim*)(bvh + 16 + 32 * (size_t)total);
    Meta* sortedMeta = (Meta*)(bvh + 16 + 32 * (size_t)total + 40 * (size_t)n);
    k_rearrange<<<grid(n), T, 0, stream>>>(prims, meta, sortedOrder, n, sortedPrims, sortedMeta); lc.count++;
    if (n > 1) {
        k_karras<<<grid(nInternal), T, 0, stream>>>(sortedCodes, n, H); lc.count++;
        uint32_t minTris = 7;
        for (int pass = 0; pass < treeletPasses; pass++) {
            if (minTris > n) break;
            k_treelet_clear<<<grid(n), T, 0, stream>>>(numTris, nInternal, baseCount); lc.count++;
            k_find_treelets<<<grid(n), T, 0, stream>>>(sortedPrims, n, minTris, H, aabb, numTris, baseCount, baseRoots); lc.count++;
            uint32_t maxRoots = n / minTris + 1;
            uint32_t blocks = (maxRoots + TL_OCTETS - 1) / TL_OCTETS; // one octet (8 lanes) per base root
            if (blocks > 148 * 32) blocks = 148 * 32;
            k_treelet_reorder<<<blocks, 128, 0, stream>>>(n, H, aabb, numTris, baseCount, baseRoots); lc.count++;
            minTris *= 2;
        }
    }
    CK(cudaMemsetAsync(arrive, 0, 8 * (size_t)n, stream));
    k_refit<<<grid(n), T, 0, stream>>>(n, H, bvh, arrive, depthDev); lc.count++;
    k_widen<<<grid(n), T, 0, stream>>>(n, bvh, out.pairs, out.tris); lc.count++;
    CK(cudaMemcpyAsync(&out.root, bvh + 16, sizeof(RefNode), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(&out.depth, depthDev, 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(stream));
    out.numPrims = n;
#undef CK
    return cudaSuccess;
}

// ---- top level: scratch layout and the build
namespace {
struct TlasScratchLayout { size_t inst, blas, loaded, codes, order, codesAlt, orderAlt, sceneBox, H, radixHist, arrive, depth, end; uint32_t sortBlocks; };
TlasScratchLayout tlas_scratch_layout(uint32_t n) {
    TlasScratchLayout L;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t at = off; off += (bytes + 255) & ~(size_t)255; return at; };
    L.sortBlocks = (n + RS_TILE - 1) / RS_TILE;
    L.inst = take(sizeof(TbInstanceDesc) * (size_t)n);
    L.blas = take(sizeof(TlasBlasInfo) * (size_t)n);
    L.loaded = take(sizeof(TlasLoaded) * (size_t)n);
    L.codes = take(4 * (size_t)n); L.order = take(4 * (size_t)n);
    L.codesAlt = take(4 * (size_t)n); L.orderAlt = take(4 * (size_t)n);
    L.sceneBox = take(6 * 4);
    L.H = take(sizeof(HNode) * (2 * (size_t)n - 1));
    L.radixHist = take(4 * 256 * ((size_t)L.sortBlocks + 1));
    L.arrive = take(8 * (size_t)n);
    L.depth = take(4);
    L.end = off;
    return L;
}
} // namespace

uint64_t tlas_scratch_bytes(uint32_t n) { return n ? tlas_scratch_layout(n).end : 0; }

// Builds the top level into `dst` (reference byte layout) and `records` (the device query's per-leaf table) from HOST
// arrays of instance descs and resolved bottom-level infos; everything else runs on `stream`. `scratch` holds at least
// tlas_scratch_bytes(n) bytes. One synchronisation at the end: the tree depth comes back to the host.
cudaError_t build_tlas(const TbInstanceDesc* h_instances, const TlasBlasInfo* h_blas, uint32_t n, uint8_t* dst, TlasInstanceRecord* records,
                       void* scratch, uint32_t* depthOut, cudaStream_t stream, LaunchCounter& lc) {
    const uint32_t T = 256;
    auto grid = [&](uint32_t c) { return (c + T - 1) / T; };
    cudaError_t err;
#define CK(x) do { err = (x); if (err != cudaSuccess) return err; } while (0)
    const TlasScratchLayout L = tlas_scratch_layout(n);
    uint8_t* base = (uint8_t*)scratch;
    TbInstanceDesc* inst = (TbInstanceDesc*)(base + L.inst);
    TlasBlasInfo* blas = (TlasBlasInfo*)(base + L.blas);
    TlasLoaded* loaded = (TlasLoaded*)(base + L.loaded);
    uint32_t *codes = (uint32_t*)(base + L.codes), *order = (uint32_t*)(base + L.order);
    uint32_t *codesAlt = (uint32_t*)(base + L.codesAlt), *orderAlt = (uint32_t*)(base + L.orderAlt);
    uint32_t* sceneBox = (uint32_t*)(base + L.sceneBox);
    HNode* H = (HNode*)(base + L.H);
    uint32_t* radixHist = (uint32_t*)(base + L.radixHist);
    uint32_t* radixDigitStart = radixHist + 256 * (size_t)L.sortBlocks;
    unsigned long long* arrive = (unsigned long long*)(base + L.arrive);
    uint32_t* depthDev = (uint32_t*)(base + L.depth);
    CK(cudaMemcpyAsync(inst, h_instances, sizeof(TbInstanceDesc) * (size_t)n, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(blas, h_blas, sizeof(TlasBlasInfo) * (size_t)n, cudaMemcpyHostToDevice, stream));
    const uint32_t boxInit[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    CK(cudaMemcpyAsync(sceneBox, boxInit, sizeof(boxInit), cudaMemcpyHostToDevice, stream));
    k_tlas_load<<<grid(n), T, 0, stream>>>(inst, blas, n, loaded, sceneBox); lc.count++;
    k_tlas_morton<<<grid(n), T, 0, stream>>>(loaded, n, sceneBox, codes, order); lc.count++;
    uint32_t *kIn = codes, *vIn = order, *kOut = codesAlt, *vOut = orderAlt;
    for (int shift = 0; shift < 30; shift += 8) { // stable LSD passes: equal codes keep the caller's order
        k_radix_hist<<<L.sortBlocks, RS_THREADS, 0, stream>>>(kIn, n, shift, L.sortBlocks, radixHist); lc.count++;
        k_radix_scan_rows<<<256, 256, 0, stream>>>(radixHist, L.sortBlocks, radixDigitStart); lc.count++;
        k_radix_scan_digits<<<1, 256, 0, stream>>>(radixDigitStart); lc.count++;
        k_radix_scatter<<<L.sortBlocks, RS_THREADS, 0, stream>>>(kIn, vIn, n, shift, L.sortBlocks, radixHist, radixDigitStart, kOut, vOut); lc.count++;
        uint32_t* t = kIn; kIn = kOut; kOut = t; t = vIn; vIn = vOut; vOut = t;
    }
    if (n > 1) { k_karras<<<grid(n - 1), T, 0, stream>>>(kIn, n, H); lc.count++; }
    k_tlas_emit<<<grid(n), T, 0, stream>>>(loaded, blas, vIn, n, dst, records); lc.count++;
    CK(cudaMemsetAsync(arrive, 0, 8 * (size_t)n, stream));
    k_tlas_refit<<<grid(n), T, 0, stream>>>(n, H, dst, arrive, depthDev); lc.count++;
    CK(cudaMemcpyAsync(depthOut, depthDev, 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(stream));
#undef CK
    return cudaSuccess;
}

} // namespace tbd
