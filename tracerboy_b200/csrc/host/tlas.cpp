// tlas.cpp — top-level acceleration structures at the SW-RT seam: BuildRaytracingAccelerationStructure
// for TYPE_TOP_LEVEL (GpuBVH2Builder.cpp:116-146, SceneType::BottomLevelBVHs) over instances of bottom-level structures
// built by tb_bvh_build_device, and the two-level ray query (TraverseFunction.hlsli with FAST_PATH 0) on the device.
//
// The top level is built on the HOST: it has one leaf per instance (tens to tens of thousands), needs each bottom-level
// root box (kept in the handle's cache / the structure's trailer) and is a few microseconds of work per instance; the
// bottom levels, where the triangles are, are built on the GPU. Stages as in the reference:
//   load        TopLevelLoadAABBs.hlsli:58-100 (world box = TransformAABB of the bottom-level root box, desc transform
//               replaced by InverseAffineTransform, ObjectToWorld kept; RayTracingHelper.hlsli:287-344)
//   scene box / Morton code of the box centre / sort by (code, index) / rearrange / Karras hierarchy; no treelet pass
//   boxes       TopLevelComputeAABBs.hlsl + ComputeAABBs.hlsli (smaller subtree left, ties keep hierarchy order: D1)
// The result buffer starts with the reference's byte layout (16-byte header, 32-byte nodes, 116-byte BVHMetadata per
// sorted leaf) and continues with a trailer and one InstanceRecord per sorted leaf for the device query.
#include <algorithm>
#include <cfloat>
#include <cstring>
#include "../common/tb_vec.h"
#include "handle.h"

using namespace tbm;

namespace tbd {
// device side (pathtrace.cu)
cudaError_t trace_rays_tlas(const uint8_t* tlasRef, const TlasInstanceRecord* records, uint32_t numInstances, const TbRay* d_rays, uint64_t n,
                            TbHit* d_hits, int numSMs, cudaStream_t stream, LaunchCounter& lc);
}

namespace {

struct Mat34 { float m[3][4]; };
struct Box { f3 mn, mx; };
struct RefNodeH { float c[3]; uint32_t flags; float h[3]; uint32_t right; };
struct MetaRec { float worldToObject[12]; uint32_t idAndMask, contribAndFlags, asLo, asHi; float objectToWorld[12]; uint32_t instanceIndex; };
static_assert(sizeof(MetaRec) == 116 && sizeof(RefNodeH) == 32, "reference layout");

inline f3 mul_point(const Mat34& a, f3 p, float w) { // mul(float3x4, float4): one dot product per row, left to right
    return mk3(((a.m[0][0] * p.x + a.m[0][1] * p.y) + a.m[0][2] * p.z) + a.m[0][3] * w,
               ((a.m[1][0] * p.x + a.m[1][1] * p.y) + a.m[1][2] * p.z) + a.m[1][3] * w,
               ((a.m[2][0] * p.x + a.m[2][1] * p.y) + a.m[2][2] * p.z) + a.m[2][3] * w);
}
inline float determinant(const Mat34& t) { // RayTracingHelper.hlsli:287-295
    return ((((t.m[0][0] * t.m[1][1] * t.m[2][2] - t.m[0][0] * t.m[2][1] * t.m[1][2]) - t.m[1][0] * t.m[0][1] * t.m[2][2]) +
             t.m[1][0] * t.m[2][1] * t.m[0][2]) + t.m[2][0] * t.m[0][1] * t.m[1][2]) - t.m[2][0] * t.m[1][1] * t.m[0][2];
}
inline Mat34 inverse_affine(const Mat34& a) { // :297-316, term by term
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
inline void instance_box(const RefNode& root, const Mat34& o2w, f3& c, f3& h) { // BoundingBoxToAABB, TransformAABB, AABBtoBoundingBox
    f3 rc = mk3(root.c[0], root.c[1], root.c[2]), rh = mk3(root.h[0], root.h[1], root.h[2]);
    Box b{rc - rh, rc + rh}, w{mk3(FLT_MAX), mk3(-FLT_MAX)};
    for (int i = 0; i < 8; i++) {
        f3 v = mul_point(o2w, mk3((i & 4) ? b.mx.x : b.mn.x, (i & 2) ? b.mx.y : b.mn.y, (i & 1) ? b.mx.z : b.mn.z), 1.0f);
        w.mn = min3(w.mn, v); w.mx = max3(w.mx, v);
    }
    c = (w.mn + w.mx) * 0.5f;
    h = w.mx - c;
}
uint32_t morton_code(f3 centroid, f3 smin, f3 smax) { // CalculateMortonCodesBindings.h:117-162
    f3 dim = max3(smax - smin, mk3(0.00001f));
    f3 unit = (centroid - smin) / dim;
    f3 adj = min3(max3(unit * 1024.0f, mk3(0.0f)), mk3(1023.0f));
    uint32_t coords[3] = {(uint32_t)adj.y, (uint32_t)adj.x, (uint32_t)adj.z};
    uint32_t code = 0;
    for (uint32_t bit = 0; bit < 10; bit++)
        for (uint32_t axis = 0; axis < 3; axis++)
            if ((1u << bit) & coords[axis]) code |= 1u << (bit * 3 + axis);
    return code;
}
struct Karras { // BuildBVHSplits.hlsli:18-141 on the sorted codes (ties broken by index)
    const uint32_t* codes; uint32_t n;
    static int clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
    int lcp(uint32_t a, uint32_t b) const {
        if (a >= n || b >= n) return -1;
        uint32_t ca = codes[a], cb = codes[b];
        return ca != cb ? clz(ca ^ cb) : clz(a ^ b) + 31;
    }
    void node(uint32_t idx, uint32_t& left, uint32_t& right) const {
        int d = lcp(idx, idx + 1) - lcp(idx, idx - 1);
        d = d < -1 ? -1 : (d > 1 ? 1 : d);
        int minPrefix = lcp(idx, idx - d), maxLength = 2;
        while (lcp(idx, idx + (uint32_t)(maxLength * d)) > minPrefix) maxLength *= 4;
        int length = 0;
        for (int t = maxLength / 2; t > 0; t /= 2)
            if (lcp(idx, idx + (uint32_t)((length + t) * d)) > minPrefix) length += t;
        uint32_t j = idx + (uint32_t)(length * d), first = std::min(idx, j), last = std::max(idx, j);
        int common = lcp(first, last), sp = (int)first, step = (int)(last - first);
        do {
            step = (step + 1) >> 1;
            int ns = sp + step;
            if ((uint32_t)ns < last && lcp(first, (uint32_t)ns) > common) sp = ns;
        } while (step > 1);
        const uint32_t split = (uint32_t)sp, nInternal = n - 1;
        left = (split == first) ? nInternal + split : split;
        right = (split + 1 == last) ? nInternal + split + 1 : split + 1;
    }
};

inline uint64_t up256(uint64_t v) { return (v + 255) & ~255ull; }
struct TlasTrailer { char magic[8]; uint32_t numInstances, depth; };
struct TlasLayout { uint64_t refBytes, trailer, records, end; };
TlasLayout tlas_layout(uint32_t n) {
    TlasLayout L;
    L.refBytes = 16ull + 32ull * (2ull * n - 1) + 116ull * n;
    L.trailer = up256(L.refBytes);
    L.records = L.trailer + 256;
    L.end = L.records + sizeof(TlasInstanceRecord) * (uint64_t)n;
    return L;
}

} // namespace

extern "C" {

TB_API int tb_tlas_prebuild_info(uint32_t numInstances, TbPrebuildInfo* out) {
    if (!out) return TB_ERR_INVALID_ARG;
    memset(out, 0, sizeof(*out));
    if (numInstances == 0) return TB_OK;
    if (numInstances > (1u << 24)) return TB_ERR_INVALID_ARG; // InstanceID and the hit-group contribution are 24 bit
    const TlasLayout L = tlas_layout(numInstances);
    out->ResultDataMaxSizeInBytes = L.end;
    out->ReferenceLayoutSizeInBytes = L.refBytes;
    out->ScratchDataSizeInBytes = 0;       // the top level is built on the host
    out->UpdateScratchDataSizeInBytes = 0;
    return TB_OK;
}

TB_API int tb_tlas_build_device(TbHandle* h, const TbInstanceDesc* instances, uint32_t n, uint32_t flags, void* dst, uint64_t dstBytes, void* cudaStream) {
    (void)flags;
    if (!h || !instances || n == 0 || !dst) return fail(h, TB_ERR_INVALID_ARG, "null/empty argument");
    if (n > (1u << 24)) return fail(h, TB_ERR_INVALID_ARG, "too many instances");
    const TlasLayout L = tlas_layout(n);
    if (dstBytes < L.end) return fail(h, TB_ERR_INVALID_ARG, "destination smaller than ResultDataMaxSizeInBytes");
    if ((uintptr_t)dst & 255) return fail(h, TB_ERR_INVALID_ARG, "dst must be 256-byte aligned");
    CUDA_OK(h, cudaSetDevice(h->device));
    cudaStream_t stream = cudaStream ? (cudaStream_t)cudaStream : h->stream;
    std::vector<DeviceBvh> blas(n);
    std::vector<RefNodeH> leaf(n);
    std::vector<MetaRec> meta(n);
    for (uint32_t i = 0; i < n; i++) {
        int rc = tbh::resolve_bottom_level(h, (const void*)(uintptr_t)instances[i].AccelerationStructure, stream, blas[i]);
        if (rc != TB_OK) return rc;
        Mat34 o2w;
        memcpy(o2w.m, instances[i].Transform, 48);
        const Mat34 w2o = inverse_affine(o2w);
        f3 c, hh;
        instance_box(blas[i].root, o2w, c, hh);
        leaf[i] = RefNodeH{{c.x, c.y, c.z}, 0x80000000u | i, {hh.x, hh.y, hh.z}, 0};
        MetaRec& m = meta[i];
        memcpy(m.worldToObject, w2o.m, 48);
        m.idAndMask = instances[i].InstanceIDAndMask; m.contribAndFlags = instances[i].InstanceContributionToHitGroupIndexAndFlags;
        m.asLo = (uint32_t)instances[i].AccelerationStructure; m.asHi = (uint32_t)(instances[i].AccelerationStructure >> 32);
        memcpy(m.objectToWorld, o2w.m, 48);
        m.instanceIndex = i;
    }
    f3 smin = mk3(FLT_MAX), smax = mk3(-FLT_MAX);
    for (uint32_t i = 0; i < n; i++) {
        f3 c = mk3(leaf[i].c[0], leaf[i].c[1], leaf[i].c[2]), hh = mk3(leaf[i].h[0], leaf[i].h[1], leaf[i].h[2]);
        smin = min3(c - hh, smin); smax = max3(c + hh, smax);
    }
    std::vector<uint32_t> codes(n), order(n);
    for (uint32_t i = 0; i < n; i++) { codes[i] = morton_code(mk3(leaf[i].c[0], leaf[i].c[1], leaf[i].c[2]), smin, smax); order[i] = i; }
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return codes[a] != codes[b] ? codes[a] < codes[b] : a < b; });
    std::vector<uint32_t> sc(n);
    for (uint32_t i = 0; i < n; i++) sc[i] = codes[order[i]];
    const uint32_t nInternal = n - 1, total = 2 * n - 1;
    std::vector<uint32_t> left(nInternal ? nInternal : 1), right(nInternal ? nInternal : 1);
    Karras K{sc.data(), n};
    for (uint32_t i = 0; i < nInternal; i++) K.node(i, left[i], right[i]);
    std::vector<uint8_t> bytes(L.end, 0);
    const uint32_t offBoxes = 16, offMeta = offBoxes + 32 * total, totalSize = offMeta + 116 * n;
    const uint32_t header[4] = {offBoxes, offMeta, 0, totalSize}; // OffsetToLeafNodeMetaDataOffset = 4 (RayTracingHelper.hlsli:46); word 2 is not written
    memcpy(bytes.data(), header, 16);
    RefNodeH* nodes = (RefNodeH*)(bytes.data() + offBoxes);
    MetaRec* sm = (MetaRec*)(bytes.data() + offMeta);
    TlasInstanceRecord* rec = (TlasInstanceRecord*)(bytes.data() + L.records);
    for (uint32_t i = 0; i < n; i++) {
        sm[i] = meta[order[i]];
        Mat34 o2w;
        memcpy(o2w.m, sm[i].objectToWorld, 48);
        const DeviceBvh& b = blas[order[i]];
        f3 c, hh;
        instance_box(b.root, o2w, c, hh); // TopLevelComputeAABBs.hlsl ComputeLeafAABB
        nodes[nInternal + i] = RefNodeH{{c.x, c.y, c.z}, i | 0x80000000u, {hh.x, hh.y, hh.z}, 1};
        TlasInstanceRecord& r = rec[i];
        memcpy(r.worldToObject, sm[i].worldToObject, 48);
        r.pairs = b.pairs; r.tris = b.tris;
        r.rootRef = (b.root.flags & 0x80000000u) ? (0x80000000u | (b.root.flags & 0x3fffffffu)) : 0u;
        r.instanceIndex = sm[i].instanceIndex;
        r.mask = sm[i].idAndMask >> 24;
        r.pad = 0;
    }
    uint32_t depth = 0;
    if (n > 1) { // bottom-up over the hierarchy (explicit post-order stack), counts decide the child swap
        std::vector<uint32_t> cnt(total, 1), height(total, 0);
        std::vector<std::pair<uint32_t, int>> st;
        st.push_back({0, 0});
        while (!st.empty()) {
            auto& top = st.back();
            const uint32_t node = top.first;
            if (node >= nInternal) { st.pop_back(); continue; }
            if (top.second == 0) { top.second = 1; st.push_back({left[node], 0}); }
            else if (top.second == 1) { top.second = 2; st.push_back({right[node], 0}); }
            else {
                uint32_t l = left[node], r = right[node];
                if (cnt[l] > cnt[r]) std::swap(l, r);
                cnt[node] = cnt[l] + cnt[r];
                height[node] = std::max(height[l], height[r]) + 1;
                f3 ac = mk3(nodes[l].c[0], nodes[l].c[1], nodes[l].c[2]), ah = mk3(nodes[l].h[0], nodes[l].h[1], nodes[l].h[2]);
                f3 bc = mk3(nodes[r].c[0], nodes[r].c[1], nodes[r].c[2]), bh = mk3(nodes[r].h[0], nodes[r].h[1], nodes[r].h[2]);
                f3 mn = min3(ac - ah, bc - bh), mx = max3(ac + ah, bc + bh);
                f3 c = (mn + mx) * 0.5f, hh = mx - c;
                nodes[node] = RefNodeH{{c.x, c.y, c.z}, l & 0x3fffffffu, {hh.x, hh.y, hh.z}, r};
                st.pop_back();
            }
        }
        depth = height[0];
    }
    if (depth > TB_TLAS_STACK_DEPTH) return fail(h, TB_ERR_NOT_IMPL, "top-level tree deeper than the query's stack");
    TlasTrailer t;
    memset(&t, 0, sizeof(t));
    memcpy(t.magic, "TBTL0001", 8);
    t.numInstances = n; t.depth = depth;
    memcpy(bytes.data() + L.trailer, &t, sizeof(t));
    CUDA_OK(h, cudaMemcpyAsync(dst, bytes.data(), bytes.size(), cudaMemcpyHostToDevice, stream));
    CUDA_OK(h, cudaStreamSynchronize(stream));
    h->topLevelBuilds[dst] = n;
    return TB_OK;
}

TB_API int tb_trace_rays_tlas_device(TbHandle* h, const void* tlas, uint64_t tlasBytes, const TbRay* dRays, uint64_t n, TbHit* dHits, void* cudaStream) {
    if (!h || !tlas || (n && (!dRays || !dHits))) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    CUDA_OK(h, cudaSetDevice(h->device));
    cudaStream_t stream = cudaStream ? (cudaStream_t)cudaStream : h->stream;
    uint32_t numInstances = 0;
    auto it = h->topLevelBuilds.find(tlas);
    if (it != h->topLevelBuilds.end()) numInstances = it->second;
    else { // built by another handle: the structure describes itself (header word 3 = 16 + 32 (2N-1) + 116 N)
        uint32_t hd[4];
        CUDA_OK(h, cudaMemcpyAsync(hd, tlas, sizeof(hd), cudaMemcpyDeviceToHost, stream));
        CUDA_OK(h, cudaStreamSynchronize(stream));
        if (hd[0] != 16 || hd[2] != 0 || hd[3] < 164 || (hd[3] + 16ull) % 180 || hd[1] + 116ull * ((hd[3] + 16ull) / 180) != hd[3]) return fail(h, TB_ERR_INVALID_ARG, "not a top-level structure built by tb_tlas_build_device");
        numInstances = (uint32_t)((hd[3] + 16ull) / 180);
        const TlasLayout L = tlas_layout(numInstances);
        if (tlasBytes < L.end) return fail(h, TB_ERR_INVALID_ARG, "top-level structure buffer too small");
        TlasTrailer t;
        CUDA_OK(h, cudaMemcpyAsync(&t, (const uint8_t*)tlas + L.trailer, sizeof(t), cudaMemcpyDeviceToHost, stream));
        CUDA_OK(h, cudaStreamSynchronize(stream));
        if (memcmp(t.magic, "TBTL0001", 8) != 0 || t.numInstances != numInstances) return fail(h, TB_ERR_INVALID_ARG, "top-level trailer is missing or corrupt");
        h->topLevelBuilds[tlas] = numInstances;
    }
    if (n == 0) return TB_OK;
    const TlasLayout L = tlas_layout(numInstances);
    CUDA_OK(h, trace_rays_tlas((const uint8_t*)tlas, (const TlasInstanceRecord*)((const uint8_t*)tlas + L.records), numInstances, dRays, n, dHits,
                               h->numSMs, stream, h->lc));
    return TB_OK;
}

} // extern "C"
