// tlas.cpp — top-level acceleration structures at the SW-RT seam: BuildRaytracingAccelerationStructure
// for TYPE_TOP_LEVEL (GpuBVH2Builder.cpp:116-146, SceneType::BottomLevelBVHs) over instances of bottom-level structures
// built by tb_bvh_build_device, and the two-level ray query (TraverseFunction.hlsli with FAST_PATH 0) on the device.
//
// The top level is built on the GPU by the bottom level's pipeline without the treelet pass (csrc/cuda/bvh_build.cu:
// k_tlas_load, Morton codes, radix sort, Karras hierarchy, k_tlas_emit, k_tlas_refit). The host only resolves each
// instance's bottom-level structure (root box and traversal-layout pointers, from the handle's cache or the structure's
// trailer) and uploads the descs. The result buffer starts with the reference's byte layout (16-byte header, 32-byte
// nodes, 116-byte BVHMetadata per sorted leaf) and continues with a trailer and one TlasInstanceRecord per sorted leaf for
// the device query.
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
    out->ScratchDataSizeInBytes = tlas_scratch_bytes(numInstances);
    out->UpdateScratchDataSizeInBytes = 0;
    return TB_OK;
}

TB_API int tb_tlas_build_device(TbHandle* h, const TbInstanceDesc* instances, uint32_t n, uint32_t flags, void* dst, uint64_t dstBytes, void* scratch,
                                uint64_t scratchBytes, void* cudaStream) {
    (void)flags;
    if (!h || !instances || n == 0 || !dst || !scratch) return fail(h, TB_ERR_INVALID_ARG, "null/empty argument");
    if (n > (1u << 24)) return fail(h, TB_ERR_INVALID_ARG, "too many instances");
    const TlasLayout L = tlas_layout(n);
    if (dstBytes < L.end) return fail(h, TB_ERR_INVALID_ARG, "destination smaller than ResultDataMaxSizeInBytes");
    if (scratchBytes < tlas_scratch_bytes(n)) return fail(h, TB_ERR_INVALID_ARG, "scratch smaller than ScratchDataSizeInBytes");
    if (((uintptr_t)dst | (uintptr_t)scratch) & 255) return fail(h, TB_ERR_INVALID_ARG, "dst and scratch must be 256-byte aligned");
    CUDA_OK(h, cudaSetDevice(h->device));
    cudaStream_t stream = cudaStream ? (cudaStream_t)cudaStream : h->stream;
    std::vector<TlasBlasInfo> blas(n);
    for (uint32_t i = 0; i < n; i++) {
        DeviceBvh b;
        int rc = tbh::resolve_bottom_level(h, (const void*)(uintptr_t)instances[i].AccelerationStructure, stream, b);
        if (rc != TB_OK) return rc;
        TlasBlasInfo& o = blas[i];
        memcpy(o.c, b.root.c, 12); memcpy(o.h, b.root.h, 12);
        o.rootRef = (b.root.flags & 0x80000000u) ? (0x80000000u | (b.root.flags & 0x3fffffffu)) : 0u;
        o.pad = 0;
        o.pairs = b.pairs; o.tris = b.tris;
    }
    uint32_t depth = 0;
    CUDA_OK(h, build_tlas(instances, blas.data(), n, (uint8_t*)dst, (TlasInstanceRecord*)((uint8_t*)dst + L.records), scratch, &depth, stream, h->lc));
    if (depth > TB_TLAS_STACK_DEPTH) return fail(h, TB_ERR_NOT_IMPL, "top-level tree deeper than the query's stack");
    TlasTrailer t;
    memset(&t, 0, sizeof(t));
    memcpy(t.magic, "TBTL0001", 8);
    t.numInstances = n; t.depth = depth;
    CUDA_OK(h, cudaMemcpyAsync((uint8_t*)dst + L.trailer, &t, sizeof(t), cudaMemcpyHostToDevice, stream));
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
