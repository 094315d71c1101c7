// pathtrace.h — host-visible interface of the CUDA translation units.
#pragma once
#include <cuda_runtime.h>
#include "device_types.h"
#include "launch.h"

namespace tbd {

// Wavefront path state, SoA, indexed by pixel (one path per pixel per frame).
struct PathState {
    float4* rayO = nullptr;        // origin.xyz, w = rand() seed (the reference's float counter)
    float4* rayD = nullptr;        // direction.xyz, w = bits: bounce | prevPerfectlySpecular << 8
    float4* thr = nullptr;         // accumulatedIndirectLightMultiplier.xyz, w = filter weight
    float4* col = nullptr;         // accumulatedColor.xyz
    float4* hit = nullptr;         // t, b1, b2, bits(primitiveIndex)
    uint32_t* hitGeom = nullptr;
    float4* neighbor = nullptr;    // neighbour camera ray origin (bounce 0 only)
    float4* neighborDir = nullptr; // neighbour camera ray direction
    uint32_t* queue[2] = {nullptr, nullptr};
    uint32_t* queueCount = nullptr; // 2 counters
    float4* accum = nullptr;       // OutputTexture
    float4* jittered = nullptr;    // JitteredOutputTexture
    float4* aovAlbedo = nullptr;   // AOVCustomOutput
    float4* aovNormal = nullptr;
    float4* aovEmissive = nullptr;
    float4* aovWorldPos[2] = {nullptr, nullptr};
    float* aovDepth = nullptr;
    uint2* primaryHit = nullptr;
    uint2* counters = nullptr;     // per pixel (TrianglesTested, BoxesTested) of the current frame
    unsigned long long* stats = nullptr; // rays, boxes, tris (cumulative)
    TbReadbackStats* readbackStats = nullptr;
};

uint64_t bvh_ref_bytes(uint32_t numPrims);
cudaError_t build_bvh(const TbGeometryRecord* d_geoms, const uint32_t* d_triPrefix, uint32_t numGeoms,
                      const float* d_positions, const uint32_t* d_indices, uint32_t numPrims, int treeletPasses,
                      DeviceBvh& out, cudaStream_t stream, LaunchCounter& lc);
cudaError_t render_frame(const DeviceBvh& bvh, const DeviceScene& sc, const FrameConstants& fc, PathState& st,
                         cudaStream_t stream, LaunchCounter& lc);
cudaError_t resolve_rgb(const float4* accum, float* rgb, uint32_t n, cudaStream_t stream, LaunchCounter& lc);
cudaError_t trace_rays(const DeviceBvh& bvh, const TbRay* d_rays, uint64_t n, TbHit* d_hits, cudaStream_t stream, LaunchCounter& lc);

} // namespace tbd
