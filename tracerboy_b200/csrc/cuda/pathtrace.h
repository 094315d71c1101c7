// pathtrace.h — host-visible interface of the CUDA translation units.
#pragma once
#include <vector>
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
    uint32_t* queueCount = nullptr; // TB_QUEUE_COUNT_WORDS words: [16..79] hits per material class, [80..143] the class sort's cursors, [0..1] queue sizes, [2..3] k_extend work counters, [4] shadow queue size, [5] its work counter, [6..9] hit/miss queue sizes, [10..11] walk queue sizes, [12..13] their work counters
    // the bounce's paths sorted by k_extend into "hit something" / "left the scene"; counters [6..9] by queue parity
    uint32_t* hitQueue = nullptr;   // entries: pixel | material class << 26
    uint32_t* hitSorted = nullptr;  // the hit queue grouped by material class (k_class_scatter), plain pixels
    uint32_t* missQueue = nullptr;
    // next-event shadow rays: queued by k_shade<0>, traced by k_extend<true>, consumed by k_shade<1>
    uint32_t* shadowQueue = nullptr;
    float4* shRayO = nullptr;
    float4* shRayD = nullptr;
    float4* shHit = nullptr;       // t, b1, b2, bits(primitiveIndex) of the first hit along the shadow feeler
    uint32_t* shHitGeom = nullptr;
    // glass / subsurface walkers: queued by k_shade, advanced one ray at a time by k_extend<WALK> + k_walk_step for
    // the first rounds, the long tail finished by the persistent k_walk. Ping-pong queues: sizes queueCount[10..11],
    // work counters [12..13].
    uint32_t* walkQueue[2] = {nullptr, nullptr};
    float4* walkA = nullptr;       // absorption.xyz, maxTravelDistance
    float4* walkB = nullptr;       // CurrentIOR, NewIOR, roughness, travelDistance of the ray in flight
    // spatial sort of a bounce's ray queue (large scenes): cell key per entry, sorted copy, cell histogram / offsets
    uint32_t* sortKeys = nullptr;
    uint32_t* sortTmp = nullptr;
    uint32_t* sortHist = nullptr;   // TB_SORT_CELLS + 1 words
    // suspended long rays: two ping-pong record buffers, one counter per round
    uint32_t* susBuf[2] = {nullptr, nullptr};
    uint32_t* susCount = nullptr;   // 4 counters
    uint32_t susCapacity = 0;       // records per buffer
    // per-frame staging consumed by k_accumulate (frames in flight finish out of order)
    float4* sample = nullptr;      // (rgb * w, w) after firefly clamp and NaN rejection
    float* sampleSeed = nullptr;   // the path's rand() seed at termination (jittered-buffer coin)
    float4* stEmissive = nullptr;  // w != 0 when written this frame
    float* stDepth = nullptr;      // < 0 when not written this frame
    // ---- everything below is shared by all frame slots ----
    float4* accum = nullptr;       // OutputTexture
    float4* jittered = nullptr;    // JitteredOutputTexture
    float4* aovAlbedo = nullptr;   // AOVCustomOutput
    float4* aovNormal = nullptr;
    float4* aovEmissive = nullptr;
    float4* aovWorldPos[2] = {nullptr, nullptr};
    float* aovDepth = nullptr;
    uint2* primaryHit = nullptr;
    uint2* counters = nullptr;     // per pixel (TrianglesTested, BoxesTested) of the current frame
    unsigned long long* stats = nullptr; // rays, boxes, tris finished in: [0..2] k_extend, [3..5] inside k_shade, [6..8] k_extend_resume
    TbReadbackStats* readbackStats = nullptr;
};

// Optional per-kernel timing (profiling mode): CUDA events recorded on the launching stream
// around every k_extend / k_shade launch; resolved by the caller after a stream sync.
struct KernelTimers {
    enum Tag { EXTEND = 0, SHADE = 1, END = 2, RESUME = 3 };
    KernelTimers() = default;
    KernelTimers(const KernelTimers&) = delete;            // owns its events
    KernelTimers& operator=(const KernelTimers&) = delete;
    KernelTimers(KernelTimers&& o) noexcept : events(std::move(o.events)), tags(std::move(o.tags)), used(o.used) { o.events.clear(); o.used = 0; }
    std::vector<cudaEvent_t> events;
    std::vector<int> tags;   // Tag | bounce << 8
    size_t used = 0;
    cudaEvent_t next(Tag t, int bounce) {
        if (used == events.size()) { cudaEvent_t e; cudaEventCreate(&e); events.push_back(e); tags.push_back(0); }
        tags[used] = (int)t | ((bounce < 31 ? bounce : 31) << 8);
        return events[used++];
    }
    // adds elapsed ms per tag (and per bounce: every stage of bounce b, extend to walk), counts the extend launches
    void resolve(double& extendMs, double& shadeMs, double& resumeMs, uint64_t& extendLaunches, double* bounceMs /*[32]*/, double* bounceExtendMs /*[32]*/) {
        for (size_t i = 0; i + 1 < used; i++) {
            const int tag = tags[i] & 0xff, bounce = tags[i] >> 8;
            if (tag == END) continue;
            float ms = 0;
            cudaEventElapsedTime(&ms, events[i], events[i + 1]);
            if (tag == EXTEND) { extendMs += ms; extendLaunches++; bounceExtendMs[bounce] += ms; }
            else if (tag == RESUME) resumeMs += ms;
            else shadeMs += ms;
            bounceMs[bounce] += ms;
        }
        used = 0;
    }
    ~KernelTimers() { for (auto e : events) cudaEventDestroy(e); }
};

// scheduling knobs; none of them changes a result
#define TB_QUEUE_COUNT_WORDS 144
#define TB_STATS_WORDS 64    // 64-bit words of PathState::stats (layout: flush_stats in pathtrace.cu)
#define TB_SORT_CELLS 32768u // 5 bits per axis of the ray origin inside the scene box

struct RenderOptions {
    int shadowMode = 2; // next-event shadow rays: 0 traced inline in k_shade, 1 own wavefront stage, 2 automatic
    uint32_t epoch = 0; // bumped by the host when the scene / BVH / frame buffers are re-created (invalidates captured frame graphs)
    int walkRounds = 1; // glass / subsurface walk: wavefront rounds (k_extend<EXT_WALK> + k_walk_step) before the persistent tail kernel
    bool suspendRays = true; // park rays over budget and resume them in k_extend_resume rounds (set by the host: on with fewer than 8 frames in flight)
    int sortRays = 2;   // spatial sort of the ray queues before traversal: 0 off, bit 0 bounce queue, bit 1 shadow queue (1 or 3), 2 automatic (3 for scenes whose BVH is far beyond L2)
    int materialSort = 2; // hit queue grouped by material class before shading: 0 off, 1 on, 2 automatic (on from four material classes)
    uint32_t sceneMaterialClasses = 1; // distinct material classes reachable from the geometry (host-side count)
    int numSMs = 148;   // of the handle's device (persistent grids are sized from it)
    bool sceneHasSSS = true; // any material with the subsurface flag (selects the k_shade variant with the inline random walk)
};

uint64_t bvh_ref_bytes(uint32_t numPrims);
uint64_t bvh_scratch_bytes(uint32_t numPrims);
uint64_t bvh_update_scratch_bytes(uint32_t numPrims);
// PERFORM_UPDATE: same hierarchy, triangles reloaded from the (moved) geometry into their sorted slots, boxes refitted
cudaError_t update_bvh(const BuildGeometry* d_geoms, uint32_t numPrims, DeviceBvh& bvh, void* scratch, cudaStream_t stream, LaunchCounter& lc);
// top level over instances of bottom-level structures (bvh_build.cu); host arrays in, device structure out
uint64_t tlas_scratch_bytes(uint32_t numInstances);
cudaError_t build_tlas(const TbInstanceDesc* h_instances, const TlasBlasInfo* h_blas, uint32_t numInstances, uint8_t* dst, TlasInstanceRecord* records,
                       void* scratch, uint32_t* depthOut, cudaStream_t stream, LaunchCounter& lc);
cudaError_t build_bvh(const BuildGeometry* d_geoms, const uint32_t* d_triPrefix, uint32_t numGeoms, uint32_t numPrims, int treeletPasses,
                      DeviceBvh& out, void* scratch, cudaStream_t stream, LaunchCounter& lc);
// The captured kernel sequence of one frame on one frame slot (see render_frame).
struct FrameGraph {
    struct Key { uint32_t epoch, width, height, maxBounces, heat, nee, shadowMode, walkRounds, sss, suspend, sort, classSort; }; // no padding: compared with memcmp
    Key key{};
    cudaGraphExec_t exec = nullptr;
    uint64_t launches = 0; // kernels inside the graph
    void reset();
};
cudaError_t render_frame(const DeviceBvh& bvh, const DeviceScene& sc, const FrameConstants& fc, FrameConstants* fcDev, PathState& st,
                         cudaStream_t stream, LaunchCounter& lc, KernelTimers* timers, const RenderOptions& opts, FrameGraph* graph);
cudaError_t accumulate_frame(const FrameConstants& fc, PathState& st, cudaStream_t stream, LaunchCounter& lc);
cudaError_t resolve_rgb(const float4* accum, float* rgb, uint32_t n, cudaStream_t stream, LaunchCounter& lc);
cudaError_t trace_rays(const DeviceBvh& bvh, const TbRay* d_rays, uint64_t n, TbHit* d_hits, int numSMs, cudaStream_t stream, LaunchCounter& lc);

} // namespace tbd
