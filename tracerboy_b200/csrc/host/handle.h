// handle.h — the object behind the opaque TbHandle* of the C ABI (private to csrc/host).
#pragma once
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../cuda/pathtrace.h"
#include "../cuda/postprocess.h"
#include "scene.h"
#include "tracerboy_b200.h"

struct TbHandle;
namespace tbh {
// Multi-GPU state of a handle (comm.cpp): one NCCL communicator, the all-gather staging and the job-wide result.
struct Comm {
    void* nccl = nullptr;            // ncclComm_t
    int rank = 0, nranks = 1;
    uint32_t mode = 0;               // TB_SHARD_SAMPLES / TB_SHARD_ROWS
    float4* gather = nullptr;        // what the all-gather receives (N whole buffers, or N packed band chunks)
    float4* pack = nullptr;          // row bands: this rank's bands of accum | jittered, contiguous
    float4* reducedAccum = nullptr;  // OutputTexture of the whole job
    float4* reducedJittered = nullptr;
    size_t pixels = 0;               // resolution the buffers above were allocated for
    bool valid = false;              // the reduced buffers reflect everything rendered so far
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint64_t reductions = 0, bytesPerReduction = 0;
    double lastMs = 0.0, totalMs = 0.0;
    // peer-memory transport: every rank's accumulation and job-wide buffers mapped into this process (CUDA IPC over NVLink)
    bool peer = false;
    float4** peerTable = nullptr;    // device: [4][nranks] pointers (reduce.h)
    std::vector<void*> imported;     // the other ranks' buffers as opened here (cudaIpcCloseMemHandle on release)
    uint32_t* barrier = nullptr;     // device: nranks words, the tiny all-gather that orders the ranks' streams
};
int resolve_bottom_level(TbHandle* h, const void* as, cudaStream_t stream, tbd::DeviceBvh& out); // api.cpp
void comm_release_buffers(TbHandle* h);
void comm_destroy(TbHandle* h);
}
using namespace tbd;

struct TbHandle {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    tb::Scene scene;
    bool sceneLoaded = false;
    uint32_t instanceMode = 0;      // TB_INSTANCES_*: what the PBRT import does with object instances
    // device scene
    DeviceScene dscene;
    std::vector<void*> sceneAllocs;
    DeviceBvh bvh;
    double bvhBuildMs = 0.0;
    // render state
    PathState st;                   // shared buffers (+ slot 0's private ones)
    // Frames in flight: each slot owns private path state + staging and its own stream, so the
    // long tail of one frame's traversal kernels overlaps the next frames' work. h->stream is
    // the accumulate stream: k_accumulate runs there in frame order.
    struct Slot { PathState st; cudaStream_t stream = nullptr; cudaEvent_t frameDone = nullptr, accDone = nullptr;
                  FrameConstants* fcDev = nullptr; FrameGraph graph; KernelTimers timers; };
    std::vector<Slot> slots;
    uint32_t framesInFlight = 0;    // 0 = automatic (memory budget), see tb_resize
    uint64_t framesIssued = 0;
    std::vector<void*> frameAllocs;
    float* resolved = nullptr;
    float4* post = nullptr;         // PostProcessCS output (float4) ...
    uchar4* post8 = nullptr;        // ... and after the UNORM8 back-buffer store
    uint32_t* lumHist = nullptr;    // LuminanceHistogram[256] + AveragedLuminance
    int numSMs = 148;
    uint32_t width = 0, height = 0;
    TbCamera camera{};
    uint32_t samplesRendered = 0; // local samples since the last invalidate
    uint32_t shardOffset = 0, shardStride = 1;
    uint32_t rowOffset = 0, rowStride = 1;
    int selX = -1, selY = -1;
    uint32_t lastMouse[2] = {0, 0}; // m_mouseX, m_mouseY (TracerBoy.cpp:511-512)
    LaunchCounter lc;
    double deviceMs = 0.0;
    uint64_t pathsStarted = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::chrono::steady_clock::time_point renderStart;
    bool timing = false;
    std::mutex statusLock;
    TbSceneLoadStatus status{TB_LOAD_IDLE, 0, 0};
    std::vector<uint8_t> blueNoiseHost;
    int profiling = 0;              // 0 off, 1 one frame at a time (exclusive kernel times), 2 frames in flight (in-situ shares)
    void* buildScratch = nullptr;   // the builder's temporaries, kept between builds
    uint64_t buildScratchBytes = 0;
    std::map<const void*, DeviceBvh> deviceBuilds; // acceleration structures built into caller memory (tb_bvh_build_device)
    std::map<const void*, uint32_t> topLevelBuilds; // top-level structures in caller memory -> number of instances
    RenderOptions options;
    double extendMs = 0.0, shadeMs = 0.0, resumeMs = 0.0;
    double bounceMs[32] = {}, bounceExtendMs[32] = {};
    uint64_t extendLaunches = 0;
    tbh::Comm* comm = nullptr;      // multi-GPU: NCCL communicator + reduction buffers (comm.cpp), nullptr on one GPU
};

namespace tbh {
std::string& create_error(); // message of a failed tb_create (there is no handle to hold it)
inline int fail(TbHandle* h, int code, const std::string& msg) {
    if (h) h->err = msg; else create_error() = msg;
    return code;
}
} // namespace tbh
using tbh::fail;
#define CUDA_OK(h, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail(h, TB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } while (0)

