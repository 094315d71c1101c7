// ref_hist.cpp — CPU ORACLE (test infrastructure): compiles the reference's own auto-exposure shaders from the mount —
// GenerateHistogramCS.hlsl (LuminanceToHistogramIndex + main(), with Tonemap.h's ColorToLuma) and
// CalculateAveragedLuminanceCS.hlsl (main()) — pre-passed into oracle/_ref/hist_gen.inc / hist_avg.inc by
// prepass.run_hist, as host C++. Both are 16x16 group shaders with groupshared counters, group barriers and interlocked
// adds: a group is 256 host threads, GroupMemoryBarrierWithGroupSync is a barrier that threads which returned early
// (pixels outside the image) drop out of, InterlockedAdd is an atomic add. Restated here: the resources (Texture2D
// element access, the byte-address buffers), the constant buffers with the values TracerBoy.cpp:2950-2951, 2985-2987
// gives them, and D3D's unsigned division (x / 0 = 0xffffffff, which the averaging shader reaches on an all-black
// image). tests/test_cpu_postprocess.py requires oracle/postprocess.cpp to match this build exactly.
#define RC_POST 1
#include "hlsl_compat.h"
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#include "tracerboy_b200.h"

namespace refcore {

#define NUM_HISTOGRAM_BINS 256          // HistogramSharedShaderStructs.h:3-4
#define BYTE_ADDRESS_BUFFER_STRIDE 4
#define groupshared static

struct GroupBarrier {
    std::mutex m; std::condition_variable cv; int expected = 0, waiting = 0; unsigned long generation = 0;
    void reset(int n) { expected = n; waiting = 0; }
    void wait() {
        std::unique_lock<std::mutex> l(m);
        unsigned long g = generation;
        if (++waiting == expected) { waiting = 0; generation++; cv.notify_all(); }
        else cv.wait(l, [&] { return generation != g; });
    }
    void drop() { // a thread that left main(): the others no longer wait for it
        std::unique_lock<std::mutex> l(m);
        expected--;
        if (expected > 0 && waiting == expected) { waiting = 0; generation++; cv.notify_all(); }
    }
};
static GroupBarrier g_barrier;
inline void GroupMemoryBarrierWithGroupSync() { g_barrier.wait(); }
inline void InterlockedAdd(uint& dst, uint v) { __atomic_fetch_add(&dst, v, __ATOMIC_SEQ_CST); }
inline float log2(float x) { return tbm::log2_(x); }
inline float exp2(float x) { return tbm::exp2_(x); }
inline uint asuint(float f) { uint u; memcpy(&u, &f, 4); return u; }

struct uint2 { uint x, y; };
struct uint3 { uint x, y, z; uint2 xy() const { return uint2{x, y}; } };
struct Texture2D {
    const TbFloat4* p = nullptr; int w = 0, h = 0;
    float4 operator[](uint2 i) const {
        if (!p || (int)i.x >= w || (int)i.y >= h) return float4(0, 0, 0, 0);
        const TbFloat4& v = p[(size_t)i.y * w + i.x];
        return float4(v.x, v.y, v.z, v.w);
    }
};
struct ByteAddressBufferShim {
    uint* words = nullptr;
    void InterlockedAdd(uint byteOffset, uint v) { __atomic_fetch_add(&words[byteOffset / 4], v, __ATOMIC_SEQ_CST); }
    uint Load(uint byteOffset) const { return words[byteOffset / 4]; }
    void Store(uint byteOffset, uint v) { words[byteOffset / 4] = v; }
};

// run one 16x16 group: thread t gets SV_GroupIndex t; f(t) calls the shader's main() for that thread
template <typename F> static void run_group(F f) {
    g_barrier.reset(256);
    std::vector<std::thread> group;
    for (uint t = 0; t < 256; t++) group.emplace_back([t, &f] { f(t); g_barrier.drop(); });
    for (auto& th : group) th.join();
}

namespace gen {
struct GenerateHistogramConstants { uint2 Resolution; float minLogLuminance; float oneOverLogLuminanceRange; }; // GenerateHistogramSharedShaderStructs.h:9-14
static Texture2D InputTexture;
static ByteAddressBufferShim LuminanceHistogram;
static GenerateHistogramConstants Constants;
#include "../_ref/hist_gen.inc"
} // namespace gen

namespace avg {
// unsigned integer with D3D's division: x / 0 = 0xffffffff
struct d3d_uint {
    uint v;
    operator float() const { return (float)v; }
};
inline d3d_uint operator-(d3d_uint a, uint b) { return d3d_uint{a.v - b}; }
inline d3d_uint operator/(uint a, d3d_uint b) { return d3d_uint{b.v ? a / b.v : 0xffffffffu}; }
inline float operator-(d3d_uint a, float b) { return (float)a.v - b; }
struct CalculateAveragedLuminanceConstants { d3d_uint PixelCount; float LogLuminanceRange; float MinLogLuminance; }; // CalculateAveragedLuminanceSharedShaderStructs.h:8-13
static ByteAddressBufferShim LuminanceHistogram, AveragedLuminance;
static CalculateAveragedLuminanceConstants Constants;
#include "../_ref/hist_avg.inc"
} // namespace avg

} // namespace refcore

static std::mutex g_serial; // the shims are process-wide statics

extern "C" __attribute__((visibility("default")))
int ref_luminance_histogram(const TbFloat4* in, uint32_t width, uint32_t height, uint32_t hist[256]) {
    using namespace refcore;
    std::lock_guard<std::mutex> l(g_serial);
    memset(hist, 0, 256 * sizeof(uint32_t));
    gen::InputTexture = Texture2D{in, (int)width, (int)height};
    gen::LuminanceHistogram.words = hist;
    gen::Constants.Resolution = uint2{width, height};
    gen::Constants.minLogLuminance = -10.0f;              // TracerBoy.cpp:2950
    gen::Constants.oneOverLogLuminanceRange = 1.0f / 16.0f; // TracerBoy.cpp:2951 (range 16)
    for (uint32_t gy = 0; gy < (height + 15) / 16; gy++)   // dispatch: TracerBoy.cpp:2966-2970
        for (uint32_t gx = 0; gx < (width + 15) / 16; gx++)
            run_group([&](uint t) { gen::shader_main(uint3{gx * 16 + (t & 15), gy * 16 + (t >> 4), 0u}, t); });
    return 0;
}

extern "C" __attribute__((visibility("default")))
int ref_averaged_luminance(const uint32_t hist[256], uint32_t pixelCount, float* out) {
    using namespace refcore;
    std::lock_guard<std::mutex> l(g_serial);
    uint32_t result = 0;
    avg::LuminanceHistogram.words = const_cast<uint32_t*>(hist);
    avg::AveragedLuminance.words = &result;
    avg::Constants.PixelCount = avg::d3d_uint{pixelCount};
    avg::Constants.LogLuminanceRange = 16.0f;  // TracerBoy.cpp:2986
    avg::Constants.MinLogLuminance = -10.0f;   // TracerBoy.cpp:2985
    run_group([&](uint t) { avg::shader_main(t); });
    memcpy(out, &result, 4);
    return 0;
}
