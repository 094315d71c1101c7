// ref_temporal.cpp — CPU ORACLE (test infrastructure): compiles the reference's own TemporalAccumulationCS.hlsl main()
// (with PlaneIntersection and Tonemap.h's ColorToLuma; pre-passed from the mount into oracle/_ref/temporal_gen.inc by
// prepass.run_temporal) as host C++. What is restated here are the RESOURCES only: Texture2D / RWTexture2D element
// access (out-of-bounds loads return 0, as D3D defines), the constant buffer (TemporalAccumulationSharedShaderStructs.h
// layout, filled as TemporalAccumulationPass.cpp:87-103 fills it), the integer vector types the shader indexes with, and
// SampleLevel(BilinearSampler, uv, 0) with clamp addressing and exact float weights (texture-unit arithmetic has no
// source text to compile; the same pinned choice as oracle/temporal.cpp). Every arithmetic statement of the shader is
// the reference's text. tests/test_cpu_temporal.py requires oracle/temporal.cpp to match this build bit for bit.
#define RC_TEMPORAL 1
#include "hlsl_compat.h"
#include <cstring>
#include <vector>
#include "tracerboy_b200.h"

namespace refcore {

struct uint2;
struct int2 {
    int x, y;
    int2(int x_, int y_) : x(x_), y(y_) {}
    explicit int2(float2 f) : x((int)f.x), y((int)f.y) {} // int2(float2): truncation toward zero
    int2(const uint2& u);
};
struct uint2 {
    uint x, y;
    uint2() : x(0), y(0) {}
    uint2(uint x_, uint y_) : x(x_), y(y_) {}
    uint2(const int2& v) : x((uint)v.x), y((uint)v.y) {}  // uint2 Index = int2(...) + int2(x, y)  (:175)
    operator float2() const { return float2((float)x, (float)y); } // float2(Constants.Resolution) (:168)
};
inline int2::int2(const uint2& u) : x((int)u.x), y((int)u.y) {}
struct uint3 { uint x, y, z; uint2 xy() const { return uint2(x, y); } };
inline int2 operator+(int2 a, int2 b) { return int2(a.x + b.x, a.y + b.y); }
struct bool2 { bool x, y; };
inline bool2 operator>(int2 a, int s) { return bool2{a.x > s, a.y > s}; }
inline bool2 operator>=(float2 a, float s) { return bool2{a.x >= s, a.y >= s}; }
inline bool2 operator<=(float2 a, float s) { return bool2{a.x <= s, a.y <= s}; }
inline bool2 operator&&(bool2 a, bool2 b) { return bool2{a.x && b.x, a.y && b.y}; }
inline bool all(bool2 b) { return b.x && b.y; }
inline bool3 operator!=(float3 a, float s) { return bool3{a.x != s, a.y != s, a.z != s}; }
inline bool any(bool3 b) { return b.x || b.y || b.z; }
inline float2 lerp(float2 a, float2 b, float s) { return float2(lerp(a.x, b.x, s), lerp(a.y, b.y, s)); }

struct SamplerState {};
struct Texture2D {
    const TbFloat4* p = nullptr; int w = 0, h = 0;
    float4 load(int x, int y) const { // out of bounds reads zero
        if (!p || x < 0 || y < 0 || x >= w || y >= h) return float4(0, 0, 0, 0);
        const TbFloat4& v = p[(size_t)y * w + x];
        return float4(v.x, v.y, v.z, v.w);
    }
    float4 operator[](uint2 i) const { return load((int)i.x, (int)i.y); }
    float4 operator[](int2 i) const { return load(i.x, i.y); }
    float4 SampleLevel(SamplerState, float2 uv, float) const { // bilinear, clamp addressing, exact float weights
        if (!p) return float4(0, 0, 0, 0);
        float fx = uv.x * (float)w - 0.5f, fy = uv.y * (float)h - 0.5f;
        float x0f = floor(fx), y0f = floor(fy);
        float tx = fx - x0f, ty = fy - y0f;
        auto cl = [](float f, int n) { int i = (int)f; return i < 0 ? 0 : (i > n - 1 ? n - 1 : i); };
        int x0 = cl(x0f, w), x1 = cl(x0f + 1.0f, w), y0 = cl(y0f, h), y1 = cl(y0f + 1.0f, h);
        float3 a = lerp(load(x0, y0).xyz(), load(x1, y0).xyz(), tx);
        float3 b = lerp(load(x0, y1).xyz(), load(x1, y1).xyz(), tx);
        return float4(lerp(a, b, ty), 0.0f);
    }
};
template <typename T> struct RWTexture2D {
    T* p = nullptr; int w = 0; T sink;
    T& operator[](uint2 i) { return p ? p[(size_t)i.y * w + i.x] : sink; }
};

struct TemporalAccumulationConstants { // TemporalAccumulationSharedShaderStructs.h:6-34
    uint2 Resolution; float CameraFocalDistance; uint IgnoreHistory;
    float3 CameraPosition; float CameraLensHeight;
    float3 CameraLookAt; float HistoryWeight;
    float3 CameraUp; uint OutputMomentInformation;
    float3 CameraRight; uint padding3;
    float3 PrevFrameCameraPosition; uint padding4;
    float3 PrevFrameCameraUp; uint padding5;
    float3 PrevFrameCameraRight; uint padding6;
    float3 PrevFrameCameraLookAt; uint padding7;
};

static thread_local Texture2D TemporalHistory, CurrentFrame, WorldPositionTexture, PreviousFrameWorldPositionTexture, MomentHistory, WorldNormalTexture;
static thread_local SamplerState BilinearSampler;
static thread_local RWTexture2D<float4> OutputTexture;
static thread_local RWTexture2D<float3> OutputMoment;
static thread_local TemporalAccumulationConstants Constants;
#include "../_ref/temporal_gen.inc"

} // namespace refcore

// Same signature as oracle_temporal_accumulate_image.
extern "C" __attribute__((visibility("default")))
int ref_temporal_accumulate_image(const TbTemporalAccumulationParams* P, uint32_t width, uint32_t height,
                                  const TbFloat4* history, const TbFloat4* current, const TbFloat4* worldPos,
                                  const TbFloat4* prevWorldPos, const TbFloat4* normals, const TbFloat4* momentHistory,
                                  TbFloat4* outColor, TbFloat4* outMoment) {
    using namespace refcore;
    const int W = (int)width, H = (int)height;
    auto F3 = [](const TbFloat3& v) { return float3(v.x, v.y, v.z); };
    std::vector<float4> color((size_t)W * H);
    std::vector<float3> moment((size_t)W * H);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        TemporalHistory = Texture2D{history, W, H}; CurrentFrame = Texture2D{current, W, H};
        WorldPositionTexture = Texture2D{worldPos, W, H}; PreviousFrameWorldPositionTexture = Texture2D{prevWorldPos, W, H};
        MomentHistory = Texture2D{momentHistory, W, H}; WorldNormalTexture = Texture2D{normals, W, H};
        OutputTexture.p = color.data(); OutputTexture.w = W;
        OutputMoment.p = moment.data(); OutputMoment.w = W;
        // TemporalAccumulationPass.cpp:87-103
        Constants.Resolution = uint2(width, height);
        Constants.CameraFocalDistance = P->Camera.FocalDistance;
        Constants.IgnoreHistory = P->IgnoreHistory;
        Constants.CameraPosition = F3(P->Camera.Position); Constants.CameraLensHeight = P->Camera.LensHeight;
        Constants.CameraLookAt = F3(P->Camera.LookAt); Constants.HistoryWeight = P->HistoryWeight;
        Constants.CameraUp = F3(P->Camera.Up); Constants.OutputMomentInformation = P->OutputMomentInformation;
        Constants.CameraRight = F3(P->Camera.Right);
        Constants.PrevFrameCameraPosition = F3(P->PrevCamera.Position); Constants.PrevFrameCameraUp = F3(P->PrevCamera.Up);
        Constants.PrevFrameCameraRight = F3(P->PrevCamera.Right); Constants.PrevFrameCameraLookAt = F3(P->PrevCamera.LookAt);
        for (int x = 0; x < W; x++) shader_main(uint3{(uint)x, (uint)y, 0u});
    }
    for (size_t i = 0; i < (size_t)W * H; i++) {
        outColor[i] = TbFloat4{color[i].x, color[i].y, color[i].z, color[i].w};
        if (outMoment && P->OutputMomentInformation) outMoment[i] = TbFloat4{moment[i].x, moment[i].y, moment[i].z, 0.0f};
    }
    return 0;
}
