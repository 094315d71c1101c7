// ref_post.cpp — CPU ORACLE (test infrastructure): compiles the reference's own TracerBoy/Tonemap.h and the
// Process* functions of TracerBoy/PostProcessCS.hlsl (pre-passed from the mount into oracle/_ref/post_gen.inc by
// prepass.run_post) as host C++. What is restated here: the resource declarations (cbuffer, ByteAddressBuffer)
// and main()'s OutputType switch (PostProcessCS.hlsl:139-196), which only dispatches to the compiled functions.
// tests/test_cpu_postprocess.py requires oracle/postprocess.cpp to match this build bit for bit.
#define RC_POST 1
#include "hlsl_compat.h"
#include <cstring>
#include "tracerboy_b200.h"

namespace refcore {

struct uint2 { uint x, y; operator float2() const { return float2((float)x, (float)y); } }; // float2(Constants.Resolution), :110
struct float3x3 {
    float3 r0, r1, r2;
    float3x3(float3 a, float3 b, float3 c) : r0(a), r1(b), r2(c) {}
    float3x3(float a, float b, float c, float d, float e, float f, float g, float h, float i) : r0(a, b, c), r1(d, e, f), r2(g, h, i) {}
};
inline float3 mul(const float3x3& m, float3 v) { return float3(dot(m.r0, v), dot(m.r1, v), dot(m.r2, v)); }
inline float3 mul(float3 v, const float3x3& m) { return (v.x * m.r0 + v.y * m.r1) + v.z * m.r2; }
struct bool4 { bool x, y, z, w; };
inline bool4 operator>(float4 a, float s) { return bool4{a.x > s, a.y > s, a.z > s, a.w > s}; }
inline bool any(bool4 b) { return b.x || b.y || b.z || b.w; }
inline float log2(float x) { return tbm::log2_(x); }
inline float smoothstep(float a, float b, float x) { return tbm::smoothstep_(a, b, x); }
inline float3 log2(float3 v) { return float3(log2(v.x), log2(v.y), log2(v.z)); }
inline float3 clamp(float3 v, float lo, float hi) { return float3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }
inline float3 saturate(float3 v) { return float3(saturate(v.x), saturate(v.y), saturate(v.z)); }
inline float3 pow(float3 v, float3 e) { return float3(pow(v.x, e.x), pow(v.y, e.y), pow(v.z, e.z)); }
inline float3 pow(float3 v, float4 e) { return float3(pow(v.x, e.x), pow(v.y, e.y), pow(v.z, e.z)); } // implicit truncation, Tonemap.h:149
inline float asfloat(uint u) { float f; memcpy(&f, &u, 4); return f; }

struct PostProcessConstants { // SharedPostProcessStructs.h:3-14
    uint2 Resolution; uint FramesRendered; float ExposureMultiplier; uint TonemapType, UseGammaCorrection, UseAutoExposure, OutputType;
    float VarianceMultiplier;
};
struct ByteAddressBufferShim { uint word; uint Load(uint) const { return word; } };
static thread_local PostProcessConstants Constants;
static thread_local ByteAddressBufferShim AveragedLuminance;
#include "../_ref/post_gen.inc"

static float3 run_pixel(float4 colorData, float4 auxData) { // main(), PostProcessCS.hlsl:149-195
    float3 outputColor;
    switch (Constants.OutputType) {
    case TB_OUTPUT_LIT: default: outputColor = ProcessLit(colorData); break;
    case TB_OUTPUT_ALBEDO: outputColor = ProcessAlbedo(colorData); break;
    case TB_OUTPUT_NORMALS: outputColor = ProcessNormal(colorData); break;
    case TB_OUTPUT_DEPTH: outputColor = PassThroughColor(colorData); break;
    case TB_OUTPUT_MOTION_VECTORS: outputColor = ProcessMotionVectors(colorData); break;
    case TB_OUTPUT_LUMINANCE: outputColor = ProcessLuminance(colorData); break;
    case TB_OUTPUT_LUMINANCE_VARIANCE: outputColor = ProcessLuminanceVariance(colorData); break;
    case TB_OUTPUT_LIVE_PIXELS: outputColor = PassThroughColor(colorData); break;
    case TB_OUTPUT_LIVE_WAVES: outputColor = ProcessLiveWaves(colorData, auxData); break;
    case TB_OUTPUT_HEATMAP: outputColor = ProcessHeatmap(colorData); break;
    }
    return outputColor;
}

} // namespace refcore

// Same signature as the PostProcessCS part of oracle_postprocess_image; the averaged luminance is an input
// (the histogram shaders are not part of this build).
extern "C" __attribute__((visibility("default")))
int ref_postprocess_image(const TbFloat4* in, const TbFloat4* aux, uint32_t width, uint32_t height, uint32_t outputType,
                          const TbPostProcessSettings* C, float averagedLuminance, TbFloat4* out) {
    using namespace refcore;
    size_t n = (size_t)width * height;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        Constants.Resolution = uint2{width, height};
        Constants.FramesRendered = 0;
        Constants.ExposureMultiplier = C->ExposureMultiplier; Constants.TonemapType = C->TonemapType;
        Constants.UseGammaCorrection = C->UseGammaCorrection; Constants.UseAutoExposure = C->UseAutoExposure;
        Constants.OutputType = outputType; Constants.VarianceMultiplier = C->VarianceMultiplier;
        memcpy(&AveragedLuminance.word, &averagedLuminance, 4);
        float4 a = aux ? float4(aux[i].x, aux[i].y, aux[i].z, aux[i].w) : float4(0, 0, 0, 0);
        float3 o = run_pixel(float4(in[i].x, in[i].y, in[i].z, in[i].w), a);
        out[i] = TbFloat4{o.x, o.y, o.z, 1.0f};
    }
    return 0;
}
