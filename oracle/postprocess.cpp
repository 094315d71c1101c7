// postprocess.cpp — CPU ORACLE (test infrastructure): the step right after the path.
//
// Hand restatement of
//   GenerateHistogramCS.hlsl:18-55          luminance histogram, 256 bins over log2 luminance
//   CalculateAveragedLuminanceCS.hlsl:15-42 histogram -> averaged luminance (note the *integer* division)
//   PostProcessCS.hlsl:23-196               per-OutputType resolve, exposure, tonemap, gamma
//   Tonemap.h:12-211                        Reinhard / ACES / clamp / Uncharted2 / Khronos PBR neutral / AgX / AgX punchy / GT
// with the host constants of TracerBoy.cpp:2948-3039 (MinLogLuminance -10, LogLuminanceRange 16).
//
// PIN: oracle/ref/ref_post.cpp compiles Tonemap.h and the Process* functions of PostProcessCS.hlsl from the
// reference mount as host C++ (oracle/_ref/libref_post.so); tests/test_cpu_postprocess.py requires this
// restatement to match that build bit for bit on random and rendered inputs -> PINNED against the
// reference's own text. The two histogram shaders are group shaders with resources: restated here and pinned against
// their own text run by a 256-thread host group (oracle/ref/ref_hist.cpp, tests/test_cpu_postprocess.py).
#include <cmath>
#include <cstring>
#include <vector>
#include "../tracerboy_b200/csrc/common/tb_vec.h"
#include "oracle.h"

using namespace tbm;

namespace oracle {
namespace {

struct M3 { f3 r0, r1, r2; };
inline f3 mul_mv(const M3& m, f3 v) { return mk3(dot(m.r0, v), dot(m.r1, v), dot(m.r2, v)); }          // mul(M, v)
inline f3 mul_vm(f3 v, const M3& m) { return (v.x * m.r0 + v.y * m.r1) + v.z * m.r2; }                  // mul(v, M)
inline f3 sat3(f3 v) { return mk3(saturate(v.x), saturate(v.y), saturate(v.z)); }

inline float ColorToLuma(f3 c) { return dot(c, mk3(0.212671f, 0.715160f, 0.072169f)); }                 // Tonemap.h:12-15
inline f3 GammaCorrect(f3 c) { return pow3(c, 1.0f / 2.2f); }                                             // :147-150

f3 RRTAndODTFit(f3 v) {                                                                                   // :33-38
    f3 a = v * (v + 0.0245786f) - 0.000090537f;
    f3 b = v * (0.983729f * v + 0.4329510f) + 0.238081f;
    return a / b;
}
f3 ACESFitted(f3 color) {                                                                                 // :40-53
    const M3 in = {mk3(0.59719f, 0.35458f, 0.04823f), mk3(0.07600f, 0.90834f, 0.01566f), mk3(0.02840f, 0.13383f, 0.83777f)};
    const M3 out = {mk3(1.60475f, -0.53108f, -0.07367f), mk3(-0.10208f, 1.10813f, -0.00605f), mk3(-0.00327f, -0.07276f, 1.07602f)};
    color = mul_mv(in, color);
    color = RRTAndODTFit(color);
    color = mul_mv(out, color);
    return sat3(color);
}
f3 Reinhard(f3 c) { return c / (1.0f + c); }                                                              // :55-59
f3 uncharted2_partial(f3 x) {                                                                             // :62-71
    float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    return ((x * (A * x + C * B) + D * E) / (x * (A * x + B) + D * F)) - E / F;
}
f3 uncharted2_filmic(f3 v) {                                                                              // :74-82
    float exposure_bias = 2.0f;
    f3 curr = uncharted2_partial(v * exposure_bias);
    f3 W = mk3(11.2f, 11.2f, 11.2f);
    f3 white_scale = mk3(1.0f, 1.0f, 1.0f) / uncharted2_partial(W);
    return curr * white_scale;
}
f3 CommerceToneMapping(f3 color) {                                                                        // :85-102
    float startCompression = 0.8f - 0.04f;
    float desaturation = 0.15f;
    float x = fminf(color.x, fminf(color.y, color.z));
    float offset = x < 0.08f ? x - 6.25f * x * x : 0.04f;
    color = color - offset;
    float peak = fmaxf(color.x, fmaxf(color.y, color.z));
    if (peak < startCompression) return color;
    float d = 1.0f - startCompression;
    float newPeak = 1.0f - d * d / (peak + d - startCompression);
    color = color * (newPeak / peak);
    float g = 1.0f - 1.0f / (desaturation * (peak - newPeak) + 1.0f);
    return mk3(lerp(color.x, newPeak * 1.0f, g), lerp(color.y, newPeak * 1.0f, g), lerp(color.z, newPeak * 1.0f, g));
}
f3 agxContrast(f3 x) {                                                                                    // :104-108
    f3 x2 = x * x;
    f3 x4 = x2 * x2;
    return +15.5f * x4 * x2 - 40.14f * x4 * x + 31.96f * x4 - 6.868f * x2 * x + 0.4298f * x2 + 0.1191f * x - 0.00232f;
}
f3 agx(f3 color) {                                                                                        // :110-123
    const M3 t = {mk3(0.842479062253094f, 0.0423282422610123f, 0.0423756549057051f),
                  mk3(0.0784335999999992f, 0.878468636469772f, 0.0784336f),
                  mk3(0.0792237451477643f, 0.0791661274605434f, 0.879142973793104f)};
    const float minEv = -12.47393f, maxEv = 4.026069f;
    color = mul_vm(color, t);
    color = mk3(clamp_(log2_(color.x), minEv, maxEv), clamp_(log2_(color.y), minEv, maxEv), clamp_(log2_(color.z), minEv, maxEv));
    color = (color - minEv) / (maxEv - minEv);
    return agxContrast(color);
}
f3 agxLook(f3 val, bool punchy) {                                                                         // :126-145
    const f3 lw = mk3(0.2126f, 0.7152f, 0.0722f);
    float luma = dot(val, lw);
    f3 offset = mk3(0.0f, 0.0f, 0.0f);
    f3 slope = mk3(1.0f, 1.0f, 1.0f);
    float power = 1.0f, sat = 1.0f;
    if (punchy) { power = 1.35f; sat = 1.4f; }
    val = pow3(val * slope + offset, power);
    return luma + sat * (val - luma);
}
float GTTonemap(float x) {                                                                                // :152-171
    float m = 0.22f, a = 1.0f, c = 1.33f, P = 1.0f, l = 0.4f;
    float l0 = ((P - m) * l) / a;
    float S0 = m + l0;
    float S1 = m + a * l0;
    float C2 = (a * P) / (P - S1);
    float L = m + a * (x - m);
    float T = m * pow_(x / m, c);
    float S = P - (P - S1) * exp_(-C2 * (x - S0) / P);
    float w0 = 1.0f - smoothstep_(0.0f, m, x);
    float w2 = (x < m + l) ? 0.0f : 1.0f;
    float w1 = 1.0f - w0 - w2;
    return T * w0 + L * w1 + S * w2;
}
f3 Tonemap(uint32_t type, f3 color) {                                                                     // :173-204
    switch (type) {
    case 0: return GammaCorrect(Reinhard(color));
    case 7: return GammaCorrect(mk3(GTTonemap(color.x), GTTonemap(color.y), GTTonemap(color.z)));
    case 1: return GammaCorrect(ACESFitted(color));
    case 3: return GammaCorrect(uncharted2_filmic(color));
    case 4: return GammaCorrect(CommerceToneMapping(color));
    case 5: return agxLook(agx(color), false);
    case 6: return agxLook(agx(color), true);
    default: return GammaCorrect(sat3(color));
    }
}

f3 xyz(const TbFloat4& c) { return mk3(c.x, c.y, c.z); }

f3 ProcessLit(const TbFloat4& color, const TbPostProcessSettings& C, float averagedLuminance) {           // PostProcessCS.hlsl:23-47
    float FrameCount = color.w;
    f3 outputColor = xyz(color) / FrameCount;
    float Exposure;
    if (C.UseAutoExposure) {
        float LinearGray = pow_(0.5f, 2.2f);
        Exposure = LinearGray / averagedLuminance;
    } else Exposure = C.ExposureMultiplier;
    outputColor = outputColor * Exposure;
    return Tonemap(C.TonemapType, outputColor);
}
f3 Lerp3(f3 c0, f3 c1, f3 c2, float v) {                                                                  // :126-136
    if (v < 0.5f) { float s = v * 2.0f; return mk3(lerp(c0.x, c1.x, s), lerp(c0.y, c1.y, s), lerp(c0.z, c1.z, s)); }
    float s = (v - 0.5f) * 2.0f;
    return mk3(lerp(c1.x, c2.x, s), lerp(c1.y, c2.y, s), lerp(c1.z, c2.z, s));
}

} // namespace

uint32_t luminance_to_histogram_index(float luminance) {                                                  // GenerateHistogramCS.hlsl:18-30
    const float epsilon = 0.00001f;
    if (luminance < epsilon) return 0;
    const float minLog = -10.0f, oneOverRange = 1.0f / 16.0f;  // TracerBoy.cpp:2950-2951, 2985-2986
    float logLuminance = saturate((log2_(luminance) - minLog) * oneOverRange);
    return (uint32_t)(logLuminance * 254.0f + 1.0f);
}

// Each 16x16 group builds a groupshared histogram and thread Gid adds bin Gid to the global one at the end -- but the
// threads outside the image have returned by then (:39), so in a partial group at the right / bottom edge the bins
// whose thread lies outside the image are never added: a pixel is counted iff thread `bin` of its group exists.
// (1920x1080: the last row of groups has 8 rows of threads, its bins 128..255 -- luminance above 0.25 -- are dropped.)
void luminance_histogram(const TbFloat4* in, uint32_t width, uint32_t height, uint32_t hist[256]) {     // :32-55
    memset(hist, 0, 256 * sizeof(uint32_t));
    for (uint32_t y = 0; y < height; y++)
        for (uint32_t x = 0; x < width; x++) {
            const TbFloat4& p = in[(size_t)y * width + x];
            f3 Color = xyz(p) / p.w;
            uint32_t bin = luminance_to_histogram_index(ColorToLuma(Color));
            bool flushed = (x & ~15u) + (bin & 15u) < width && (y & ~15u) + (bin >> 4) < height;
            if (flushed) hist[bin]++;
        }
}

float averaged_luminance(const uint32_t hist[256], uint32_t pixelCount) {                                 // CalculateAveragedLuminanceCS.hlsl:15-42
    uint32_t sum = 0;
    for (uint32_t b = 0; b < 256; b++) sum += hist[b] * b;  // uint arithmetic, wraps like InterlockedAdd
    uint32_t denom = pixelCount - hist[0];                  // the first thread's BinCount is bin 0
    uint32_t q = denom ? sum / denom : 0xffffffffu;         // D3D: unsigned division by zero yields 0xffffffff
    float averagedLogLuminance = ((float)q - 1.0f) / 254.0f;
    return exp2_(averagedLogLuminance * 16.0f + -10.0f);
}

TbFloat4 postprocess_pixel(const TbFloat4& colorData, const TbFloat4& auxData, uint32_t outputType, uint32_t width, uint32_t height,
                           const TbPostProcessSettings& C, float averagedLuminance) {                     // PostProcessCS.hlsl:149-196
    f3 o;
    switch (outputType) {
    default: o = ProcessLit(colorData, C, averagedLuminance); break;
    case TB_OUTPUT_ALBEDO: {
        o = Tonemap(C.TonemapType, xyz(colorData) * C.ExposureMultiplier);
        if (C.UseGammaCorrection) o = GammaCorrect(o);
        break;
    }
    case TB_OUTPUT_NORMALS: {
        uint32_t frameCount = (uint32_t)colorData.w;
        o = frameCount > 0 ? abs3(normalize(xyz(colorData) / (float)frameCount)) : mk3(0.0f, 0.0f, 0.0f);
        break;
    }
    case TB_OUTPUT_DEPTH:
    case TB_OUTPUT_LIVE_PIXELS: o = Tonemap(C.TonemapType, xyz(colorData) * C.ExposureMultiplier); break;
    case TB_OUTPUT_MOTION_VECTORS: {
        o = mk3(colorData.x / (float)width, colorData.y / (float)height, 0.0f);
        if (C.UseGammaCorrection) o = GammaCorrect(o);
        break;
    }
    case TB_OUTPUT_LUMINANCE: {
        f3 c = xyz(colorData) / colorData.w;
        c = c * C.ExposureMultiplier;
        c = Tonemap(C.TonemapType, c);
        o = mk3(ColorToLuma(c));
        if (C.UseGammaCorrection) o = GammaCorrect(o);
        break;
    }
    case TB_OUTPUT_LUMINANCE_VARIANCE: o = C.VarianceMultiplier * mk3(colorData.x, 0.0f, 0.0f); break;
    case TB_OUTPUT_LIVE_WAVES: {
        o = ProcessLit(colorData, C, averagedLuminance);
        if (auxData.x > 0.1f || auxData.y > 0.1f || auxData.z > 0.1f || auxData.w > 0.1f) o = xyz(auxData);
        break;
    }
    case TB_OUTPUT_HEATMAP: {
        uint32_t TrianglesTested = (uint32_t)colorData.x, BoxesTested = (uint32_t)colorData.y;
        uint32_t TotalTests = TrianglesTested + BoxesTested;
        float lerpValue = (float)TotalTests / 100.0f;
        o = Lerp3(mk3(0.0f, 1.0f, 0.0f), mk3(1.0f, 1.0f, 0.0f), mk3(1.0f, 0.0f, 0.0f), lerpValue);
        o = Tonemap(C.TonemapType, o * C.ExposureMultiplier);
        break;
    }
    }
    return TbFloat4{o.x, o.y, o.z, 1.0f};
}

// float -> UNORM8 as a typed UAV store does it (D3D11.3 functional spec 3.2.3.6: NaN -> 0, clamp, *255, +0.5, truncate)
uint8_t float_to_unorm8(float c) {
    if (c != c) return 0;
    c = fminf(fmaxf(c, 0.0f), 1.0f);
    return (uint8_t)(uint32_t)(c * 255.0f + 0.5f);
}

} // namespace oracle

#define ORACLE_API extern "C" __attribute__((visibility("default")))

// in / aux: float4 images (aux may be null -> zeros). out: float4, rgba8: 4 bytes per pixel (may be null),
// hist: 256 words (may be null), avgLum: 1 float (may be null). The histogram pass runs only with auto exposure,
// as in TracerBoy.cpp:2948.
ORACLE_API int oracle_postprocess_image(const TbFloat4* in, const TbFloat4* aux, uint32_t width, uint32_t height, uint32_t outputType,
                                        const TbPostProcessSettings* C, TbFloat4* out, uint8_t* rgba8, uint32_t* hist, float* avgLum) {
    using namespace oracle;
    size_t n = (size_t)width * height;
    uint32_t h[256];
    memset(h, 0, sizeof(h));
    float avg = 0.0f;
    if (C->UseAutoExposure) {
        luminance_histogram(in, width, height, h);
        avg = averaged_luminance(h, (uint32_t)n);
    }
    if (hist) memcpy(hist, h, sizeof(h));
    if (avgLum) *avgLum = avg;
    const TbFloat4 zero{0, 0, 0, 0};
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        TbFloat4 o = postprocess_pixel(in[i], aux ? aux[i] : zero, outputType, width, height, *C, avg);
        out[i] = o;
        if (rgba8) { rgba8[4 * i] = float_to_unorm8(o.x); rgba8[4 * i + 1] = float_to_unorm8(o.y); rgba8[4 * i + 2] = float_to_unorm8(o.z); rgba8[4 * i + 3] = float_to_unorm8(o.w); }
    }
    return 0;
}
