// postprocess.cu — the step right after the path for sm_100a.
//
// Role: the auto-exposure passes GenerateHistogramCS.hlsl:18-55 and CalculateAveragedLuminanceCS.hlsl:15-42
// (host side TracerBoy.cpp:2948-3039) and PostProcessCS.hlsl:23-196 with Tonemap.h:12-211.
//
//   k_luminance_histogram  persistent blocks (a multiple of the SM count), grid-stride over pixels, one
//                          256-bin histogram per block in shared memory, 256 global atomics per block.
//                          HBM-bound: 16 B read per pixel.
//   k_postprocess          every block first reduces the 256-bin histogram to the averaged luminance (1 KB,
//                          L2 resident; replaces the reference's single-group CalculateAveragedLuminance dispatch),
//                          then resolves / exposes / tonemaps its pixels. HBM-bound: 16 B read + 20 B written
//                          per pixel (float4 result + the UNORM8 back buffer).
//
// Arithmetic follows the reference's order of operations on the pinned intrinsics (common/tb_math.h, -fmad=false),
// so the float4 result equals the CPU oracle's bit for bit; the histogram is integer work and order independent.
#include "../common/tb_vec.h"
#include "postprocess.h"

using namespace tbm;

namespace tbd {
namespace {

struct M3 { f3 r0, r1, r2; };
__device__ __forceinline__ f3 mul_mv(const M3& m, f3 v) { return mk3(dot(m.r0, v), dot(m.r1, v), dot(m.r2, v)); }
__device__ __forceinline__ f3 mul_vm(f3 v, const M3& m) { return (v.x * m.r0 + v.y * m.r1) + v.z * m.r2; }
__device__ __forceinline__ f3 sat3(f3 v) { return mk3(saturate(v.x), saturate(v.y), saturate(v.z)); }
__device__ __forceinline__ float color_to_luma(f3 c) { return dot(c, mk3(0.212671f, 0.715160f, 0.072169f)); }
__device__ __forceinline__ f3 gamma_correct(f3 c) { return pow3(c, 1.0f / 2.2f); }

__device__ f3 rrt_odt_fit(f3 v) {
    f3 a = v * (v + 0.0245786f) - 0.000090537f;
    f3 b = v * (0.983729f * v + 0.4329510f) + 0.238081f;
    return a / b;
}
__device__ f3 aces_fitted(f3 color) {
    const M3 in = {mk3(0.59719f, 0.35458f, 0.04823f), mk3(0.07600f, 0.90834f, 0.01566f), mk3(0.02840f, 0.13383f, 0.83777f)};
    const M3 out = {mk3(1.60475f, -0.53108f, -0.07367f), mk3(-0.10208f, 1.10813f, -0.00605f), mk3(-0.00327f, -0.07276f, 1.07602f)};
    color = mul_mv(in, color);
    color = rrt_odt_fit(color);
    color = mul_mv(out, color);
    return sat3(color);
}
__device__ f3 uncharted2_partial(f3 x) {
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    return ((x * (A * x + C * B) + D * E) / (x * (A * x + B) + D * F)) - E / F;
}
__device__ f3 uncharted2_filmic(f3 v) {
    f3 curr = uncharted2_partial(v * 2.0f);
    f3 white_scale = mk3(1.0f, 1.0f, 1.0f) / uncharted2_partial(mk3(11.2f, 11.2f, 11.2f));
    return curr * white_scale;
}
__device__ f3 commerce_tonemapping(f3 color) {
    const float startCompression = 0.8f - 0.04f;
    const float desaturation = 0.15f;
    float x = fminf(color.x, fminf(color.y, color.z));
    float offset = x < 0.08f ? x - 6.25f * x * x : 0.04f;
    color = color - offset;
    float peak = fmaxf(color.x, fmaxf(color.y, color.z));
    if (peak < startCompression) return color;
    float d = 1.0f - startCompression;
    float newPeak = 1.0f - d * d / (peak + d - startCompression);
    color = color * (newPeak / peak);
    float g = 1.0f - 1.0f / (desaturation * (peak - newPeak) + 1.0f);
    return lerp(color, mk3(newPeak * 1.0f, newPeak * 1.0f, newPeak * 1.0f), g);
}
__device__ f3 agx_contrast(f3 x) {
    f3 x2 = x * x;
    f3 x4 = x2 * x2;
    return 15.5f * x4 * x2 - 40.14f * x4 * x + 31.96f * x4 - 6.868f * x2 * x + 0.4298f * x2 + 0.1191f * x - 0.00232f;
}
__device__ f3 agx(f3 color) {
    const M3 t = {mk3(0.842479062253094f, 0.0423282422610123f, 0.0423756549057051f),
                  mk3(0.0784335999999992f, 0.878468636469772f, 0.0784336f),
                  mk3(0.0792237451477643f, 0.0791661274605434f, 0.879142973793104f)};
    const float minEv = -12.47393f, maxEv = 4.026069f;
    color = mul_vm(color, t);
    color = mk3(clamp_(log2_(color.x), minEv, maxEv), clamp_(log2_(color.y), minEv, maxEv), clamp_(log2_(color.z), minEv, maxEv));
    color = (color - minEv) / (maxEv - minEv);
    return agx_contrast(color);
}
__device__ f3 agx_look(f3 val, bool punchy) {
    float luma = dot(val, mk3(0.2126f, 0.7152f, 0.0722f));
    const float power = punchy ? 1.35f : 1.0f, sat = punchy ? 1.4f : 1.0f;
    val = pow3(val * mk3(1.0f, 1.0f, 1.0f) + mk3(0.0f, 0.0f, 0.0f), power);
    return luma + sat * (val - luma);
}
__device__ float gt_tonemap(float x) {
    const float m = 0.22f, a = 1.0f, c = 1.33f, P = 1.0f, l = 0.4f;
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
__device__ f3 tonemap(uint32_t type, f3 color) {
    switch (type) {
    case TB_TONEMAP_REINHARD: return gamma_correct(color / (1.0f + color));
    case TB_TONEMAP_GT: return gamma_correct(mk3(gt_tonemap(color.x), gt_tonemap(color.y), gt_tonemap(color.z)));
    case TB_TONEMAP_ACES: return gamma_correct(aces_fitted(color));
    case TB_TONEMAP_UNCHARTED: return gamma_correct(uncharted2_filmic(color));
    case TB_TONEMAP_KHRONOS_PBR_NEUTRAL: return gamma_correct(commerce_tonemapping(color));
    case TB_TONEMAP_AGX: return agx_look(agx(color), false);
    case TB_TONEMAP_AGX_PUNCHY: return agx_look(agx(color), true);
    default: return gamma_correct(sat3(color));
    }
}

__device__ __forceinline__ f3 process_lit(float4 color, const TbPostProcessSettings& C, float averagedLuminance) {
    f3 o = mk3(color.x, color.y, color.z) / color.w;
    float exposure = C.UseAutoExposure ? pow_(0.5f, 2.2f) / averagedLuminance : C.ExposureMultiplier;
    return tonemap(C.TonemapType, o * exposure);
}
__device__ __forceinline__ f3 lerp3(f3 c0, f3 c1, f3 c2, float v) {
    return v < 0.5f ? lerp(c0, c1, v * 2.0f) : lerp(c1, c2, (v - 0.5f) * 2.0f);
}

__device__ __forceinline__ uint32_t luminance_to_bin(float luminance) {
    if (luminance < 0.00001f) return 0;
    float logLuminance = saturate((log2_(luminance) - -10.0f) * (1.0f / 16.0f));
    return (uint32_t)(logLuminance * 254.0f + 1.0f);
}

// scalar inputs (AOVDepth, R32_FLOAT) read as float4 the way a Texture2D<float4> load does: (r, 0, 0, 1)
__device__ __forceinline__ float4 load_px(const void* __restrict__ in, uint32_t scalar, size_t i) {
    if (scalar) return make_float4(__ldg((const float*)in + i), 0.0f, 0.0f, 1.0f);
    return __ldg((const float4*)in + i);
}

// The reference's 16x16 groups add bin Gid of their groupshared histogram from thread Gid, after the threads outside
// the image have returned (GenerateHistogramCS.hlsl:39, 54): in a partial group at the right / bottom edge the bins
// whose thread does not exist are never added. A pixel therefore counts iff thread `bin` of its group is inside the
// image; the blocks here are free to cover the pixels any way they like.
__global__ void __launch_bounds__(256) k_luminance_histogram(const void* __restrict__ in, uint32_t scalar, uint32_t width, uint32_t height,
                                                             uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const size_t n = (size_t)width * height;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float4 a = load_px(in, scalar, i);
        f3 c = mk3(a.x, a.y, a.z) / a.w;
        const uint32_t bin = luminance_to_bin(color_to_luma(c));
        const uint32_t x = (uint32_t)(i % width), y = (uint32_t)(i / width);
        if ((x & ~15u) + (bin & 15u) < width && (y & ~15u) + (bin >> 4) < height) atomicAdd(&sh[bin], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

__device__ __forceinline__ uint8_t to_unorm8(float c) { // typed UAV store: NaN -> 0, clamp, scale, round half up
    if (c != c) return 0;
    c = fminf(fmaxf(c, 0.0f), 1.0f);
    return (uint8_t)(uint32_t)(c * 255.0f + 0.5f);
}

__global__ void __launch_bounds__(256) k_postprocess(const void* __restrict__ in, uint32_t scalar, const float4* __restrict__ aux,
                                                     uint32_t width, uint32_t height, uint32_t outputType, TbPostProcessSettings C,
                                                     uint32_t* __restrict__ hist, float4* __restrict__ out, uchar4* __restrict__ out8) {
    const uint32_t n = width * height;
    __shared__ uint32_t shSum;
    __shared__ float shAvg;
    float averagedLuminance = 0.0f;
    if (C.UseAutoExposure) { // CalculateAveragedLuminanceCS: sum of bin * count in uint arithmetic, integer division
        if (threadIdx.x == 0) shSum = 0;
        __syncthreads();
        uint32_t v = hist[threadIdx.x] * threadIdx.x;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&shSum, v);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t denom = n - hist[0];
            uint32_t q = denom ? shSum / denom : 0xffffffffu;
            float averagedLogLuminance = ((float)q - 1.0f) / 254.0f;
            shAvg = exp2_(averagedLogLuminance * 16.0f + -10.0f);
            if (blockIdx.x == 0) hist[256] = __float_as_uint(shAvg); // AveragedLuminance
        }
        __syncthreads();
        averagedLuminance = shAvg;
    }
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float4 c = load_px(in, scalar, i);
        f3 o;
        switch (outputType) {
        default: o = process_lit(c, C, averagedLuminance); break;
        case TB_OUTPUT_ALBEDO:
            o = tonemap(C.TonemapType, mk3(c.x, c.y, c.z) * C.ExposureMultiplier);
            if (C.UseGammaCorrection) o = gamma_correct(o);
            break;
        case TB_OUTPUT_NORMALS: {
            uint32_t frameCount = (uint32_t)c.w;
            o = frameCount > 0 ? abs3(normalize(mk3(c.x, c.y, c.z) / (float)frameCount)) : mk3(0.0f, 0.0f, 0.0f);
            break;
        }
        case TB_OUTPUT_DEPTH:
        case TB_OUTPUT_LIVE_PIXELS: o = tonemap(C.TonemapType, mk3(c.x, c.y, c.z) * C.ExposureMultiplier); break;
        case TB_OUTPUT_MOTION_VECTORS:
            o = mk3(c.x / (float)width, c.y / (float)height, 0.0f);
            if (C.UseGammaCorrection) o = gamma_correct(o);
            break;
        case TB_OUTPUT_LUMINANCE: {
            f3 t = mk3(c.x, c.y, c.z) / c.w;
            t = tonemap(C.TonemapType, t * C.ExposureMultiplier);
            o = mk3(color_to_luma(t));
            if (C.UseGammaCorrection) o = gamma_correct(o);
            break;
        }
        case TB_OUTPUT_LUMINANCE_VARIANCE: o = C.VarianceMultiplier * mk3(c.x, 0.0f, 0.0f); break;
        case TB_OUTPUT_LIVE_WAVES: {
            o = process_lit(c, C, averagedLuminance);
            float4 f = aux ? __ldg(aux + i) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (f.x > 0.1f || f.y > 0.1f || f.z > 0.1f || f.w > 0.1f) o = mk3(f.x, f.y, f.z);
            break;
        }
        case TB_OUTPUT_HEATMAP: {
            uint32_t total = (uint32_t)c.x + (uint32_t)c.y;
            o = lerp3(mk3(0.0f, 1.0f, 0.0f), mk3(1.0f, 1.0f, 0.0f), mk3(1.0f, 0.0f, 0.0f), (float)total / 100.0f);
            o = tonemap(C.TonemapType, o * C.ExposureMultiplier);
            break;
        }
        }
        out[i] = make_float4(o.x, o.y, o.z, 1.0f);
        if (out8) out8[i] = make_uchar4(to_unorm8(o.x), to_unorm8(o.y), to_unorm8(o.z), 255);
    }
}

// ------------------------------------------------------------------ realtime temporal accumulation
// TemporalAccumulationCS.hlsl:100-235 (NEIGHBORHOOD_CLAMPING 0, WORLD_POSITION_HISTORY_REJECTION 1). One thread per
// pixel, 32x8 tiles so that a warp reads one row segment (coalesced float4 loads; the 3x3 world-position neighbourhood
// and the 2x2 reprojected history taps come from L1/L2). HBM-bound: 5 float4 images read once + 2 written = 112 B per
// pixel of compulsory traffic.
struct Img4 {
    const float4* p; int w, h;
    __device__ __forceinline__ float4 load(int x, int y) const { // Texture2D::operator[]: out of bounds reads zero
        if (!p || x < 0 || y < 0 || x >= w || y >= h) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        return __ldg(p + (size_t)y * w + x);
    }
    __device__ __forceinline__ f3 load3(int x, int y) const { float4 v = load(x, y); return mk3(v.x, v.y, v.z); }
    __device__ f3 bilinear_clamp(float u, float v) const { // SampleLevel(BilinearSampler, uv, 0), clamp addressing
        if (!p) return mk3(0.0f);
        float fx = u * (float)w - 0.5f, fy = v * (float)h - 0.5f;
        float x0f = floorf(fx), y0f = floorf(fy);
        float tx = fx - x0f, ty = fy - y0f;
        int x0 = min(max((int)x0f, 0), w - 1), x1 = min(max((int)(x0f + 1.0f), 0), w - 1);
        int y0 = min(max((int)y0f, 0), h - 1), y1 = min(max((int)(y0f + 1.0f), 0), h - 1);
        f3 a = lerp(load3(x0, y0), load3(x1, y0), tx);
        f3 b = lerp(load3(x0, y1), load3(x1, y1), tx);
        return lerp(a, b, ty);
    }
};
__device__ __forceinline__ f3 F3(const TbFloat3& v) { return mk3(v.x, v.y, v.z); }

__global__ void __launch_bounds__(256) k_temporal_accumulate(TbTemporalAccumulationParams P, int W, int H, const float4* __restrict__ history,
                                                             const float4* __restrict__ current, const float4* __restrict__ worldPos,
                                                             const float4* __restrict__ prevWorldPos, const float4* __restrict__ normals,
                                                             const float4* __restrict__ momentHistory, float4* __restrict__ outColor,
                                                             float4* __restrict__ outMoment) {
    const int px = blockIdx.x * 32 + (threadIdx.x & 31), py = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (px >= W || py >= H) return;
    Img4 hist{history, W, H}, cur{current, W, H}, wp{worldPos, W, H}, pwp{prevWorldPos, W, H}, nrm{normals, W, H}, mom{momentHistory, W, H};
    const f3 prevPos = F3(P.PrevCamera.Position);
    f3 WorldPosition = wp.load3(px, py);
    f3 WorldNormal = nrm.load3(px, py);
    bool bHitValid = WorldNormal.x != 0.0f || WorldNormal.y != 0.0f || WorldNormal.z != 0.0f;
    float aspectRatio = (float)W / (float)H;
    float lensHeight = P.Camera.LensHeight;
    float lensWidth = lensHeight * aspectRatio;
    f3 PrevFrameCameraDir = normalize(F3(P.PrevCamera.LookAt) - prevPos);
    f3 PrevFrameFocalPoint = prevPos - P.Camera.FocalDistance * PrevFrameCameraDir;
    f3 PrevFrameRayDirection = normalize(WorldPosition - PrevFrameFocalPoint);
    f3 RawOutputColor = cur.load3(px, py);
    f3 NMin = WorldPosition, NMax = WorldPosition;
#pragma unroll
    for (int x = -1; x <= 1; x++)
#pragma unroll
        for (int y = -1; y <= 1; y++) {
            int cx = px + x, cy = py + y;
            bool valid = cx > 0 && cy > 0 && cx < W && cy < H; // all(coord > 0), :139
            if (valid && !(x == 0 && y == 0)) {
                f3 w = wp.load3(cx, cy);
                NMin = min3(NMin, w);
                NMax = max3(NMax, w);
            }
        }
    f3 PrevFrameColor = mk3(0.0f), PrevMomentData = mk3(0.0f);
    float denom = dot(PrevFrameCameraDir, PrevFrameRayDirection); // PlaneIntersection, :73-82
    float t = fabsf(denom) > 0.0f ? dot(prevPos - PrevFrameFocalPoint, PrevFrameCameraDir) / denom : -1.0f;
    bool bValidHistory = false;
    if (!P.IgnoreHistory && t >= 0.0f && bHitValid) {
        f3 LensPosition = PrevFrameFocalPoint + PrevFrameRayDirection * t;
        f3 OffsetFromCenter = LensPosition - prevPos;
        float u = dot(OffsetFromCenter, F3(P.PrevCamera.Right)) / (lensWidth / 2.0f);
        float v = dot(OffsetFromCenter, F3(P.PrevCamera.Up)) / (lensHeight / 2.0f);
        u = (u + 1.0f) / 2.0f; v = (v + 1.0f) / 2.0f;
        v = 1.0f - v;
        if (u >= 0.0f && u <= 1.0f && v >= 0.0f && v <= 1.0f) {
            float distanceToNeighbor = length(NMax - NMin);
            float fx = u * (float)W - 0.5f, fy = v * (float)H - 0.5f;
            float SummedWeight = 0.0f;
#pragma unroll
            for (int x = 0; x < 2; x++)
#pragma unroll
                for (int y = 0; y < 2; y++) {
                    int ix = (int)fx + x, iy = (int)fy + y;
                    f3 PrevWP = pwp.load3(ix, iy);
                    if (length(PrevWP - WorldPosition) < distanceToNeighbor) {
                        float xWeight = x == 0 ? 1.0f - frac(fx) : frac(fx);
                        float yWeight = y == 0 ? 1.0f - frac(fy) : frac(fy);
                        float weight = xWeight * yWeight;
                        PrevFrameColor += hist.load3(ix, iy) * weight;
                        SummedWeight += weight;
                        if (P.OutputMomentInformation) PrevMomentData += mom.load3(ix, iy) * weight;
                    }
                }
            bValidHistory = SummedWeight > 0.0f;
            if (bValidHistory) { PrevFrameColor /= SummedWeight; PrevMomentData /= SummedWeight; }
            PrevMomentData = mom.bilinear_clamp(u, v); // :199, overrides the weighted sum
        }
    }
    float outputAlpha = 1.0f;
    if (P.OutputMomentInformation) {
        float luminance = color_to_luma(RawOutputColor);
        float luminanceSquared = luminance * luminance;
        float sampleCount = PrevMomentData.z + 1.0f;
        float lerpFactor = 1.0f / fminf(sampleCount, 32.0f);
        float m1 = lerp(PrevMomentData.x, luminance, lerpFactor), m2 = lerp(PrevMomentData.y, luminanceSquared, lerpFactor);
        if (outMoment) outMoment[(size_t)py * W + px] = make_float4(m1, m2, sampleCount, 0.0f);
        outputAlpha = fmaxf(m2 - m1 * m1, 0.0f);
    }
    f3 OutputColor = lerp(RawOutputColor, PrevFrameColor, bValidHistory ? P.HistoryWeight : 0.0f);
    outColor[(size_t)py * W + px] = make_float4(OutputColor.x, OutputColor.y, OutputColor.z, outputAlpha);
}

} // namespace

cudaError_t temporal_accumulate(const TbTemporalAccumulationParams& p, uint32_t width, uint32_t height, const float4* history,
                                const float4* current, const float4* worldPos, const float4* prevWorldPos, const float4* normals,
                                const float4* momentHistory, float4* outColor, float4* outMoment, cudaStream_t stream, LaunchCounter& lc) {
    dim3 grid((width + 31) / 32, (height + 7) / 8);
    k_temporal_accumulate<<<grid, 256, 0, stream>>>(p, (int)width, (int)height, history, current, worldPos, prevWorldPos, normals,
                                                    momentHistory, outColor, outMoment); lc.count++;
    return cudaGetLastError();
}

cudaError_t postprocess(const void* in, bool scalarInput, const float4* aux, uint32_t width, uint32_t height, uint32_t outputType,
                        const TbPostProcessSettings& s, uint32_t* hist257, float4* out, uchar4* out8, int numSMs,
                        cudaStream_t stream, LaunchCounter& lc) {
    const uint32_t n = width * height;
    uint32_t blocks = (n + 255) / 256;
    const uint32_t cap = (uint32_t)numSMs * 8; // 8 resident blocks of 256 threads per SM
    if (blocks > cap) blocks = cap;
    cudaMemsetAsync(hist257, 0, 257 * sizeof(uint32_t), stream);
    if (s.UseAutoExposure) { // the histogram passes run only with auto exposure (TracerBoy.cpp:2948)
        k_luminance_histogram<<<blocks, 256, 0, stream>>>(in, scalarInput ? 1u : 0u, width, height, hist257); lc.count++;
    }
    k_postprocess<<<blocks, 256, 0, stream>>>(in, scalarInput ? 1u : 0u, aux, width, height, outputType, s, hist257, out, out8); lc.count++;
    return cudaGetLastError();
}

} // namespace tbd
