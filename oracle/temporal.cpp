// temporal.cpp — CPU ORACLE (test infrastructure): realtime temporal accumulation.
//
// Hand restatement of TemporalAccumulationCS.hlsl:100-235 (NEIGHBORHOOD_CLAMPING 0, WORLD_POSITION_HISTORY_REJECTION 1)
// with the constants TemporalAccumulationPass.cpp:72-127 fills. Pinned against the reference's own shader text:
// oracle/_ref/libref_temporal.so compiles main() from the mount with only the resources shimmed (oracle/ref/
// ref_temporal.cpp) and tests/test_cpu_temporal.py requires this file to match it bit for bit. What stays a pinned
// CHOICE (texture-unit behaviour that has no source text): out-of-bounds texture loads return 0 (D3D); the bilinear
// fetch of the moment history (SampleLevel, clamp addressing) uses exact float weights.
#include <cmath>
#include <cstring>
#include "../tracerboy_b200/csrc/common/tb_vec.h"
#include "oracle.h"

using namespace tbm;

namespace {

inline f3 F3(const TbFloat3& v) { return mk3(v.x, v.y, v.z); }
inline f3 xyz(const TbFloat4& v) { return mk3(v.x, v.y, v.z); }

struct Img {
    const TbFloat4* p; int w, h;
    TbFloat4 load(int x, int y) const { // Texture2D::operator[]: out of bounds reads zero
        if (!p || x < 0 || y < 0 || x >= w || y >= h) return TbFloat4{0, 0, 0, 0};
        return p[(size_t)y * w + x];
    }
    f3 bilinear_clamp(float u, float v) const { // SampleLevel(BilinearSampler, uv, 0), clamp addressing
        if (!p) return mk3(0.0f);
        float fx = u * (float)w - 0.5f, fy = v * (float)h - 0.5f;
        float x0f = floorf(fx), y0f = floorf(fy);
        float tx = fx - x0f, ty = fy - y0f;
        auto cl = [](float f, int n) { int i = (int)f; return i < 0 ? 0 : (i > n - 1 ? n - 1 : i); };
        int x0 = cl(x0f, w), x1 = cl(x0f + 1.0f, w), y0 = cl(y0f, h), y1 = cl(y0f + 1.0f, h);
        f3 a = lerp(xyz(p[(size_t)y0 * w + x0]), xyz(p[(size_t)y0 * w + x1]), tx);
        f3 b = lerp(xyz(p[(size_t)y1 * w + x0]), xyz(p[(size_t)y1 * w + x1]), tx);
        return lerp(a, b, ty);
    }
};

float plane_intersection(f3 ro, f3 rd, f3 po, f3 pn) { // TemporalAccumulationCS.hlsl:73-82
    float denom = dot(pn, rd);
    if (fabsf(denom) > 0.0f) return dot(po - ro, pn) / denom;
    return -1.0f;
}

} // namespace

extern "C" __attribute__((visibility("default")))
int oracle_temporal_accumulate_image(const TbTemporalAccumulationParams* P, uint32_t width, uint32_t height,
                                     const TbFloat4* history, const TbFloat4* current, const TbFloat4* worldPos,
                                     const TbFloat4* prevWorldPos, const TbFloat4* normals, const TbFloat4* momentHistory,
                                     TbFloat4* outColor, TbFloat4* outMoment) {
    const int W = (int)width, H = (int)height;
    Img hist{history, W, H}, cur{current, W, H}, wp{worldPos, W, H}, pwp{prevWorldPos, W, H}, nrm{normals, W, H}, mom{momentHistory, W, H};
    const f3 prevPos = F3(P->PrevCamera.Position);
#pragma omp parallel for schedule(static)
    for (int py = 0; py < H; py++)
        for (int px = 0; px < W; px++) {
            f3 WorldPosition = xyz(wp.load(px, py));
            f3 WorldNormal = xyz(nrm.load(px, py));
            bool bHitValid = WorldNormal.x != 0.0f || WorldNormal.y != 0.0f || WorldNormal.z != 0.0f;
            float aspectRatio = (float)width / (float)height;
            float lensHeight = P->Camera.LensHeight;
            float lensWidth = lensHeight * aspectRatio;
            f3 PrevFrameCameraDir = normalize(F3(P->PrevCamera.LookAt) - prevPos);
            f3 PrevFrameFocalPoint = prevPos - P->Camera.FocalDistance * PrevFrameCameraDir;
            f3 PrevFrameRayDirection = normalize(WorldPosition - PrevFrameFocalPoint);
            f3 RawOutputColor = xyz(cur.load(px, py));
            f3 NMin = WorldPosition, NMax = WorldPosition;
            for (int x = -1; x <= 1; x++)
                for (int y = -1; y <= 1; y++) {
                    int cx = px + x, cy = py + y;
                    bool valid = cx > 0 && cy > 0 && cx < W && cy < H; // all(coord > 0): row and column 0 are excluded (:139)
                    if (valid && !(x == 0 && y == 0)) {
                        f3 w = xyz(wp.load(cx, cy));
                        NMin = min3(NMin, w);
                        NMax = max3(NMax, w);
                    }
                }
            f3 PrevFrameColor = mk3(0.0f), PrevMomentData = mk3(0.0f);
            float t = plane_intersection(PrevFrameFocalPoint, PrevFrameRayDirection, prevPos, PrevFrameCameraDir);
            bool bValidHistory = false;
            if (!P->IgnoreHistory && t >= 0.0f && bHitValid) {
                f3 LensPosition = PrevFrameFocalPoint + PrevFrameRayDirection * t;
                f3 OffsetFromCenter = LensPosition - prevPos;
                float u = dot(OffsetFromCenter, F3(P->PrevCamera.Right)) / (lensWidth / 2.0f);
                float v = dot(OffsetFromCenter, F3(P->PrevCamera.Up)) / (lensHeight / 2.0f);
                u = (u + 1.0f) / 2.0f; v = (v + 1.0f) / 2.0f;
                v = 1.0f - v;
                if (u >= 0.0f && u <= 1.0f && v >= 0.0f && v <= 1.0f) {
                    float distanceToNeighbor = length(NMax - NMin);
                    float fx = u * (float)width - 0.5f, fy = v * (float)height - 0.5f;
                    float SummedWeight = 0.0f;
                    for (int x = 0; x < 2; x++)
                        for (int y = 0; y < 2; y++) {
                            int ix = (int)fx + x, iy = (int)fy + y;
                            f3 PrevWP = xyz(pwp.load(ix, iy));
                            if (length(PrevWP - WorldPosition) < distanceToNeighbor) {
                                float xWeight = x == 0 ? 1.0f - frac(fx) : frac(fx);
                                float yWeight = y == 0 ? 1.0f - frac(fy) : frac(fy);
                                float weight = xWeight * yWeight;
                                PrevFrameColor += xyz(hist.load(ix, iy)) * weight;
                                SummedWeight += weight;
                                if (P->OutputMomentInformation) PrevMomentData += xyz(mom.load(ix, iy)) * weight;
                            }
                        }
                    bValidHistory = SummedWeight > 0.0f;
                    if (bValidHistory) { PrevFrameColor /= SummedWeight; PrevMomentData /= SummedWeight; }
                    PrevMomentData = mom.bilinear_clamp(u, v); // :199, overrides the weighted sum
                }
            }
            float outputAlpha = 1.0f;
            if (P->OutputMomentInformation) {
                float luminance = dot(RawOutputColor, mk3(0.212671f, 0.715160f, 0.072169f));
                float luminanceSquared = luminance * luminance;
                float sampleCount = PrevMomentData.z + 1.0f;
                float lerpFactor = 1.0f / fminf(sampleCount, 32.0f);
                float m1 = lerp(PrevMomentData.x, luminance, lerpFactor), m2 = lerp(PrevMomentData.y, luminanceSquared, lerpFactor);
                if (outMoment) outMoment[(size_t)py * W + px] = TbFloat4{m1, m2, sampleCount, 0.0f};
                outputAlpha = fmaxf(m2 - m1 * m1, 0.0f);
            }
            f3 OutputColor = lerp(RawOutputColor, PrevFrameColor, bValidHistory ? P->HistoryWeight : 0.0f);
            outColor[(size_t)py * W + px] = TbFloat4{OutputColor.x, OutputColor.y, OutputColor.z, outputAlpha};
        }
    return 0;
}
