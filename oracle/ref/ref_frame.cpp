// ref_frame.cpp — CPU ORACLE (test infrastructure): the per-pixel wrapper around PathTrace compiled from the mount —
// RayGenCommon.h: Halton / Halton23 / ApplyLDSToNoise / GetBlueNoise (:48-122), the AOV writers (:524-654), hash13
// (:662-667), RayTraceCommon (:690-728: NaN samples dropped whole, world-position ping-pong by frame parity, "frame 0
// overwrites" accumulation, the jittered half-buffer and its rand() coin) and the per-pixel part of
// SoftwareRayTraceCS.hlsl's main() (ClearAOVs, the hash13 seed, RayTraceCommon) — pre-passed into
// oracle/_ref/frame_gen.inc by prepass.run_frame. PathTrace is the deterministic stand-in of synthetic_tracer.h, the same
// one the oracle's render_frame is driven with in the test. Restated here: the resources (RWTexture2D element access, the
// two blue-noise textures as UNORM8 loads, the stats buffer), the PerFrameConstants members the text reads, integer
// vector types, and rand() (kernel.glsl:39-40). Single-threaded: the shader's statics are plain statics here.
#define RC_FRAME 1
#include "hlsl_compat.h"
#include <cstring>
#include "../glue.h"
#include "synthetic_tracer.h"

namespace refcore {

struct bool2 { bool x, y; };
struct bool4 { bool x, y, z, w; };
inline bool all(bool2 b) { return b.x && b.y; }
inline bool all(bool4 b) { return b.x && b.y && b.z && b.w; }
inline bool4 operator!(bool4 b) { return bool4{!b.x, !b.y, !b.z, !b.w}; }
inline bool4 isnan(float4 v) { return bool4{v.x != v.x, v.y != v.y, v.z != v.z, v.w != v.w}; }
struct uint2 {
    uint x, y;
    uint2() : x(0), y(0) {}
    uint2(uint a, uint b) : x(a), y(b) {}
    uint2 xy() const { return *this; }
    operator float2() const { return float2((float)x, (float)y); } // float3(Resolution, 1), float2(uint2)
};
inline uint2 operator%(uint2 a, int m) { return uint2(a.x % (uint)m, a.y % (uint)m); }
inline float2 operator+(uint2 a, float s) { return float2((float)a.x + s, (float)a.y + s); }
inline bool2 operator==(uint2 a, uint2 b) { return bool2{a.x == b.x, a.y == b.y}; }
inline uint asuint(float f) { uint u; memcpy(&u, &f, 4); return u; }
inline float2 frac(float2 v) { return float2(frac(v.x), frac(v.y)); }

template <typename T> struct RWTexture2D {
    T* p = nullptr; uint w = 0;
    T& operator[](uint2 i) { return p[(size_t)i.y * w + i.x]; }
};
struct UnormTexture256 { // Texture2D of DXGI_FORMAT_R8G8B8A8_UNORM, 256 x 256
    const uint8_t* p = nullptr;
    float4 operator[](uint2 i) const {
        const uint8_t* t = p + 4 * ((size_t)i.y * 256 + i.x);
        return float4((float)t[0] / 255.0f, (float)t[1] / 255.0f, (float)t[2] / 255.0f, (float)t[3] / 255.0f);
    }
};
struct StatsShim { uint words[8]; void Store(uint byteOffset, uint v) { words[byteOffset / 4] = v; } };
struct PerFrame { float Time; uint GlobalFrameCount, UseBlueNoise, OutputMode, IsRealTime, SelectedPixelX, SelectedPixelY; float MaxZ; };

static PerFrame perFrameConstants;
static RWTexture2D<float4> OutputTexture, JitteredOutputTexture, AOVNormals, AOVWorldPosition0, AOVWorldPosition1, AOVCustomOutput, AOVEmissive;
static RWTexture2D<float> AOVDepth;
static UnormTexture256 BlueNoise0Texture, BlueNoise1Texture;
static StatsShim StatsBuffer;
static float seed;                                                                                                  // kernel.glsl:39
inline float rand() { float s = seed; seed = s + 1.0f; return frac(sin(s + perFrameConstants.Time) * 43758.5453123f); } // kernel.glsl:40
float4 PathTrace(float2 pixelCoord);
#define IS_COMPUTE_SHADER 1
#define USE_ADAPTIVE_RAY_DISPATCHING 0
#define OUTPUT_TYPE_HEATMAP 9

#include "../_ref/frame_gen.inc"

struct RefSink {
    float rand() { return refcore::rand(); }
    void blue_noise(float o[8]) {
        BlueNoiseData d = GetBlueNoise();
        o[0] = d.PrimaryJitter.x; o[1] = d.PrimaryJitter.y; o[2] = d.SecondaryRayDirection.x; o[3] = d.SecondaryRayDirection.y;
        o[4] = d.AreaLightJitter.x; o[5] = d.AreaLightJitter.y; o[6] = d.DOFJitter.x; o[7] = d.DOFJitter.y;
    }
    void albedo(const float* c, float k) { OutputPrimaryAlbedo(float3(c[0], c[1], c[2]), k); }
    void normal(const float* n) { OutputPrimaryNormal(float3(n[0], n[1], n[2])); }
    void world_position(const float* p, float d) { OutputPrimaryWorldPosition(float3(p[0], p[1], p[2]), d); }
    void distance(float d) { OutputDistanceToFirstHit(d); }
    void material(int id) { OutputMaterial(id); }
    void emissive(const float* e) { OutputPrimaryEmissive(float3(e[0], e[1], e[2])); }
};
float4 PathTrace(float2 pixelCoord) {
    RefSink s;
    float c[4];
    synthetic_path(s, pixelCoord.x, pixelCoord.y, perFrameConstants.GlobalFrameCount, c);
    return float4(c[0], c[1], c[2], c[3]);
}

} // namespace refcore

// `frames` dispatches of the wrapper starting at GlobalFrameCount = firstFrame on caller-owned buffers (float4 images,
// depth float image, stats = 4 words: [2] SelectedPixelDistance bits, [3] SelectedMaterialID).
extern "C" __attribute__((visibility("default")))
int ref_render_synthetic(const void* scene, const TbOutputSettings* S, uint32_t width, uint32_t height, uint32_t firstFrame, uint32_t frames,
                         int selX, int selY, float time, TbFloat4* accum, TbFloat4* jittered, TbFloat4* normals, TbFloat4* worldPos0,
                         TbFloat4* worldPos1, TbFloat4* albedo, TbFloat4* emissive, float* depth, uint32_t* stats4) {
    using namespace refcore;
    const oracle::Scene& sc = *(const oracle::Scene*)scene;
    static_assert(sizeof(float4) == sizeof(TbFloat4), "float4 layout");
    OutputTexture = {(float4*)accum, width}; JitteredOutputTexture = {(float4*)jittered, width}; AOVNormals = {(float4*)normals, width};
    AOVWorldPosition0 = {(float4*)worldPos0, width}; AOVWorldPosition1 = {(float4*)worldPos1, width};
    AOVCustomOutput = {(float4*)albedo, width}; AOVEmissive = {(float4*)emissive, width}; AOVDepth = {depth, width};
    BlueNoise0Texture.p = sc.blueNoise.data(); BlueNoise1Texture.p = sc.blueNoise.data() + 256 * 256 * 4;
    memset(&StatsBuffer, 0, sizeof(StatsBuffer));
    perFrameConstants.Time = time; perFrameConstants.UseBlueNoise = S->EnableBlueNoise; perFrameConstants.OutputMode = S->OutputType;
    perFrameConstants.IsRealTime = S->RenderMode == TB_RENDER_REALTIME; perFrameConstants.MaxZ = S->MaxZ;
    perFrameConstants.SelectedPixelX = (uint)selX; perFrameConstants.SelectedPixelY = (uint)selY;
    for (uint32_t f = 0; f < frames; f++) {
        perFrameConstants.GlobalFrameCount = firstFrame + f;
        for (uint32_t y = 0; y < height; y++)
            for (uint32_t x = 0; x < width; x++) {
                Resolution = uint2(width, height);   // OutputTexture.GetDimensions, SoftwareRayTraceCS.hlsl:13
                DispatchIndex = uint2(x, y);         // :32 (the thread-group tiling only permutes which thread gets which pixel)
                pixel_main();
            }
    }
    memcpy(stats4, StatsBuffer.words, 16);
    return 0;
}
