// pathtrace.cu — wavefront path tracer for sm_100a.
//
// Role: SoftwareRayTraceCS.hlsl:9-51 -> RayTraceCommon (RayGenCommon.h:690-728) ->
// PathTrace (kernel.glsl:1805-1921) -> Trace (kernel.glsl:1278-1776), i.e. the
// reference's one-thread-per-pixel megakernel, re-designed as a wavefront:
//
//   k_raygen   one thread per pixel: seed (hash13), jitter, camera ray, path state, AOV clear
//   k_extend   persistent threads over the ray queue: closest hit (traverse.cuh)
//   k_shade    persistent threads over the same queue: one iteration of the bounce loop
//              (material fetch, emissive, NEE + inline shadow query, BSDF sampling,
//              throughput update, russian roulette for the next bounce); survivors are
//              appended to the next queue with one warp-aggregated atomic per warp;
//              terminated paths are accumulated in-kernel (OutputTexture +=)
//
// Per-pixel arithmetic follows the reference's order of operations with the intrinsics
// pinned in common/tb_math.h / tb_vec.h (compiled with -fmad=false), and consumes the
// reference's sequential per-pixel rand() stream in the reference's order (SURVEY
// Appendix A) so that results are bit-identical to the CPU oracle.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../common/tb_vec.h"
#include "device_types.h"
#include "launch.h"
#include "pathtrace.h"
#include "traverse.cuh"

using namespace tbm;

namespace tbd {

namespace {

#define EPSILON 0.000001f
#define PI_K 3.1415926535f
#define LARGE_NUMBER 1e20f
#define AIR_IOR 1.0f
#define MIN_ROUGHNESS 0.04f
#define MIN_ROUGHNESS_SQUARED (0.04f * 0.04f)
#define MIN_T 0.001f
#define FAR_T 999999.0f
#define AOV_FULL 1u     // this frame is the last one of the render call: it owns the overwritten-every-frame AOVs
#define AOV_WORLDPOS 2u // last or last-1 frame: owns its ping-pong world-position buffer

__device__ __forceinline__ f3 F3(const TbFloat3& v) { return mk3(v.x, v.y, v.z); }

struct Mat {
    f3 albedo; uint32_t albedoIndex, normalMapIndex, emissiveIndex, specularMapIndex;
    float IOR; f3 absorption; float roughness; f3 scattering; f3 emissive; int Flags; float SpecularCoef;
};

__device__ __forceinline__ Mat load_mat(const TbMaterial* __restrict__ mats, uint32_t id) {
    // 84 B = 21 words; read-only path
    const uint32_t* p = (const uint32_t*)(mats + id);
    uint32_t w[21];
#pragma unroll
    for (int i = 0; i < 21; i++) w[i] = __ldg(p + i);
    Mat m;
    m.albedo = mk3(__uint_as_float(w[0]), __uint_as_float(w[1]), __uint_as_float(w[2]));
    m.albedoIndex = w[3]; m.normalMapIndex = w[5]; m.emissiveIndex = w[6]; m.specularMapIndex = w[7];
    m.IOR = __uint_as_float(w[8]);
    m.absorption = mk3(__uint_as_float(w[9]), __uint_as_float(w[10]), __uint_as_float(w[11]));
    m.roughness = __uint_as_float(w[12]);
    m.scattering = mk3(__uint_as_float(w[13]), __uint_as_float(w[14]), __uint_as_float(w[15]));
    m.emissive = mk3(__uint_as_float(w[16]), __uint_as_float(w[17]), __uint_as_float(w[18]));
    m.Flags = (int)w[19];
    m.SpecularCoef = __uint_as_float(w[20]);
    return m;
}

struct Rng {
    float seed, time;
    __device__ __forceinline__ float next() { float s = seed; seed = seed + 1.0f; return frac(sin_(s + time) * 43758.5453123f); }
};

// ------------------------------------------------------------------ textures
// The sampler's sRGB -> linear conversion of an _SRGB format (D3D11.3 functional spec 7.2.1.2), pinned intrinsics.
__device__ __forceinline__ float srgb_to_linear(float c) { return c <= 0.04045f ? c / 12.92f : pow_((c + 0.055f) / 1.055f, 2.4f); }
__device__ __forceinline__ f4 fetch_texel(const DeviceScene::ImageRef& im, int x, int y) {
    if (im.format == 0) return *((const f4*)im.data + ((size_t)y * im.width + x));
    uchar4 c = *((const uchar4*)im.data + ((size_t)y * im.width + x));
    f4 v = mk4((float)c.x / 255.0f, (float)c.y / 255.0f, (float)c.z / 255.0f, (float)c.w / 255.0f);
    if (im.format == 2) { v.x = srgb_to_linear(v.x); v.y = srgb_to_linear(v.y); v.z = srgb_to_linear(v.z); } // R8G8B8A8_UNORM_SRGB: per texel, before filtering
    return v;
}
__device__ __forceinline__ int wrapi(int i, int n) { i %= n; return i < 0 ? i + n : i; }
__device__ __forceinline__ f4 lerp4(f4 a, f4 b, float s) { return mk4(lerp(a.x, b.x, s), lerp(a.y, b.y, s), lerp(a.z, b.z, s), lerp(a.w, b.w, s)); }
__device__ f4 sample_bilinear_wrap(const DeviceScene::ImageRef& im, f2 uv) {
    const float nanv = __uint_as_float(0x7fc00000u);
    if (!(fabsf(uv.x) <= 1.0e6f) || !(fabsf(uv.y) <= 1.0e6f)) return mk4(nanv, nanv, nanv, nanv);
    float fx = uv.x * (float)im.width - 0.5f, fy = uv.y * (float)im.height - 0.5f;
    float x0f = floorf(fx), y0f = floorf(fy);
    float tx = fx - x0f, ty = fy - y0f;
    int x0 = wrapi((int)x0f, (int)im.width), y0 = wrapi((int)y0f, (int)im.height);
    int x1 = wrapi(x0 + 1, (int)im.width), y1 = wrapi(y0 + 1, (int)im.height);
    f4 a = lerp4(fetch_texel(im, x0, y0), fetch_texel(im, x1, y0), tx);
    f4 b = lerp4(fetch_texel(im, x0, y1), fetch_texel(im, x1, y1), tx);
    return lerp4(a, b, ty);
}
__device__ f4 texture_nonrecursive(const DeviceScene& sc, const TbTextureData& t, f2 uv) {
    f4 data = mk4(0, 0, 0, 0);
    if (t.TextureType == TB_IMAGE_TEXTURE_TYPE) data = sample_bilinear_wrap(sc.images[t.DescriptorHeapIndex], uv);
    else if (t.TextureType == TB_CHECKER_TEXTURE_TYPE) {
        f2 scaled = uv * mk2(t.UScale, t.VScale);
        data = mk4(F3(t.CheckerColor1), 1.0f);
        if ((((int)scaled.x + (int)scaled.y) % 2) == 0) data = mk4(F3(t.CheckerColor2), 1.0f);
    }
    if (t.TextureFlags & TB_NEEDS_GAMMA_CORRECTION_TEXTURE_FLAG) { data.x = pow_(data.x, 2.2f); data.y = pow_(data.y, 2.2f); data.z = pow_(data.z, 2.2f); }
    return data;
}
__device__ f4 get_texture_data(const DeviceScene& sc, uint32_t index, f2 uv) {
    if (index == TB_INVALID_TEXTURE) return mk4(0, 0, 0, 0);
    if (sc.flipTextureUVs) uv = mk2(0.0f, 1.0f) + uv * mk2(1.0f, -1.0f);
    TbTextureData t = sc.textures[index];
    if (t.TextureType == TB_SCALE_TEXTURE_TYPE) {
        f4 c1 = texture_nonrecursive(sc, sc.textures[t.TextureIndex1], uv);
        f4 c2 = texture_nonrecursive(sc, sc.textures[t.TextureIndex2], uv);
        return c1 * mk4(F3(t.ScaleColor1), 1.0f) + c2 * mk4(F3(t.ScaleColor2), 1.0f);
    }
    return texture_nonrecursive(sc, t, uv);
}
__device__ f3 sample_environment_map(const DeviceScene& sc, f3 v) {
    if (sc.envImage < 0) return mk3(0.0f);
    v = mk3(dot(v, mk3(sc.envTransform[0][0], sc.envTransform[0][1], sc.envTransform[0][2])),
            dot(v, mk3(sc.envTransform[1][0], sc.envTransform[1][1], sc.envTransform[1][2])),
            dot(v, mk3(sc.envTransform[2][0], sc.envTransform[2][1], sc.envTransform[2][2])));
    f3 viewDir = normalize(v);
    float p = atan2_(viewDir.y, viewDir.x);
    p = p > 0.0f ? p : p + 2.0f * 3.14f;
    f2 uv;
    uv.x = p / (2.0f * 3.14f);
    uv.y = acos_(viewDir.z) / 3.14f;
    f4 t = sample_bilinear_wrap(sc.images[sc.envImage], uv);
    return mk3(t.x, t.y, t.z) * mk3(sc.envColorScale[0], sc.envColorScale[1], sc.envColorScale[2]);
}

// -------------------------------------------------------------------- noise
__device__ __forceinline__ float hash13(f3 p3) {
    p3 = frac3(p3 * 0.1031f);
    float d = dot(p3, mk3(p3.y, p3.z, p3.x) + 33.33f);
    p3 = p3 + d;
    return frac((p3.x + p3.y) * p3.z);
}
struct BlueNoise { f2 primary, dof; };
__device__ __forceinline__ BlueNoise get_blue_noise(const DeviceScene& sc, const FrameConstants& fc, Rng& rng, uint32_t px, uint32_t py) {
    BlueNoise d;
    if (!fc.settings.EnableBlueNoise) {
        float a, b;
        a = rng.next(); b = rng.next(); d.primary = mk2(a, b);
        a = rng.next(); b = rng.next();
        a = rng.next(); b = rng.next();
        a = rng.next(); b = rng.next(); d.dof = mk2(a, b);
    } else {
        f2 h = mk2(fc.halton2, fc.halton3);
        const uchar4* t = (const uchar4*)sc.blueNoise + ((size_t)(py % 256) * 256 + (px % 256));
        uchar4 t0 = t[0], t1 = t[256 * 256];
        d.primary = frac2(mk2((float)t0.x / 255.0f, (float)t0.y / 255.0f) + h);
        d.dof = frac2(mk2((float)t1.z / 255.0f, (float)t1.w / 255.0f) + h);
    }
    return d;
}

// ------------------------------------------------------------------ materials
__device__ __forceinline__ bool AllowsSpecular(const Mat& m) { return (m.Flags & TB_NO_SPECULAR_MATERIAL_FLAG) == 0; }
__device__ __forceinline__ bool IsMetallic(const Mat& m) { return (m.Flags & TB_METALLIC_MATERIAL_FLAG) != 0; }
__device__ __forceinline__ bool IsSSS(const Mat& m) { return (m.Flags & TB_SUBSURFACE_SCATTER_MATERIAL_FLAG) != 0; }
__device__ __forceinline__ bool IsHair(const Mat& m) { return (m.Flags & TB_HAIR_MATERIAL_FLAG) != 0; }
__device__ __forceinline__ bool IsLight(const Mat& m) { return (m.Flags & TB_LIGHT_MATERIAL_FLAG) != 0; }

__device__ Mat get_material(const DeviceScene& sc, Rng& rng, int id, f2 uv, bool backside) {
    Mat mat = load_mat(sc.materials, (uint32_t)id);
    bool ignoreEmissive = backside;
    if (ignoreEmissive) mat.emissive = mk3(0.0f);
    if ((mat.Flags & TB_MIX_MATERIAL_FLAG) != 0) {
        if (rng.next() < mat.albedo.z) mat = load_mat(sc.materials, (uint32_t)mat.albedo.x);
        else mat = load_mat(sc.materials, (uint32_t)mat.albedo.y);
    } else {
        if (mat.albedoIndex != TB_INVALID_TEXTURE) { f4 t = get_texture_data(sc, mat.albedoIndex, uv); mat.albedo = mk3(t.x, t.y, t.z); }
        if (mat.emissiveIndex != TB_INVALID_TEXTURE && !ignoreEmissive) { f4 t = get_texture_data(sc, mat.emissiveIndex, uv); mat.emissive = mk3(t.x, t.y, t.z); }
        if (mat.specularMapIndex != TB_INVALID_TEXTURE) {
            f4 t = get_texture_data(sc, mat.specularMapIndex, uv);
            mat.roughness = t.y;
            if (t.z > 0.5f) mat.Flags |= TB_METALLIC_MATERIAL_FLAG;
        }
    }
    if (IsSSS(mat) && (mat.albedo.x != 0.0f || mat.albedo.y != 0.0f || mat.albedo.z != 0.0f)) {
        f3 color = mat.albedo;
        f3 mfp = mk3(1.0f) / mat.scattering;
        f3 alpha = mk3(1.0f) - exp3(((-5.09406f * color) + ((2.61188f * color) * color)) - (((4.31805f * color) * color) * color));
        f3 s = (mk3(1.9f) - color) + ((3.5f * (color - 0.8f)) * (color - 0.8f));
        f3 transmission = mk3(1.0f) / (s * mfp);
        mat.scattering = transmission * alpha;
        mat.absorption = transmission - mat.scattering;
        mat.albedo = mk3(0.0f);
    }
    return mat;
}
__device__ f3 get_detail_normal(const DeviceScene& sc, const FrameConstants& fc, const Mat& mat, f3 normal, f3 tangent, f2 uv) {
    if (mat.normalMapIndex != TB_INVALID_TEXTURE && fc.settings.EnableNormalMaps) {
        f3 bitangent = cross(tangent, normal);
        f4 nm = get_texture_data(sc, mat.normalMapIndex, uv);
        f3 tbn = mk3((0.5f - nm.x) * 2.0f, (0.5f - nm.y) * 2.0f, 0.0f);
        tbn.z = sqrtf(1.0f - (tbn.x * tbn.x + tbn.y * tbn.y));
        return normalize((tangent * tbn.x + bitangent * tbn.y) + normal * fmaxf(tbn.z, 0.02f));
    }
    return normal;
}

// --------------------------------------------------------------------- lights
__device__ __forceinline__ f3 random_barycentric(Rng& rng) {
    float u = rng.next();
    float v = rng.next();
    if (u + v > 1.0f) { u = 1.0f - u; v = 1.0f - v; }
    return mk3(u, v, 1.0f - u - v);
}
__device__ __forceinline__ float light_target_pdf(const TbLight& l, f3 bary, f3 pos) {
    f3 lp = (F3(l.P0) * bary.x + F3(l.P1) * bary.y) + F3(l.P2) * bary.z;
    float d = length(lp - pos);
    return (l.SurfaceArea * dot(F3(l.LightColor), mk3(0.212671f, 0.715160f, 0.072169f))) / d * d;
}
__device__ void get_one_light_sample(const DeviceScene& sc, const FrameConstants& fc, Rng& rng, f3 pos, f3& LightDirection,
                                     f3& LightColor, float& PDFValue, f3& LightNormal, float& LightAttenuation) {
    LightDirection = LightColor = LightNormal = mk3(0.0f);
    LightAttenuation = 0.0f;
    PDFValue = 0.0f;
    const uint32_t lightCount = sc.numLights;
    if (lightCount > 0 && fc.settings.EnableNextEventEstimation) {
        if (fc.settings.EnableSamplingImportanceResampling) {
            uint32_t selIndex = 0; f3 selBary = mk3(0.0f); float weightSum = 0.0f;
            for (uint32_t i = 0; i < 16; i++) {
                uint32_t li = (uint32_t)(rng.next() * (float)lightCount);
                TbLight light = sc.lights[li];
                f3 bary = random_barycentric(rng);
                float target = light_target_pdf(light, bary, pos);
                float proposal = 1.0f / (float)lightCount;
                float w = target / (proposal * 16.0f);
                weightSum += w;
                if (rng.next() < w / weightSum) { selIndex = li; selBary = bary; }
            }
            TbLight light = sc.lights[selIndex];
            float sirPDF = light_target_pdf(light, selBary, pos) / weightSum;
            PDFValue = sirPDF / light.SurfaceArea;
            f3 lp = (F3(light.P0) * selBary.x + F3(light.P1) * selBary.y) + F3(light.P2) * selBary.z;
            LightDirection = lp - pos;
            LightNormal = (F3(light.N0) * selBary.x + F3(light.N1) * selBary.y) + F3(light.N2) * selBary.z;
            LightColor = F3(light.LightColor);
        } else {
            uint32_t li = (uint32_t)(rng.next() * (float)lightCount);
            TbLight light = sc.lights[li];
            f3 bary = random_barycentric(rng);
            if (light.LightType == TB_LIGHT_TYPE_AREA) {
                f3 lp = (F3(light.P0) * bary.x + F3(light.P1) * bary.y) + F3(light.P2) * bary.z;
                LightDirection = lp - pos;
                LightNormal = (F3(light.N0) * bary.x + F3(light.N1) * bary.y) + F3(light.N2) * bary.z;
                float d = length(LightDirection);
                LightAttenuation = 1.0f / (d * d);
                LightDirection = LightDirection / d;
            } else if (light.LightType == TB_LIGHT_TYPE_DIRECTIONAL) {
                LightDirection = -F3(light.Direction);
                if (fc.settings.DebugValue > 0.0f) {
                    LightDirection.x = sin_(fc.settings.DebugValue);
                    LightDirection.y = sin_(fc.settings.DebugValue2);
                    LightDirection = normalize(LightDirection);
                }
                LightNormal = -LightDirection;
                LightAttenuation = 1.0f;
            }
            LightColor = F3(light.LightColor);
            PDFValue = 1.0f / (float)lightCount;
            if (light.LightType == TB_LIGHT_TYPE_AREA) PDFValue /= light.SurfaceArea;
        }
    }
}

// ------------------------------------------------------------------- sampling
__device__ __forceinline__ f3 reorient_around_normal(f3 v, f3 normal) {
    f3 tangent;
    if (fabsf(normal.x) > fabsf(normal.y)) tangent = mk3(-normal.z, 0.0f, normal.x) / sqrtf(normal.x * normal.x + normal.z * normal.z);
    else tangent = mk3(0.0f, normal.z, -normal.y) / sqrtf(normal.y * normal.y + normal.z * normal.z);
    f3 bitangent = cross(normal, tangent);
    return normalize((v.x * tangent + v.y * normal) + v.z * bitangent);
}
__device__ __forceinline__ f3 importance_sampled_direction(f3 normal, float roughness, float rand0, float rand1, float& pdf) {
    float lobe = pow_(1.0f - roughness, 5.0f) * 1000.0f;
    float theta = 2.0f * PI_K * rand1;
    float phi = acos_(sqrtf(pow_(rand0, 1.0f / (lobe + 1.0f))));
    f3 d = mk3(sin_(phi) * cos_(theta), cos_(phi), sin_(phi) * sin_(theta));
    pdf = (lobe + 1.0f) * pow_(cos_(phi), lobe) / (2.0f * PI_K);
    return reorient_around_normal(d, normal);
}
__device__ __forceinline__ float ggx_pdf(f3 normal, f3 outgoing, f3 halfVector, float roughness) {
    roughness = fmaxf(MIN_ROUGHNESS, roughness);
    float a = roughness * roughness;
    float a2 = a * a;
    float cosTheta = fabsf(dot(normal, halfVector));
    float e = ((a2 - 1.0f) * cosTheta) * cosTheta + 1.0f;
    if (e <= 0.0f) return LARGE_NUMBER;
    float d = a2 / ((PI_K * e) * e);
    return d * fabsf(dot(halfVector, normal)) / (4.0f * fabsf(dot(outgoing, halfVector)));
}
__device__ __forceinline__ float ggx_ndf(f3 normal, f3 halfVector, float roughnessSquared) {
    roughnessSquared = fmaxf(roughnessSquared, MIN_ROUGHNESS_SQUARED);
    float a2 = roughnessSquared * roughnessSquared;
    float nDotH = dot(normal, halfVector);
    float denom = PI_K * pow_((nDotH * nDotH) * (a2 - 1.0f) + 1.0f, 2.0f);
    return a2 / denom;
}
__device__ __forceinline__ float diffuse_brdf(f3 l, f3 n) { return fmaxf(dot(l, n), 0.0f) / PI_K; }
__device__ __forceinline__ f3 half_vector_safe(f3 a, f3 b, f3 normal) {
    if (dot(a, b) > (-1.0f + EPSILON)) return normalize(a + b);
    return normal;
}
enum RefractResult { REFRACTED, REFLECTED, GIVE_UP };
__device__ RefractResult refract_or_reflect(Rng& rng, f3& dir, f3 normal, float nr, float RdotN, bool perfectSpec, float roughness, bool& prevPerfectlySpecular) {
    float discriminant = 1.0f - (nr * nr) * (1.0f - RdotN * RdotN);
    if (discriminant > EPSILON) {
        f3 refr = normalize(nr * (dir - normal * RdotN) - normal * sqrtf(discriminant));
        if (perfectSpec) { dir = refr; prevPerfectlySpecular = true; }
        else {
            float pdf;
            float r0 = rng.next();
            float r1 = rng.next();
            dir = importance_sampled_direction(refr, roughness, r0, r1, pdf);
            if (pdf < EPSILON) {
                r0 = rng.next();
                r1 = rng.next();
                dir = importance_sampled_direction(refr, roughness, r0, r1, pdf);
                if (pdf < EPSILON) return GIVE_UP;
            }
        }
        return REFRACTED;
    }
    dir = reflect(dir, normal);
    return REFLECTED;
}

// ------------------------------------------------------------- hit attributes
struct Surface { f3 normal, tangent; f2 uv; int material; };
__device__ __forceinline__ Surface surface_from_hit(const DeviceScene& sc, float b1, float b2, uint32_t geom, uint32_t prim) {
    // GetGeometryInfo/GetHitInfo, SharedHitGroup.h:48-151
    Surface s;
    TbGeometryRecord G = sc.geoms[geom];
    f3 bary = mk3(1.0f - b1 - b2, b1, b2);
    const uint32_t* idx = sc.indices + G.IndexFirst + 3 * (size_t)prim;
    uint32_t i0 = __ldg(idx), i1 = __ldg(idx + 1), i2 = __ldg(idx + 2);
    const float4* va = (const float4*)(sc.vertices + G.VertexFirst + i0);
    const float4* vb = (const float4*)(sc.vertices + G.VertexFirst + i1);
    const float4* vc = (const float4*)(sc.vertices + G.VertexFirst + i2);
    float4 a0 = __ldg(va), a1 = __ldg(va + 1), b0 = __ldg(vb), b1v = __ldg(vb + 1), c0 = __ldg(vc), c1 = __ldg(vc + 1);
    // Vertex = normal3, uv2, tangent3
    f2 uv0 = mk2(a0.w, a1.x), uv1 = mk2(b0.w, b1v.x), uv2 = mk2(c0.w, c1.x);
    s.uv = (bary.x * uv0 + bary.y * uv1) + bary.z * uv2;
    s.normal = normalize((bary.x * mk3(a0.x, a0.y, a0.z) + bary.y * mk3(b0.x, b0.y, b0.z)) + bary.z * mk3(c0.x, c0.y, c0.z));
    s.tangent = normalize((bary.x * mk3(a1.y, a1.z, a1.w) + bary.y * mk3(b1v.y, b1v.z, b1v.w)) + bary.z * mk3(c1.y, c1.z, c1.w));
    s.material = (int)G.MaterialIndex;
    return s;
}

struct RayCount { uint32_t rays, tris, boxes; };

// Intersect() for the rays that stay inside the shading stage (shadow feelers, SSS walk)
__device__ __forceinline__ bool intersect_inline(const DeviceBvh& bvh, const DeviceScene& sc, f3 org, f3 dir, RayCount& rc, Surface& s, float& t) {
    HitRec h;
    trace_ray(bvh, org, dir, MIN_T, FAR_T, h);
    rc.rays++; rc.tris += h.tris; rc.boxes += h.boxes;
    if (h.t >= 0.0f) { s = surface_from_hit(sc, h.b1, h.b2, h.geom, h.prim); t = h.t; return true; }
    s.normal = mk3(0.0f); s.tangent = mk3(0.0f); s.uv = mk2(0, 0); s.material = -1; t = -1.0f;
    return false;
}

__device__ __forceinline__ f3 lens_position(const TbCamera& cam, f2 uv, float aspect) {
    f3 p = F3(cam.Position);
    float lensWidth = cam.LensHeight * aspect;
    p += ((F3(cam.Right) * (uv.x * 2.0f - 1.0f)) * lensWidth) / 2.0f;
    p += ((F3(cam.Up) * (uv.y * 2.0f - 1.0f)) * cam.LensHeight) / 2.0f;
    return p;
}
__device__ __forceinline__ float gaussian(float x, float mu, float sigma) {
    float d = x - mu;
    return 1.0f / sqrtf((2.0f * PI_K) * sigma * sigma) * exp_(-(d * d) / ((2.0f * sigma) * sigma));
}

// state word packing in rayD.w: bits 0..7 bounce index, bit 8 prevPerfectlySpecular
__device__ __forceinline__ uint32_t pack_state(int bounce, bool prevSpec) { return (uint32_t)bounce | (prevSpec ? 0x100u : 0u); }
// walkers (k_shade -> k_walk) also carry bit 9: the bounce's perfect-specular flag

// ---------------------------------------------------------------------- raygen
__global__ void __launch_bounds__(256) k_raygen(DeviceScene sc, const FrameConstants* __restrict__ fcp, PathState st) {
    const FrameConstants& fc = *fcp;
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t pi = slot;
    uint32_t n = fc.width * fc.height;
    // Thread -> pixel mapping (the role of ThreadGroupTilingX, ComputeShaderUtil.h:12-67; results do not
    // depend on it): a warp owns an 8x4 pixel tile, so the 32 primary rays that enter k_extend together
    // walk nearly the same nodes and their loads coalesce into few L1 wavefronts.
    if ((fc.width & 7u) == 0 && (fc.height & 3u) == 0) {
        const uint32_t tile = slot >> 5, l = slot & 31u, tilesX = fc.width >> 3;
        pi = ((tile / tilesX) * 4u + (l >> 3)) * fc.width + (tile % tilesX) * 8u + (l & 7u);
        if (slot >= n) pi = n;
    }
    // (all queue / suspension counters were zeroed by the memset node in front of this kernel)
    // row-band sharding (partitioning 1): bands of 8 rows, band b belongs to the shard b % stride
    const bool owned = pi < n && ((pi / fc.width) / 8u) % fc.rowStride == fc.rowOffset;
    if (fc.rowStride == 1) { // every pixel: the first queue is the identity
        if (slot == 0) st.queueCount[0] = n;
        if (slot < n) st.queue[0][slot] = pi;
    } else {
        uint32_t m = __ballot_sync(0xffffffffu, owned);
        if (m) {
            uint32_t lane = threadIdx.x & 31, base = 0;
            if (lane == 0) base = atomicAdd(&st.queueCount[0], (uint32_t)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (owned) st.queue[0][base + __popc(m & ((1u << lane) - 1u))] = pi;
        }
    }
    if (!owned) return;
    uint32_t px = pi % fc.width, py = pi / fc.width;
    Rng rng;
    rng.time = fc.time;
    rng.seed = hash13(mk3((float)px, (float)py, (float)fc.frame));
    f2 res = mk2((float)fc.width, (float)fc.height);
    f2 dispatchUV = mk2((float)px + 0.5f, (float)py + 0.5f) / res;
    f2 uv0 = mk2(0.0f, 1.0f) + dispatchUV * mk2(1.0f, -1.0f);
    f2 pixelCoord = uv0 * res;
    // PathTrace
    f2 pixelUVSize = mk2(1.0f / res.x, 1.0f / res.y);
    f2 uv = pixelCoord * pixelUVSize;
    BlueNoise bn = get_blue_noise(sc, fc, rng, px, py);
    f2 off = bn.primary - mk2(0.5f, 0.5f);
    float pixelRadius = fc.settings.FilterWidth / 2.0f;
    float filterWeight = 1.0f;
    if (fc.settings.FilterType == TB_FILTER_TRIANGLE) filterWeight = fmaxf(0.5f - fabsf(off.x), 0.5f - fabsf(off.y));
    else if (fc.settings.FilterType == TB_FILTER_GAUSSIAN) {
        float sigma = 0.8f;
        float eX = gaussian(1.0f, 0.0f, sigma), eY = gaussian(1.0f, 0.0f, sigma);
        filterWeight = fmaxf(0.0f, gaussian(off.x * 2.0f, 0.0f, sigma) - eX) * fmaxf(0.0f, gaussian(off.y * 2.0f, 0.0f, sigma) - eY);
    }
    uv = uv + (off * pixelUVSize) * (pixelRadius * 2.0f);
    float aspect = res.x / res.y;
    f3 camPos = F3(fc.camera.Position);
    f3 focalPoint = camPos - fc.camera.FocalDistance * normalize(F3(fc.camera.LookAt) - camPos);
    f3 lensPoint = lens_position(fc.camera, uv, aspect);
    f3 neighborLensPoint = lens_position(fc.camera, uv + pixelUVSize, aspect);
    f3 org = focalPoint, dir = normalize(lensPoint - focalPoint);
    f3 ndir = normalize(neighborLensPoint - focalPoint);
    if (fc.settings.DOFFocalDistance > 0.0f) {
        f3 FocusPoint = org + dir * fc.settings.DOFFocalDistance;
        float Radius = sqrtf(bn.dof.x) * fc.settings.ApertureWidth;
        float Theta = (bn.dof.y * 2.0f) * PI_K;
        f2 fj = mk2(cos_(Theta) * Radius, sin_(Theta) * Radius);
        org = org + (fj.x * F3(fc.camera.Right) + fj.y * F3(fc.camera.Up));
        dir = normalize(FocusPoint - org);
    }
    // Trace(): GetBlueNoise() again (values unused, burns 8 rand() when blue noise is off), kernel.glsl:1283
    if (!fc.settings.EnableBlueNoise) { for (int k = 0; k < 8; k++) (void)rng.next(); }
    st.rayO[pi] = make_float4(org.x, org.y, org.z, rng.seed);
    st.rayD[pi] = make_float4(dir.x, dir.y, dir.z, __uint_as_float(pack_state(0, false)));
    st.thr[pi] = make_float4(1.0f, 1.0f, 1.0f, filterWeight);
    st.col[pi] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    st.neighbor[pi] = make_float4(focalPoint.x, focalPoint.y, focalPoint.z, 0.0f);
    st.neighborDir[pi] = make_float4(ndir.x, ndir.y, ndir.z, 0.0f);
    // ClearAOVs (RayGenCommon.h:650-654) + zeroed world-position accumulators (:693-694)
    if (fc.aovMask & AOV_FULL) {
        st.aovAlbedo[pi] = make_float4(0, 0, 0, 1.0f);
        st.aovNormal[pi] = make_float4(0, 0, 0, 1.0f);
        st.primaryHit[pi] = make_uint2(0xffffffffu, 0xffffffffu);
        st.counters[pi] = make_uint2(0, 0);
    }
    if (fc.aovMask & AOV_WORLDPOS) st.aovWorldPos[fc.worldPosSlot][pi] = make_float4(0, 0, 0, 0);
    // per-frame staging of the two AOVs that persist when a frame does not write them
    st.stEmissive[pi] = make_float4(0, 0, 0, 0);
    st.stDepth[pi] = -1.0f;
    if (fc.settings.MaxBounces <= 0) { // the bounce loop never runs: Trace returns black
        st.sample[pi] = make_float4(0.0f * filterWeight, 0.0f * filterWeight, 0.0f * filterWeight, filterWeight);
        st.sampleSeed[pi] = rng.seed;
    }
}

// ---------------------------------------------------------------------- extend
// Persistent warps with per-lane dynamic ray fetch. A lane whose ray has finished does not
// wait for the slowest ray of its warp: once fewer than REFILL_THRESHOLD lanes are still
// traversing, the warp regroups, idle lanes grab the next rays of the queue with one
// warp-aggregated atomic, and traversal resumes. (A static assignment ran at 3.5 active
// lanes per instruction on the incoherent bounces of the Teapot scene — profiles/.)
#ifndef EXTEND_MIN_BLOCKS
#define EXTEND_MIN_BLOCKS 8
#endif
#ifndef SHADE_MIN_BLOCKS
#define SHADE_MIN_BLOCKS 8
#endif
#define REFILL_THRESHOLD 14
// Step budgets: a ray that is still traversing after `budget` node visits in one kernel is
// suspended (Traversal::suspend) and resumed by the next k_extend_resume round, where it shares
// a warp with other long rays instead of pinning a 1-active-lane warp of the main kernel.
// On the Teapot scene the median ray needs 8 visits, the 99th percentile 155 and the maximum
// about 6600 (triangle fans at the teapot's poles).
#define EXTEND_BUDGET_MAIN 160u
#define EXTEND_RESUME_ROUNDS 3

// returns true when the ray hit something (the path then goes to the hit queue, otherwise to the miss queue)
__device__ __forceinline__ bool write_hit(PathState& st, const Traversal& tr, uint32_t pi, int bounceIsZero, uint32_t outputHeatmap,
                                          uint32_t aovMask, uint32_t& rays, uint32_t& ntris, uint32_t& nboxes) {
    HitRec h;
    tr.result(h);
    if (h.t >= 0.0f) { // a miss needs no hit record: k_shade_miss only reads the ray
        st.hit[pi] = make_float4(h.t, h.b1, h.b2, __uint_as_float(h.prim));
        st.hitGeom[pi] = h.geom;
    }
    if (aovMask & AOV_FULL) {
        uint2 c = st.counters[pi];
        c.x += h.tris; c.y += h.boxes;
        st.counters[pi] = c;
        if (bounceIsZero) st.primaryHit[pi] = make_uint2(h.geom, h.prim);
        if (outputHeatmap) st.aovAlbedo[pi] = make_float4((float)h.tris, (float)h.boxes, 0.0f, 0.0f);
    }
    rays++; ntris += h.tris; nboxes += h.boxes;
    return h.t >= 0.0f;
}

// Material class of a hit: the key of the shading stage's queue (north star: "hit queues ... sorted by material to cut
// divergence"). The divergent code of the bounce is selected by the material's flags (kernel.glsl:1394-1417 lobe
// choice, :1519-1697 specular / subsurface / diffuse branches; RayGenCommon.h:298-341 mix and textured fetch), so
// the key is those flag bits plus "albedo comes from a texture": 6 bits, 64 classes.
#define TB_CLASS_BITS 6
#define TB_CLASS_SHIFT 26            // a queue entry is pixel | class << 26 (tb_resize caps the image at 2^26 pixels)
#define TB_PIXEL_MASK 0x03ffffffu
#define TB_CLASS_COUNT_BASE 16       // queueCount[16 .. 79]: hits per class of the current bounce, [80 .. 143]: scatter cursors
// The class of every geometry is tabulated by the host when the scene (or a material) changes: (Flags & 0x1f) |
// (albedo textured ? 0x20 : 0) of the geometry's material, i.e. METALLIC 1, SUBSURFACE 2, NO_SPECULAR 4, MIX 8, LIGHT 16,
// textured 32. One byte load from a table that lives in L1 (two dependent loads through the geometry and material
// records stalled the whole warp's service phase).
__device__ __forceinline__ uint32_t material_class(const uint8_t* __restrict__ classOfGeom, uint32_t geom) { return __ldg(classOfGeom + geom); }

// hit / miss queues of the bounce: counters [6 + 2*qi] (hits) and [7 + 2*qi] (misses)
__device__ __forceinline__ void push_sorted(PathState& st, int qi, uint32_t pi, bool hit, uint32_t cls, bool classSort) {
    uint32_t slot = atomicAdd(&st.queueCount[6 + 2 * qi + (hit ? 0 : 1)], 1u);
    if (hit) {
        st.hitQueue[slot] = pi | (cls << TB_CLASS_SHIFT);
        if (classSort) atomicAdd(&st.queueCount[TB_CLASS_COUNT_BASE + cls], 1u);
    } else st.missQueue[slot] = pi;
}

// stats layout (TB_STATS_WORDS 64-bit words): [0..2] rays / boxes / tris finished in k_extend<EXT_MAIN>, [3..5] in the
// shadow / walk / inline traversals, [6..8] in k_extend_resume, [16 + b] rays of bounce b of any kind (b clamped to 31)
__device__ __forceinline__ void flush_stats(PathState& st, int slot, int bounce, uint32_t rays, uint32_t ntris, uint32_t nboxes) {
    for (int o = 16; o > 0; o >>= 1) {
        rays += __shfl_xor_sync(0xffffffffu, rays, o); ntris += __shfl_xor_sync(0xffffffffu, ntris, o); nboxes += __shfl_xor_sync(0xffffffffu, nboxes, o);
    }
    if ((threadIdx.x & 31) == 0 && rays) {
        atomicAdd(&st.stats[slot], (unsigned long long)rays); atomicAdd(&st.stats[slot + 1], (unsigned long long)nboxes); atomicAdd(&st.stats[slot + 2], (unsigned long long)ntris);
        atomicAdd(&st.stats[16 + (bounce < 31 ? bounce : 31)], (unsigned long long)rays);
    }
}

// park the ray; returns false when the suspension buffer is full (caller keeps traversing)
__device__ __forceinline__ bool try_suspend(PathState& st, int round, const Traversal& tr, const uint32_t* stack, uint32_t pi) {
    uint32_t slot = atomicAdd(&st.susCount[round], 1u);
    if (slot >= st.susCapacity) { atomicSub(&st.susCount[round], 1u); return false; }
    tr.suspend(st.susBuf[round & 1] + (size_t)slot * Traversal::kRecordWords, stack, pi);
    return true;
}

// KIND 0 (EXT_MAIN): the bounce's extension rays (queue qi -> st.hit, hit / miss queues, suspension).
// KIND 1 (EXT_SHADOW): the next-event shadow feelers queued by k_shade<0> (st.shadowQueue, st.shRayO/D -> st.shHit).
// KIND 2 (EXT_WALK): the current rays of the glass / subsurface walkers of walk queue `qi` (0/1) -> st.hit, read by
//         k_walk_step. Kinds 1 and 2 have no suspension, no AOVs and no output queues.
#define EXT_MAIN 0
#define EXT_SHADOW 1
#define EXT_WALK 2
template <int KIND>
__global__ void __launch_bounds__(128, EXTEND_MIN_BLOCKS) k_extend(DeviceBvh bvh, PathState st, int qi, int bounce, uint32_t outputHeatmap, const FrameConstants* __restrict__ fcp, uint32_t budgetMain, uint32_t refillBelow,
                                                                   const uint8_t* __restrict__ classOfGeom) {
    const bool classSort = KIND == EXT_MAIN && classOfGeom != nullptr; // hits carry their material class and are counted per class
    // per-block class histogram: the scene's hits fall into a handful of classes, so counting them in global memory would
    // be millions of same-address atomics per bounce (measured: -13 % on Teapot); shared counters, flushed once per block
    __shared__ uint32_t s_classCount[KIND == EXT_MAIN ? (1u << TB_CLASS_BITS) : 1u];
    if (classSort) {
        if (threadIdx.x < (1u << TB_CLASS_BITS)) s_classCount[threadIdx.x] = 0;
        __syncthreads();
    }
    const uint32_t aovMask = fcp->aovMask;
    const int bounceIsZero = bounce == 0;
    const uint32_t count = KIND == EXT_SHADOW ? st.queueCount[4] : KIND == EXT_WALK ? st.queueCount[10 + qi] : st.queueCount[qi];
    if (KIND == EXT_MAIN && blockIdx.x == 0 && threadIdx.x == 0) {
        st.queueCount[qi ^ 1] = 0;                       // next queue starts empty (consumed by k_shade)
        st.queueCount[4] = 0; st.queueCount[5] = 0;      // shadow queue of this bounce + its work counter
        st.queueCount[10] = 0; st.queueCount[12] = 0;    // walk queue 0 of this bounce (filled by k_shade) + its work counter
    }
    if (KIND == EXT_WALK && blockIdx.x == 0 && threadIdx.x == 0) {
        st.queueCount[10 + (qi ^ 1)] = 0; st.queueCount[12 + (qi ^ 1)] = 0; // the other walk queue: filled by the k_walk_step after this kernel
    }
    uint32_t* __restrict__ next = KIND == EXT_SHADOW ? &st.queueCount[5] : KIND == EXT_WALK ? &st.queueCount[12 + qi] : &st.queueCount[2 + qi]; // work counter, zeroed by an earlier kernel
    const uint32_t* __restrict__ queue = KIND == EXT_SHADOW ? st.shadowQueue : KIND == EXT_WALK ? st.walkQueue[qi] : st.queue[qi];
    const float4* __restrict__ rayO = KIND == EXT_SHADOW ? st.shRayO : st.rayO;
    const float4* __restrict__ rayD = KIND == EXT_SHADOW ? st.shRayD : st.rayD;
    const float4* __restrict__ pairs = (const float4*)bvh.pairs;
    const float4* __restrict__ tris = (const float4*)bvh.tris;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t rays = 0, ntris = 0, nboxes = 0;
    Traversal tr;
    uint32_t stack[TB_STACK_WORDS];
    tr.idle(stack);
    bool haveRay = false, exhausted = false; // haveRay: this lane owns a ray (traversing or finished, not yet retired)
    uint32_t pi = 0, suspendAt = budgetMain; // suspendAt: node visits (tr.steps()) after which the ray is parked
    while (true) {
        // ---- service phase (whole warp): retire finished rays, park rays over budget, refill idle lanes
        bool retired = false, retiredHit = false;
        if (haveRay && tr.done()) {
            if (KIND != EXT_MAIN) {
                HitRec h;
                tr.result(h);
                float4 rec = make_float4(h.t, h.b1, h.b2, __uint_as_float(h.prim));
                if (KIND == EXT_SHADOW) { st.shHit[pi] = rec; st.shHitGeom[pi] = h.geom; }
                else { st.hit[pi] = rec; st.hitGeom[pi] = h.geom; }
                if (aovMask & AOV_FULL) { uint2 c = st.counters[pi]; c.x += h.tris; c.y += h.boxes; st.counters[pi] = c; }
                rays++; ntris += h.tris; nboxes += h.boxes;
            } else {
                retiredHit = write_hit(st, tr, pi, bounceIsZero, outputHeatmap, aovMask, rays, ntris, nboxes);
                retired = true;
            }
            haveRay = false;
        } else if (KIND == EXT_MAIN && haveRay && tr.steps() >= suspendAt) {
            if (try_suspend(st, 0, tr, stack, pi)) { haveRay = false; tr.cur = TB_NO_NODE; }
            else suspendAt += budgetMain; // buffer full: keep going here
        }
        if (KIND == EXT_MAIN) { // sort retired paths into the hit / miss queues (material-class split of the shading stage)
            uint32_t mh = __ballot_sync(0xffffffffu, retired && retiredHit), mm = __ballot_sync(0xffffffffu, retired && !retiredHit);
            uint32_t cls = 0;
            if (classSort && retired && retiredHit) { // one counter update per distinct class per warp
                cls = material_class(classOfGeom, tr.hitGeom);
                const uint32_t peers = __match_any_sync(mh, cls);
                if (lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&s_classCount[cls], (uint32_t)__popc(peers));
            }
            if (mh | mm) {
                uint32_t bh = 0, bm = 0;
                if (lane == 0) {
                    if (mh) bh = atomicAdd(&st.queueCount[6 + 2 * qi], (uint32_t)__popc(mh));
                    if (mm) bm = atomicAdd(&st.queueCount[7 + 2 * qi], (uint32_t)__popc(mm));
                }
                bh = __shfl_sync(0xffffffffu, bh, 0); bm = __shfl_sync(0xffffffffu, bm, 0);
                const uint32_t below = (1u << lane) - 1u;
                if (retired && retiredHit) st.hitQueue[bh + __popc(mh & below)] = pi | (cls << TB_CLASS_SHIFT);
                if (retired && !retiredHit) st.missQueue[bm + __popc(mm & below)] = pi;
            }
        }
        uint32_t idle = __ballot_sync(0xffffffffu, !haveRay);
        if (idle && !exhausted) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(next, (uint32_t)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (!haveRay) {
                uint32_t i = base + __popc(idle & ((1u << lane) - 1u));
                if (i < count) {
                    pi = __ldg(queue + i);
                    float4 o = rayO[pi], d = rayD[pi];
                    tr.begin(bvh, stack, mk3(o.x, o.y, o.z), mk3(d.x, d.y, d.z), MIN_T, FAR_T);
                    haveRay = true;
                    suspendAt = budgetMain;
                }
            }
            if (base + (uint32_t)__popc(idle) >= count) exhausted = true;
        }
        if (!__any_sync(0xffffffffu, haveRay)) break;
        // ---- traversal phase: every iteration the whole warp executes ONE of the two code paths, the
        // one more lanes are waiting for (box-pair test vs. triangle test), so the two never serialise
        // inside an iteration; lanes on the minority path wait and soon become the majority. The phase
        // ends when enough lanes have finished to make a service phase worthwhile.
        // Once the queue is exhausted there is nothing to refill with, so the phase would only end when the last
        // ray does; it is cut every 64 iterations instead so that the service phase can park rays over budget.
        // (A lane without a ray always has tr.cur == TB_NO_NODE, so `cur` alone says what the lane wants.)
        uint32_t minBusy = exhausted ? 1u : max(refillBelow, 1u);
        uint32_t left = (KIND == EXT_MAIN && exhausted) ? 64u : 0xffffffffu;
        asm volatile("" : "+r"(minBusy), "+r"(left)); // loop constants live in registers (ptxas re-derived them every iteration)
        while (true) {
            const bool busy = !tr.done();
            const bool wantLeaf = (int32_t)tr.cur < 0;
            const uint32_t nB = __popc(__ballot_sync(0xffffffffu, busy)), nL = __popc(__ballot_sync(0xffffffffu, wantLeaf));
            const bool stop = (nB < minBusy) | (left == 0u);
            left--;
            if (stop) break;
            if (2 * nL > nB) {
                if (wantLeaf) tr.step_leaf(stack, tris);
            } else {
                if (busy && !wantLeaf) tr.step_internal(stack, pairs);
            }
        }
    }
    flush_stats(st, KIND == EXT_MAIN ? 0 : 3, bounce, rays, ntris, nboxes);
    if (classSort) {
        __syncthreads();
        if (threadIdx.x < (1u << TB_CLASS_BITS) && s_classCount[threadIdx.x]) atomicAdd(&st.queueCount[TB_CLASS_COUNT_BASE + threadIdx.x], s_classCount[threadIdx.x]);
    }
}

// Resume round `round` (1-based): continues the rays parked by round-1; the last round has no budget.
__global__ void __launch_bounds__(128) k_extend_resume(DeviceBvh bvh, PathState st, int qi, int round, uint32_t budget, int bounce,
                                                       uint32_t outputHeatmap, const FrameConstants* __restrict__ fcp,
                                                       const uint8_t* __restrict__ classOfGeom) {
    const bool classSort = classOfGeom != nullptr;
    const uint32_t aovMask = fcp->aovMask;
    const int bounceIsZero = bounce == 0;
    const uint32_t count = min(st.susCount[round - 1], st.susCapacity);
    const float4* __restrict__ pairs = (const float4*)bvh.pairs;
    const float4* __restrict__ tris = (const float4*)bvh.tris;
    const uint32_t* __restrict__ in = st.susBuf[(round - 1) & 1];
    uint32_t rays = 0, ntris = 0, nboxes = 0;
    Traversal tr;
    uint32_t stack[TB_STACK_WORDS];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        uint32_t pi = tr.resume(bvh, in + (size_t)i * Traversal::kRecordWords, stack, st.rayO, st.rayD, MIN_T, FAR_T);
        uint32_t suspendAt = tr.steps() + budget;
        while (true) {
            if (tr.done()) {
                const bool hit = write_hit(st, tr, pi, bounceIsZero, outputHeatmap, aovMask, rays, ntris, nboxes);
                push_sorted(st, qi, pi, hit, (hit && classSort) ? material_class(classOfGeom, tr.hitGeom) : 0u, classSort);
                break;
            }
            if (budget && tr.steps() >= suspendAt) {
                if (try_suspend(st, round, tr, stack, pi)) break;
                suspendAt += budget;
            }
            tr.step(stack, pairs, tris);
        }
    }
    flush_stats(st, 6, bounce, rays, ntris, nboxes);
}

// ------------------------------------------------------------------ ray sort
// Spatial sort of a bounce's ray queue (counting sort by the Morton cell of the ray origin, 5 bits per axis inside
// the scene box). The queue order k_shade leaves behind is the order in which rays retired from the previous
// traversal, i.e. random within a window of ~150 k paths; on a scene whose BVH is far larger than L2 every lane of a
// warp then walks its own part of the tree through DRAM. Sorted, the rays a warp (and its neighbours in time) pick up
// start in the same cell and share the lower levels of the tree in L1 / L2. Path results do not depend on queue
// order (every path is independent and accumulation is ordered per frame), so this changes timing only.
__device__ __forceinline__ uint32_t spread5(uint32_t v) { // bit k of a 5-bit value -> bit 3k
    return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4) | ((v & 8u) << 6) | ((v & 16u) << 8);
}
// Rays of one warp mostly fall into a few cells (that is the point of sorting them), so the cell counters are
// updated once per distinct key per warp (__match_any_sync), not once per ray: plain per-ray atomics serialised
// millions of same-address operations on scenes whose rays crowd a few cells.
__global__ void __launch_bounds__(256) k_sort_keys(DeviceBvh bvh, PathState st, const uint32_t* __restrict__ queue, const uint32_t* __restrict__ countPtr,
                                                   const float4* __restrict__ rayO) {
    const uint32_t count = *countPtr;
    const uint32_t lane = threadIdx.x & 31;
    const float sx = 16.0f / fmaxf(bvh.root.h[0], 1e-30f), sy = 16.0f / fmaxf(bvh.root.h[1], 1e-30f), sz = 16.0f / fmaxf(bvh.root.h[2], 1e-30f);
    const float ox = bvh.root.c[0] - bvh.root.h[0], oy = bvh.root.c[1] - bvh.root.h[1], oz = bvh.root.c[2] - bvh.root.h[2];
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < count; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + lane;
        uint32_t key = 0xffffffffu;
        if (i < count) {
            const uint32_t pi = queue[i];
            const float4 o = rayO[pi];
            // NaN / out-of-box origins land in cell 0 or 31 of the axis: any cell is fine, the key only orders work
            const uint32_t cx = (uint32_t)fminf(fmaxf((o.x - ox) * sx, 0.0f), 31.0f);
            const uint32_t cy = (uint32_t)fminf(fmaxf((o.y - oy) * sy, 0.0f), 31.0f);
            const uint32_t cz = (uint32_t)fminf(fmaxf((o.z - oz) * sz, 0.0f), 31.0f);
            key = spread5(cx) | (spread5(cy) << 1) | (spread5(cz) << 2);
            st.sortKeys[i] = key;
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, key);
        if (i < count && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&st.sortHist[key], (uint32_t)__popc(peers));
    }
}
// exclusive scan of the TB_SORT_CELLS counters in place (one block of 1024 threads, 32 cells per thread)
__global__ void __launch_bounds__(1024) k_sort_scan(PathState st) {
    __shared__ uint32_t s_warp[32];
    const uint32_t t = threadIdx.x, per = TB_SORT_CELLS / 1024u;
    uint32_t local[TB_SORT_CELLS / 1024u];
    uint32_t sum = 0;
#pragma unroll
    for (uint32_t k = 0; k < per; k++) { local[k] = st.sortHist[t * per + k]; sum += local[k]; }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if ((t & 31) >= (uint32_t)o) incl += v; }
    if ((t & 31) == 31) s_warp[t >> 5] = incl;
    __syncthreads();
    if (t < 32) {
        uint32_t w = s_warp[t], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, wi, o); if (t >= (uint32_t)o) wi += v; }
        s_warp[t] = wi - w;
    }
    __syncthreads();
    uint32_t run = s_warp[t >> 5] + incl - sum;
#pragma unroll
    for (uint32_t k = 0; k < per; k++) { st.sortHist[t * per + k] = run; run += local[k]; }
}
__global__ void __launch_bounds__(256) k_sort_scatter(PathState st, const uint32_t* __restrict__ queue, const uint32_t* __restrict__ countPtr) {
    const uint32_t count = *countPtr;
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < count; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + lane;
        const uint32_t key = i < count ? st.sortKeys[i] : 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, key);
        const uint32_t leader = (uint32_t)(__ffs(peers) - 1);
        uint32_t pos = 0;
        if (i < count && lane == leader) pos = atomicAdd(&st.sortHist[key], (uint32_t)__popc(peers));
        pos = __shfl_sync(0xffffffffu, pos, leader) + (uint32_t)__popc(peers & ((1u << lane) - 1u));
        if (i < count) st.sortTmp[pos] = queue[i];
    }
}

// Counting sort of the bounce's hit queue by material class: the histogram was filled by k_extend as the rays
// retired, so one pass is left (every block derives the class offsets from the 64 counters itself). A warp of the
// shading stage then works on hits of ONE class: the same branch of the lobe choice / specular / subsurface /
// diffuse code, the same texture path. Order inside a class is arbitrary (every path is independent).
#define CS_ITEMS 8 // queue entries per thread: a block of 256 threads sorts a chunk of 2048
__global__ void __launch_bounds__(256) k_class_scatter(PathState st, int qi) {
    __shared__ uint32_t s_start[1u << TB_CLASS_BITS]; // first slot of the class in the sorted queue
    __shared__ uint32_t s_count[1u << TB_CLASS_BITS]; // this chunk's entries per class, then the chunk's cursor inside the class
    __shared__ uint32_t s_base[1u << TB_CLASS_BITS];  // where this chunk's entries of the class start (one global atomic per class per block)
    const uint32_t count = st.queueCount[6 + 2 * qi];
    uint32_t* __restrict__ cursor = st.queueCount + TB_CLASS_COUNT_BASE + (1u << TB_CLASS_BITS);
    const uint32_t chunk = 256u * CS_ITEMS;
    if (threadIdx.x < (1u << TB_CLASS_BITS)) {
        uint32_t run = 0;
        for (uint32_t c = 0; c < threadIdx.x; c++) run += st.queueCount[TB_CLASS_COUNT_BASE + c];
        s_start[threadIdx.x] = run;
    }
    for (uint32_t first = blockIdx.x * chunk; first < count; first += gridDim.x * chunk) {
        if (threadIdx.x < (1u << TB_CLASS_BITS)) s_count[threadIdx.x] = 0;
        __syncthreads();
        // one shared-memory atomic per distinct class per warp (a warp's 32 entries fall into one or two classes)
        const uint32_t lane = threadIdx.x & 31;
        uint32_t e[CS_ITEMS];
#pragma unroll
        for (int k = 0; k < CS_ITEMS; k++) {
            const uint32_t i = first + k * 256u + threadIdx.x;
            e[k] = i < count ? st.hitQueue[i] : 0xffffffffu;
            const uint32_t key = i < count ? (e[k] >> TB_CLASS_SHIFT) : 0xffffffffu;
            const uint32_t peers = __match_any_sync(0xffffffffu, key);
            if (i < count && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&s_count[key], (uint32_t)__popc(peers));
        }
        __syncthreads();
        if (threadIdx.x < (1u << TB_CLASS_BITS)) {
            const uint32_t c = s_count[threadIdx.x];
            s_base[threadIdx.x] = c ? atomicAdd(&cursor[threadIdx.x], c) : 0u;
            s_count[threadIdx.x] = 0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CS_ITEMS; k++) {
            const uint32_t i = first + k * 256u + threadIdx.x;
            const uint32_t key = i < count ? (e[k] >> TB_CLASS_SHIFT) : 0xffffffffu;
            const uint32_t peers = __match_any_sync(0xffffffffu, key);
            const uint32_t leader = (uint32_t)(__ffs(peers) - 1);
            uint32_t pos = 0;
            if (i < count && lane == leader) pos = atomicAdd(&s_count[key], (uint32_t)__popc(peers));
            pos = __shfl_sync(0xffffffffu, pos, leader) + (uint32_t)__popc(peers & ((1u << lane) - 1u));
            if (i < count) st.hitSorted[s_start[key] + s_base[key] + pos] = e[k] & TB_PIXEL_MASK;
        }
        __syncthreads();
    }
}

// ----------------------------------------------------------------------- shade
// End of a path: firefly clamp + filter weight (kernel.glsl:1907-1920) and NaN rejection
// (RayGenCommon.h:704-707). The sample and the path's rand() seed are staged per frame;
// k_accumulate applies them to OutputTexture in frame order.
__device__ __forceinline__ void finish_path(const FrameConstants& fc, PathState& st, uint32_t pi, f3 color, float filterWeight, Rng& rng) {
    if (fc.settings.FireflyClampValue >= EPSILON) color = min3(color, fc.settings.FireflyClampValue);
    f4 c = mk4(color * filterWeight, filterWeight);
    f4 outc = mk4(0, 0, 0, 0);
    if (!isnan_(c.x) && !isnan_(c.y) && !isnan_(c.z) && !isnan_(c.w)) outc = outc + c;
    st.sample[pi] = make_float4(outc.x, outc.y, outc.z, outc.w);
    st.sampleSeed[pi] = rng.seed;
}

// STAGE 0: every path of the bounce; a path that needs a next-event shadow ray writes the ray,
//          joins the shadow queue and is NOT advanced (no state is written for it).
// STAGE 1: the paths of the shadow queue, after k_extend<SHADOW> has traced their rays: the same
//          code from the top (same inputs, same rand() draws), now with the occluder known.
// STAGE 2: single stage with the shadow ray traced inline (small scenes, tb_set_shadow_mode(0)).
// STAGE 3: single stage for frames that cannot cast shadow rays (no lights, or NEE off).
// SSS:     the scene has subsurface / glass materials, whose random walk traces rays inline. Scenes
//          without them (checked on the host) get a kernel with no traversal code at all in stages
//          0, 1 and 3: fewer registers, no local-memory stack.
template <int STAGE, bool SSS>
__global__ void __launch_bounds__(128, SHADE_MIN_BLOCKS) k_shade(DeviceBvh bvh, DeviceScene sc, const FrameConstants* __restrict__ fcp, PathState st, int qi, int bounceIndex, int sortedHits) {
    const FrameConstants& fc = *fcp;
    const uint32_t count = STAGE == 1 ? st.queueCount[4] : st.queueCount[6 + 2 * qi];
    const uint32_t* __restrict__ inQueue = STAGE == 1 ? st.shadowQueue : (sortedHits ? st.hitSorted : st.hitQueue);
    if (STAGE != 1 && blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            st.queueCount[2 + (qi ^ 1)] = 0; // work counter of the next k_extend
            st.queueCount[6 + 2 * (qi ^ 1)] = 0; st.queueCount[7 + 2 * (qi ^ 1)] = 0; // hit / miss queues of the next bounce
            for (int r = 0; r < 4; r++) st.susCount[r] = 0; // suspension counters of the next bounce
        }
        st.queueCount[TB_CLASS_COUNT_BASE + threadIdx.x] = 0; // 128 threads: class counters + scatter cursors of the next bounce
    }
    const TbOutputSettings& S = fc.settings;
    const int MaxBounces = S.MaxBounces;
    RayCount rc = {0, 0, 0};
    const uint32_t stride = gridDim.x * blockDim.x;
    // loop bound rounded up to a warp multiple so every lane reaches the ballot below
    const uint32_t countUp = (count + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < countUp; i += stride) {
        bool alive = false;    // does this path continue to the next bounce?
        bool deferred = false; // STAGE 0: waits for its shadow ray
        bool walker = false, walkerPerfectSpec = false; // entered a subsurface / glass medium: continues in k_walk
        uint32_t pi = 0;
        if (i < count) {
            pi = inQueue[i] & TB_PIXEL_MASK; // (hit-queue entries carry the material class in their top bits)
            float4 o4 = st.rayO[pi], d4 = st.rayD[pi], t4 = st.thr[pi], c4 = st.col[pi], h4 = st.hit[pi];
            Rng rng; rng.seed = o4.w; rng.time = fc.time;
            uint32_t sw = __float_as_uint(d4.w);
            int bounce = (int)(sw & 0xffu);
            bool bPrevSpec = (sw & 0x100u) != 0;
            f3 org = mk3(o4.x, o4.y, o4.z), dir = mk3(d4.x, d4.y, d4.z);
            f3 thr = mk3(t4.x, t4.y, t4.z), acc = mk3(c4.x, c4.y, c4.z);
            const float filterWeight = t4.w;
            const bool bFirstRay = bounce == 0;
            bool terminated = false;
            uint2 cnt0 = make_uint2(rc.tris, rc.boxes);

            do { // one iteration of the bounce loop, `break` = path ends
                if (thr.x < EPSILON && thr.y < EPSILON && thr.z < EPSILON) { terminated = true; break; }
                if (h4.x < 0.0f) { // miss
                    acc += thr * sample_environment_map(sc, dir);
                    if (bFirstRay) st.stEmissive[pi] = make_float4(acc.x, acc.y, acc.z, 1.0f);
                    terminated = true; break;
                }
                Surface sf = surface_from_hit(sc, h4.y, h4.z, st.hitGeom[pi], __float_as_uint(h4.w));
                f3 normal = sf.normal;
                f3 RayPoint = org + dir * h4.x;
                org = RayPoint + normal * EPSILON;
                float RdotN = dot(normal, dir);
                bool backside = RdotN > 0.0f;
                Mat material = get_material(sc, rng, sf.material, sf.uv, backside);
                f3 detailNormal = get_detail_normal(sc, fc, material, normal, sf.tangent, sf.uv);
                if (bFirstRay) {
                    if (fc.aovMask & AOV_WORLDPOS) {
                        float4 nb = st.neighbor[pi], nd = st.neighborDir[pi];
                        f3 nrp = mk3(nb.x, nb.y, nb.z) + mk3(nd.x, nd.y, nd.z) * h4.x;
                        f3 wp = mk3(0.0f) + RayPoint;
                        float dn = 0.0f + length(nrp - RayPoint);
                        st.aovWorldPos[fc.worldPosSlot][pi] = make_float4(wp.x, wp.y, wp.z, dn);
                    }
                    if (fc.aovMask & AOV_FULL) st.aovNormal[pi] = make_float4(detailNormal.x, detailNormal.y, detailNormal.z, 1.0f);
                    st.stDepth[pi] = saturate(h4.x / S.MaxZ);
                    if ((fc.aovMask & AOV_FULL) && (int)(pi % fc.width) == fc.selectedX && (int)(pi / fc.width) == fc.selectedY) {
                        st.readbackStats->SelectedPixelDistance = h4.x;
                        st.readbackStats->SelectedMaterialID = sf.material;
                    }
                    if (S.OutputType == TB_OUTPUT_HEATMAP) { terminated = true; break; }
                }
                float CurrentIOR = backside ? material.IOR : AIR_IOR;
                float NewIOR = backside ? AIR_IOR : material.IOR;
                if (backside) { normal = -normal; RdotN = -RdotN; detailNormal = -detailNormal; }
                float ReflectionCoefficient = material.SpecularCoef;
                bool bSpecularRay = false;
                if (AllowsSpecular(material)) {
                    if (IsMetallic(material) || IsHair(material)) bSpecularRay = true;
                    else bSpecularRay = rng.next() < 0.5f;
                }
                bool bPerfectSpec = bSpecularRay && material.roughness < 0.05f;
                if (bPrevSpec || bFirstRay || !IsLight(material) || !S.EnableNextEventEstimation) acc += thr * material.emissive;
                if (IsLight(material)) { terminated = true; break; }

                float lightPDF = 0.0f, lightAttenuation = 0.0f;
                f3 lightDirection = mk3(0.0f), lightColor = mk3(0.0f), lightNormal = mk3(0.0f);
                if (STAGE != 3) get_one_light_sample(sc, fc, rng, RayPoint, lightDirection, lightColor, lightPDF, lightNormal, lightAttenuation);
                if (STAGE != 3 && !bPerfectSpec && lightPDF > EPSILON && dot(lightDirection, lightNormal) < 0.0f) {
                    f3 ShadowMultiplier = mk3(1.0f);
                    Surface ss; float stt;
                    bool occluded;
                    if (STAGE == 0) { // hand the shadow feeler to the traversal stage and stop here
                        f3 so = RayPoint + normal * EPSILON;
                        st.shRayO[pi] = make_float4(so.x, so.y, so.z, 0.0f);
                        st.shRayD[pi] = make_float4(lightDirection.x, lightDirection.y, lightDirection.z, 0.0f);
                        deferred = true;
                        break;
                    } else if (STAGE == 1) {
                        float4 sh = st.shHit[pi];
                        occluded = sh.x >= 0.0f;
                        if (occluded) ss = surface_from_hit(sc, sh.y, sh.z, st.shHitGeom[pi], __float_as_uint(sh.w));
                    } else if (STAGE == 2) {
                        occluded = intersect_inline(bvh, sc, RayPoint + normal * EPSILON, lightDirection, rc, ss, stt);
                    } else occluded = false;
                    if (occluded) {
                        bool shBack = dot(ss.normal, lightDirection) > 0.0f;
                        Mat sm = get_material(sc, rng, ss.material, ss.uv, shBack);
                        if (!IsLight(sm)) ShadowMultiplier = mk3(0.0f);
                    }
                    float lightMultiplier = lightAttenuation * diffuse_brdf(lightDirection, detailNormal) * fabsf(dot(lightNormal, lightDirection)) / lightPDF;
                    acc += (((thr * material.albedo) * lightMultiplier) * ShadowMultiplier) * lightColor;
                }

                f3 previousDirection = dir;
                bPrevSpec = bPerfectSpec;
                bool skipBrdf = false;
                if (bSpecularRay) {
                    // ImportanceSampleGGX, kernel.glsl:1066-1082
                    float roughness = fmaxf(MIN_ROUGHNESS, material.roughness);
                    float a = roughness * roughness;
                    float a2 = a * a;
                    float u1 = rng.next(), u2 = rng.next();
                    float theta = 2.0f * PI_K * u2;
                    float phi = acos_(sqrtf((1.0f - u1) / ((a2 - 1.0f) * u1 + 1.0f)));
                    f3 dd = mk3(sin_(phi) * cos_(theta), cos_(phi), sin_(phi) * sin_(theta));
                    f3 hh = reorient_around_normal(dd, normal);
                    dir = reflect(dir, hh);
                } else if (SSS && IsSSS(material)) {
                    float nr = CurrentIOR / NewIOR;
                    if (refract_or_reflect(rng, dir, normal, nr, RdotN, bPerfectSpec, material.roughness, bPrevSpec) == GIVE_UP) { terminated = true; break; }
                    bool noScatter = material.scattering.x < EPSILON;
                    float DistancePerScatter = 1.0f / (((material.scattering.x + material.scattering.y) + material.scattering.z) / 3.0f);
                    float maxTravelDistance = noScatter ? LARGE_NUMBER : DistancePerScatter;
                    bool exitting = (material.Flags & TB_SINGLE_SIDED_MATERIAL_FLAG) != 0;
                    if (!exitting) {
                        // The random walk inside the medium (kernel.glsl:1571-1688) traces up to 100 rays. It runs in
                        // its own stages (k_extend<EXT_WALK> + k_walk_step, then k_walk), in warps made of walkers only:
                        // inline here it ran at 3.2 active lanes per instruction on the vw-van scene (profiles/).
                        // Hand over what the walk reads besides the path state.
                        float travelDistance = fmaxf(-log_(rng.next()), 0.1f) * maxTravelDistance; // first iteration, kernel.glsl:1573
                        st.walkA[pi] = make_float4(material.absorption.x, material.absorption.y, material.absorption.z, maxTravelDistance);
                        st.walkB[pi] = make_float4(CurrentIOR, NewIOR, material.roughness, travelDistance);
                        walker = true; walkerPerfectSpec = bPerfectSpec;
                        break;
                    }
                    skipBrdf = true; // `continue`, kernel.glsl:1690
                } else {
                    // GenerateCosineWeightedDirection, kernel.glsl:1025-1046
                    float rand0 = rng.next();
                    float rand1 = rng.next();
                    float r = sqrtf(rand0);
                    float theta = 2.0f * PI_K * rand1;
                    float x = r * cos_(theta);
                    float y = sqrtf(fmaxf(EPSILON, 1.0f - rand0));
                    float z = r * sin_(theta);
                    dir = reorient_around_normal(mk3(x, y, z), normal);
                }
                if (!skipBrdf) {
                    float DiffusePDF = dot(dir, normal) / PI_K;
                    if (AllowsSpecular(material)) {
                        f3 halfVector = half_vector_safe(-previousDirection, dir, normal);
                        float SpecularPDF = ggx_pdf(normal, dir, halfVector, material.roughness);
                        float PDFValue = IsMetallic(material) ? SpecularPDF : lerp(SpecularPDF, DiffusePDF, 0.5f);
                        thr /= PDFValue;
                    } else thr /= DiffusePDF;
                    if (bFirstRay) st.stEmissive[pi] = make_float4(material.emissive.x, material.emissive.y, material.emissive.z, 1.0f);
                    bool bRemoveAlbedo = (S.RenderMode == TB_RENDER_REALTIME) && bFirstRay;
                    f3 albedo = bRemoveAlbedo ? mk3(1.0f) : material.albedo;
                    if (IsMetallic(material)) {
                        f3 halfVector = normalize(-previousDirection + dir);
                        float roughnessSquared = fmaxf(material.roughness * material.roughness, MIN_ROUGHNESS_SQUARED);
                        float specular = ggx_ndf(detailNormal, halfVector, roughnessSquared) /
                                         ((4.0f * fabsf(dot(-previousDirection, halfVector))) * fmaxf(fabsf(dot(-previousDirection, normal)), fabsf(dot(dir, normal))));
                        thr *= (specular * albedo) * saturate(dot(dir, normal));
                    } else if (AllowsSpecular(material)) {
                        f3 halfVector = half_vector_safe(-previousDirection, dir, normal);
                        float fresnel = ReflectionCoefficient + (1.0f - ReflectionCoefficient) * pow_(fabsf(1.0f - dot(-previousDirection, halfVector)), 5.0f);
                        float diffuseMultiplier = (((28.0f / (23.0f * PI_K)) * (1.0f - ReflectionCoefficient)) *
                                                   (1.0f - pow_(1.0f - 0.5f * dot(-previousDirection, normal), 5.0f))) *
                                                  (1.0f - pow_(1.0f - 0.5f * dot(dir, normal), 5.0f));
                        f3 diffuse = albedo * diffuseMultiplier;
                        float roughnessSquared = fmaxf(material.roughness * material.roughness, MIN_ROUGHNESS_SQUARED);
                        float specular = ggx_ndf(detailNormal, halfVector, roughnessSquared) /
                                         ((4.0f * fabsf(dot(-previousDirection, halfVector))) * fmaxf(fabsf(dot(-previousDirection, normal)), fabsf(dot(dir, normal))));
                        thr *= (diffuse + fresnel * specular) * saturate(dot(dir, normal));
                    } else {
                        thr *= albedo * diffuse_brdf(dir, detailNormal);
                    }
                    if (bFirstRay && (fc.aovMask & AOV_FULL) && S.OutputType != TB_OUTPUT_HEATMAP) st.aovAlbedo[pi] = make_float4(material.albedo.x, material.albedo.y, material.albedo.z, 1.0f);
                }
                // top of the next loop iteration: bounce limit, then russian roulette (kernel.glsl:1286-1302)
                int next = bounce + 1;
                if (next >= MaxBounces) { terminated = true; break; }
                if (next >= 2) {
                    float p = fmaxf(fmaxf(thr.x, thr.y), thr.z);
                    p = fmaxf(p, EPSILON);
                    if (p < rng.next()) { terminated = true; break; }
                    thr *= 1.0f / p;
                }
                bounce = next;
            } while (false);

            // per-pixel counters for the inline rays of this stage
            if ((fc.aovMask & AOV_FULL) && (rc.tris != cnt0.x || rc.boxes != cnt0.y)) {
                uint2 c = st.counters[pi];
                c.x += rc.tris - cnt0.x; c.y += rc.boxes - cnt0.y;
                st.counters[pi] = c;
            }
            if (deferred) { /* nothing is written: STAGE 1 redoes this path from the same inputs */ }
            else if (walker) {
                st.rayO[pi] = make_float4(org.x, org.y, org.z, rng.seed);
                st.rayD[pi] = make_float4(dir.x, dir.y, dir.z, __uint_as_float(pack_state(bounce, bPrevSpec) | (walkerPerfectSpec ? 0x200u : 0u)));
                st.thr[pi] = make_float4(thr.x, thr.y, thr.z, filterWeight);
                st.col[pi] = make_float4(acc.x, acc.y, acc.z, 0.0f);
            }
            else if (terminated) finish_path(fc, st, pi, acc, filterWeight, rng);
            else {
                st.rayO[pi] = make_float4(org.x, org.y, org.z, rng.seed);
                st.rayD[pi] = make_float4(dir.x, dir.y, dir.z, __uint_as_float(pack_state(bounce, bPrevSpec)));
                st.thr[pi] = make_float4(thr.x, thr.y, thr.z, filterWeight);
                st.col[pi] = make_float4(acc.x, acc.y, acc.z, 0.0f);
                alive = true;
            }
        }
        // queue compaction: one atomic per warp
        uint32_t ballot = __ballot_sync(0xffffffffu, alive);
        if (ballot) {
            uint32_t lane = threadIdx.x & 31, base = 0;
            if (lane == 0) base = atomicAdd(&st.queueCount[qi ^ 1], (uint32_t)__popc(ballot));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (alive) st.queue[qi ^ 1][base + __popc(ballot & ((1u << lane) - 1u))] = pi;
        }
        if (SSS) { // walk queue, same warp-aggregated append
            uint32_t wballot = __ballot_sync(0xffffffffu, walker);
            if (wballot) {
                uint32_t lane = threadIdx.x & 31, base = 0;
                if (lane == 0) base = atomicAdd(&st.queueCount[10], (uint32_t)__popc(wballot));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (walker) st.walkQueue[0][base + __popc(wballot & ((1u << lane) - 1u))] = pi;
            }
        }
        if (STAGE == 0) { // shadow queue, same warp-aggregated append
            uint32_t dballot = __ballot_sync(0xffffffffu, deferred);
            if (dballot) {
                uint32_t lane = threadIdx.x & 31, base = 0;
                if (lane == 0) base = atomicAdd(&st.queueCount[4], (uint32_t)__popc(dballot));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (deferred) st.shadowQueue[base + __popc(dballot & ((1u << lane) - 1u))] = pi;
            }
        }
    }
    uint32_t rays = rc.rays, tris = rc.tris, boxes = rc.boxes;
    for (int o = 16; o > 0; o >>= 1) {
        rays += __shfl_xor_sync(0xffffffffu, rays, o); tris += __shfl_xor_sync(0xffffffffu, tris, o); boxes += __shfl_xor_sync(0xffffffffu, boxes, o);
    }
    if ((threadIdx.x & 31) == 0 && rays) { // slots 3..5: rays traced inside the shading stage (shadow feelers, SSS walk)
        atomicAdd(&st.stats[3], (unsigned long long)rays); atomicAdd(&st.stats[4], (unsigned long long)boxes); atomicAdd(&st.stats[5], (unsigned long long)tris);
        atomicAdd(&st.stats[16 + (bounceIndex < 31 ? bounceIndex : 31)], (unsigned long long)rays);
    }
}

// ------------------------------------------------------------------------ walk
// The random walk of a path inside a subsurface / glass medium (kernel.glsl:1571-1688), one ray per step.
// State word `sw` (rayD.w): bits 0..7 bounce, 8 prevPerfectlySpecular, 9 the bounce's perfect-specular flag, 16..23 k.
struct Walker {
    uint32_t pi, sw;
    Rng rng;
    f3 org, dir, thr, absorption;
    float maxTravelDistance, CurrentIOR, NewIOR, roughness, travelDistance;
};
__device__ __forceinline__ void walker_load(const PathState& st, const FrameConstants& fc, uint32_t pi, Walker& w) {
    float4 o = st.rayO[pi], d = st.rayD[pi], t4 = st.thr[pi], a = st.walkA[pi], b = st.walkB[pi];
    w.pi = pi;
    w.org = mk3(o.x, o.y, o.z); w.rng.seed = o.w; w.rng.time = fc.time;
    w.dir = mk3(d.x, d.y, d.z); w.sw = __float_as_uint(d.w);
    w.thr = mk3(t4.x, t4.y, t4.z);
    w.absorption = mk3(a.x, a.y, a.z); w.maxTravelDistance = a.w;
    w.CurrentIOR = b.x; w.NewIOR = b.y; w.roughness = b.z; w.travelDistance = b.w;
}
// the walker waits for its next ray: what changed since walker_load goes back to memory
__device__ __forceinline__ void walker_store(PathState& st, const Walker& w) {
    st.rayO[w.pi] = make_float4(w.org.x, w.org.y, w.org.z, w.rng.seed);
    st.rayD[w.pi] = make_float4(w.dir.x, w.dir.y, w.dir.z, __uint_as_float(w.sw));
    float4 t4 = st.thr[w.pi];
    st.thr[w.pi] = make_float4(w.thr.x, w.thr.y, w.thr.z, t4.w);
    st.walkB[w.pi] = make_float4(w.CurrentIOR, w.NewIOR, w.roughness, w.travelDistance);
}
// One iteration of the walk loop after its ray came back (t < 0: nothing hit). Returns true when the walk goes on:
// org / dir / travelDistance then describe the next ray.
__device__ bool walk_step(const DeviceScene& sc, Walker& w, float t, float b1, float b2, uint32_t geom, uint32_t prim) {
    const bool noScatter = w.maxTravelDistance == LARGE_NUMBER; // DistancePerScatter < 1/EPSILON otherwise
    const bool bPerfectSpec = (w.sw & 0x200u) != 0;
    bool bPrevSpec = (w.sw & 0x100u) != 0;
    uint32_t k = (w.sw >> 16) & 0xffu;
    bool walkOn = false;
    if (t < 0.0f) w.thr = mk3(0.0f); // left the medium without meeting a surface: `break`, kernel.glsl:1584
    else {
        Surface ws = surface_from_hit(sc, b1, b2, geom, prim);
        f3 normal = ws.normal;
        float tt = fminf(w.travelDistance, t);
        bool exitting = tt < w.travelDistance || noScatter;
        if (k == 99u && !exitting) w.thr = mk3(0.0f);
        f3 RayPoint = w.org + w.dir * tt;
        w.org = RayPoint + normal * EPSILON;
        w.thr *= exp3((-tt) * w.absorption);
        bool giveUp = false;
        if (exitting) {
            float RdotN = dot(normal, w.dir);
            if (RdotN >= 0.0f) { normal = -normal; RdotN = -RdotN; }
            RefractResult rr = refract_or_reflect(w.rng, w.dir, normal, w.NewIOR / w.CurrentIOR, RdotN, bPerfectSpec, w.roughness, bPrevSpec);
            if (rr == GIVE_UP) giveUp = true;
            if (rr == REFLECTED) exitting = false;
        } else {
            // GenerateRandomDirection(), kernel.glsl:991-999
            float u1 = w.rng.next(), u2 = w.rng.next();
            float r = sqrtf(1.0f - u1 * u1);
            float phi = 2.0f * 3.14f * u2;
            w.dir = mk3(cos_(phi) * r, sin_(phi) * r, u1);
            w.thr /= 1.0f;
        }
        k++;
        walkOn = !giveUp && k < 100u && !exitting;
    }
    w.sw = (w.sw & 0x0000feffu) | (bPrevSpec ? 0x100u : 0u) | (k << 16);
    if (walkOn) w.travelDistance = fmaxf(-log_(w.rng.next()), 0.1f) * w.maxTravelDistance;
    return walkOn;
}
// The walk is over: `continue` (kernel.glsl:1690) to the top of the next loop iteration — bounce limit, then russian
// roulette (kernel.glsl:1286-1302). Returns true when the path joins the next bounce's queue.
__device__ bool walk_finish(const FrameConstants& fc, PathState& st, Walker& w) {
    float4 t4 = st.thr[w.pi], c4 = st.col[w.pi];
    const bool bPrevSpec = (w.sw & 0x100u) != 0;
    int bounce = (int)(w.sw & 0xffu) + 1;
    bool terminated = bounce >= fc.settings.MaxBounces;
    if (!terminated && bounce >= 2) {
        float p = fmaxf(fmaxf(w.thr.x, w.thr.y), w.thr.z);
        p = fmaxf(p, EPSILON);
        if (p < w.rng.next()) terminated = true;
        else w.thr *= 1.0f / p;
    }
    if (terminated) { finish_path(fc, st, w.pi, mk3(c4.x, c4.y, c4.z), t4.w, w.rng); return false; }
    st.rayO[w.pi] = make_float4(w.org.x, w.org.y, w.org.z, w.rng.seed);
    st.rayD[w.pi] = make_float4(w.dir.x, w.dir.y, w.dir.z, __uint_as_float(pack_state(bounce, bPrevSpec)));
    st.thr[w.pi] = make_float4(w.thr.x, w.thr.y, w.thr.z, t4.w);
    return true;
}
__device__ __forceinline__ void append_warp(uint32_t* counter, uint32_t* queue, bool pred, uint32_t value) {
    uint32_t ballot = __ballot_sync(0xffffffffu, pred);
    if (ballot) {
        uint32_t lane = threadIdx.x & 31, base = 0;
        if (lane == 0) base = atomicAdd(counter, (uint32_t)__popc(ballot));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (pred) queue[base + __popc(ballot & ((1u << lane) - 1u))] = value;
    }
}

// Walk round `wr`: one step for every walker of walk queue wr, whose rays k_extend<EXT_WALK> has just traced. Walkers
// that go on wait in queue wr ^ 1 for the next round; the others end their bounce. No traversal code in here, so the
// (frequent) one- and two-ray walks of clear glass run at the occupancy and coherence of the regular stages.
__global__ void __launch_bounds__(256) k_walk_step(DeviceScene sc, const FrameConstants* __restrict__ fcp, PathState st, int qi, int wr) {
    const FrameConstants& fc = *fcp;
    const uint32_t count = st.queueCount[10 + wr];
    const uint32_t countUp = (count + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < countUp; i += gridDim.x * blockDim.x) {
        bool walkOn = false, alive = false;
        uint32_t pi = 0;
        if (i < count) {
            pi = st.walkQueue[wr][i];
            Walker w;
            walker_load(st, fc, pi, w);
            float4 h = st.hit[pi];
            walkOn = walk_step(sc, w, h.x, h.y, h.z, st.hitGeom[pi], __float_as_uint(h.w));
            if (walkOn) walker_store(st, w);
            else alive = walk_finish(fc, st, w);
        }
        append_warp(&st.queueCount[10 + (wr ^ 1)], st.walkQueue[wr ^ 1], walkOn, pi);
        append_warp(&st.queueCount[qi ^ 1], st.queue[qi ^ 1], alive, pi);
    }
}

// The long tail: walkers still inside their medium after the wavefront rounds (total internal reflection in closed
// glass bodies, real subsurface scattering: up to 100 rays). Same persistent-warp scheme as k_extend: every lane owns
// a walker, the traversal phase steps all lanes' rays together (one code path per iteration), and the service phase
// runs walk_step for the lanes whose ray has finished, then either starts the walker's next ray or ends its bounce
// and takes a new walker from queue `wr`.
__global__ void __launch_bounds__(128, 4) k_walk(DeviceBvh bvh, DeviceScene sc, const FrameConstants* __restrict__ fcp, PathState st, int qi, int wr, int bounce, uint32_t serviceAt) {
    const FrameConstants& fc = *fcp;
    const uint32_t count = st.queueCount[10 + wr];
    uint32_t* __restrict__ next = &st.queueCount[12 + wr];
    const float4* __restrict__ pairs = (const float4*)bvh.pairs;
    const float4* __restrict__ tris = (const float4*)bvh.tris;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t rays = 0, ntris = 0, nboxes = 0;
    Traversal tr;
    uint32_t stack[TB_STACK_WORDS];
    tr.idle(stack);
    bool have = false, exhausted = false;
    Walker w;
    w.pi = 0; w.sw = 0; w.rng.seed = 0.0f; w.rng.time = fc.time;
    w.org = w.dir = w.thr = w.absorption = mk3(0.0f);
    w.maxTravelDistance = w.roughness = w.travelDistance = 0.0f; w.CurrentIOR = w.NewIOR = 1.0f;
    while (true) {
        // ---- service phase
        bool alive = false;
        if (have && tr.done()) {
            HitRec h;
            tr.result(h);
            rays++; ntris += h.tris; nboxes += h.boxes;
            if (fc.aovMask & AOV_FULL) { uint2 c = st.counters[w.pi]; c.x += h.tris; c.y += h.boxes; st.counters[w.pi] = c; }
            if (walk_step(sc, w, h.t, h.b1, h.b2, h.geom, h.prim)) tr.begin(bvh, stack, w.org, w.dir, MIN_T, FAR_T);
            else { alive = walk_finish(fc, st, w); have = false; }
        }
        append_warp(&st.queueCount[qi ^ 1], st.queue[qi ^ 1], alive, w.pi); // survivors join the next bounce's queue
        uint32_t idle = __ballot_sync(0xffffffffu, !have);
        if (idle && !exhausted) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(next, (uint32_t)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (!have) {
                uint32_t i = base + __popc(idle & ((1u << lane) - 1u));
                if (i < count) {
                    walker_load(st, fc, st.walkQueue[wr][i], w);
                    tr.begin(bvh, stack, w.org, w.dir, MIN_T, FAR_T);
                    have = true;
                }
            }
            if (base + (uint32_t)__popc(idle) >= count) exhausted = true;
        }
        if (!__any_sync(0xffffffffu, have)) break;
        // ---- traversal phase (see k_extend)
        while (true) {
            bool busy = have && !tr.done();
            bool wantLeaf = busy && tr.at_leaf();
            uint32_t mB = __ballot_sync(0xffffffffu, busy), mL = __ballot_sync(0xffffffffu, wantLeaf);
            uint32_t nB = __popc(mB), nL = __popc(mL);
            uint32_t nHave = __popc(__ballot_sync(0xffffffffu, have));
            if (nB == 0 || nHave - nB >= serviceAt) break; // enough finished rays to make a walk step (and a refill) worthwhile
            if (2 * nL > nB) {
                if (wantLeaf) tr.step_leaf(stack, tris);
            } else {
                if (busy && !wantLeaf) tr.step_internal(stack, pairs);
            }
        }
    }
    flush_stats(st, 3, bounce, rays, ntris, nboxes);
}

// Paths whose extension ray left the scene (kernel.glsl:1328-1343): radiance += throughput * environment,
// then the path ends. Split from k_shade so that the (large) miss population neither diverges against
// surface shading nor pays for its register footprint; reads only the ray, throughput and colour.
__global__ void __launch_bounds__(256) k_shade_miss(DeviceScene sc, const FrameConstants* __restrict__ fcp, PathState st, int qi) {
    const FrameConstants& fc = *fcp;
    const uint32_t count = st.queueCount[7 + 2 * qi];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        uint32_t pi = st.missQueue[i];
        float4 o4 = st.rayO[pi], d4 = st.rayD[pi], t4 = st.thr[pi], c4 = st.col[pi];
        Rng rng; rng.seed = o4.w; rng.time = fc.time;
        f3 thr = mk3(t4.x, t4.y, t4.z), acc = mk3(c4.x, c4.y, c4.z);
        if (!(thr.x < EPSILON && thr.y < EPSILON && thr.z < EPSILON)) { // kernel.glsl:1319-1326 comes first
            acc += thr * sample_environment_map(sc, mk3(d4.x, d4.y, d4.z));
            if ((__float_as_uint(d4.w) & 0xffu) == 0u) st.stEmissive[pi] = make_float4(acc.x, acc.y, acc.z, 1.0f);
        }
        finish_path(fc, st, pi, acc, t4.w, rng);
    }
}

// RayTraceCommon tail (RayGenCommon.h:709-727), one launch per frame, in frame order:
// OutputTexture += sample, the jittered-buffer coin (one more rand() of the path's stream),
// and the ordered merge of the two persistent AOVs.
__global__ void __launch_bounds__(256) k_accumulate(FrameConstants fc, PathState st) {
    uint32_t pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= fc.width * fc.height) return;
    if (((pi / fc.width) / 8u) % fc.rowStride != fc.rowOffset) return; // not this shard's row band
    float4 outc = st.sample[pi];
    Rng rng; rng.seed = st.sampleSeed[pi]; rng.time = fc.time;
    bool realtime = fc.settings.RenderMode == TB_RENDER_REALTIME;
    bool clear = realtime || fc.clearAccum;
    float4 prev = clear ? make_float4(0, 0, 0, 0) : st.accum[pi];
    st.accum[pi] = make_float4(outc.x + prev.x, outc.y + prev.y, outc.z + prev.z, outc.w + prev.w);
    if (!realtime) {
        bool take = fc.frame == 0 || rng.next() < 0.5f; // the coin is not drawn on frame 0 (short circuit, :723)
        if (take) {
            float4 pj = fc.clearAccum ? make_float4(0, 0, 0, 0) : st.jittered[pi];
            st.jittered[pi] = make_float4(outc.x + pj.x, outc.y + pj.y, outc.z + pj.z, outc.w + pj.w);
        } else if (fc.clearAccum) st.jittered[pi] = make_float4(0, 0, 0, 0);
    }
    float4 e = st.stEmissive[pi];
    if (e.w != 0.0f) st.aovEmissive[pi] = e;
    float d = st.stDepth[pi];
    if (d >= 0.0f) st.aovDepth[pi] = d;
}

__global__ void k_resolve(const float4* __restrict__ accum, float* __restrict__ rgb, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = accum[i]; // ProcessLit: color / color.w, PostProcessCS.hlsl:23-27
    rgb[3 * (size_t)i] = a.x / a.w; rgb[3 * (size_t)i + 1] = a.y / a.w; rgb[3 * (size_t)i + 2] = a.z / a.w;
}

__global__ void k_trace_rays(DeviceBvh bvh, const TbRay* __restrict__ rays, uint64_t n, TbHit* __restrict__ hits) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        TbRay r = rays[i];
        HitRec h;
        trace_ray(bvh, mk3(r.Origin[0], r.Origin[1], r.Origin[2]), mk3(r.Direction[0], r.Direction[1], r.Direction[2]), r.TMin, r.TMax, h);
        TbHit o;
        o.t = h.t; o.b1 = h.b1; o.b2 = h.b2; o.PrimitiveIndex = h.prim; o.GeometryIndex = h.geom; o.InstanceIndex = 0;
        o.TrianglesTested = h.tris; o.BoxesTested = h.boxes;
        hits[i] = o;
    }
}

// Two-level ray query (TraverseFunction.hlsli:537-785 with FAST_PATH 0), one thread per ray: the
// top-level tree is walked on the reference's own 32-byte nodes with the world-space ray; at an instance leaf whose mask
// passes, the ray goes to object space (origin as a point, direction as a vector, so t keeps its meaning), and the
// instance's bottom level is walked on the traversal layout from its root without a root box test, with the t committed
// so far. One committed t and both counters run across the levels; on exactly equal t the lower (instance, geometry,
// primitive) triple wins (D3 extended).
__global__ void __launch_bounds__(128) k_trace_rays_tlas(const uint8_t* __restrict__ tlas, const TlasInstanceRecord* __restrict__ records,
                                                         const TbRay* __restrict__ rays, uint64_t n, TbHit* __restrict__ hits) {
    const RefNode* __restrict__ tnodes = (const RefNode*)(tlas + 16);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const TbRay r = rays[i];
        const f3 worg = mk3(r.Origin[0], r.Origin[1], r.Origin[2]), wdir = mk3(r.Direction[0], r.Direction[1], r.Direction[2]);
        TbHit o;
        o.t = -1.0f; o.b1 = o.b2 = 0.0f; o.PrimitiveIndex = o.GeometryIndex = 0xffffffffu; o.InstanceIndex = 0; o.TrianglesTested = o.BoxesTested = 0;
        const bool nanRay = worg.x != worg.x || worg.y != worg.y || worg.z != worg.z || wdir.x != wdir.x || wdir.y != wdir.y || wdir.z != wdir.z;
        if (!nanRay) { // D7
            const f3 inv = mk3(1.0f / wdir.x, 1.0f / wdir.y, 1.0f / wdir.z), oinv = worg * inv, ainv = abs3(inv);
            const int zmask = (wdir.x == 0.0f ? 1 : 0) | (wdir.y == 0.0f ? 2 : 0) | (wdir.z == 0.0f ? 4 : 0); // D6
            auto box = [&](const RefNode& b, float closest) {
                return zmask ? slab_zero(closest, worg, zmask, oinv, inv, ainv, b.c[0], b.c[1], b.c[2], b.h[0], b.h[1], b.h[2])
                             : slab(closest, oinv, inv, ainv, b.c[0], b.c[1], b.c[2], b.h[0], b.h[1], b.h[2]);
            };
            float committedT = r.TMax, hb1 = 0.0f, hb2 = 0.0f;
            bool haveHit = false;
            uint32_t hitInst = 0, hitGeom = 0, hitPrim = 0, tris = 0, boxes = 0;
            // A waiting top-level node is kept as its two reference words (flags | left, right): they were read with its box,
            // so a pop goes straight to the children's boxes (or the instance record) without fetching the node again.
            uint2 tstack[TB_TLAS_STACK_DEPTH + 1];
            int tsp = 0;
            { const RefNode root = tnodes[0]; const SlabRange rr = box(root, committedT); if (rr.enter < rr.exit) tstack[tsp++] = make_uint2(root.flags, root.right); }
            uint32_t stack[TB_STACK_WORDS];
            while (tsp > 0) {
                const uint2 nd = tstack[--tsp];
                if (nd.x & 0x80000000u) {
                    const TlasInstanceRecord rec = records[nd.x & 0x3fffffffu];
                    if (rec.mask == 0) continue; // GetInstanceMask & InstanceInclusionMask
                    const float* w = rec.worldToObject;
                    const f3 oorg = mk3(((w[0] * worg.x + w[1] * worg.y) + w[2] * worg.z) + w[3] * 1.0f,
                                        ((w[4] * worg.x + w[5] * worg.y) + w[6] * worg.z) + w[7] * 1.0f,
                                        ((w[8] * worg.x + w[9] * worg.y) + w[10] * worg.z) + w[11] * 1.0f);
                    const f3 odir = mk3(((w[0] * wdir.x + w[1] * wdir.y) + w[2] * wdir.z) + w[3] * 0.0f,
                                        ((w[4] * wdir.x + w[5] * wdir.y) + w[6] * wdir.z) + w[7] * 0.0f,
                                        ((w[8] * wdir.x + w[9] * wdir.y) + w[10] * wdir.z) + w[11] * 0.0f);
                    // what an equal-t hit inside this instance has to beat: the committed triple if it is this instance's,
                    // anything if this instance's index is lower than the committed one, nothing if it is higher
                    uint32_t seedGeom = hitGeom, seedPrim = hitPrim;
                    if (haveHit && rec.instanceIndex < hitInst) seedGeom = seedPrim = 0xffffffffu;
                    else if (haveHit && rec.instanceIndex > hitInst) seedGeom = seedPrim = 0u;
                    DeviceBvh blas;
                    blas.root.c[0] = blas.root.c[1] = blas.root.c[2] = 0.0f; blas.root.h[0] = blas.root.h[1] = blas.root.h[2] = 0.0f; blas.root.flags = 0; blas.root.right = 0;
                    Traversal tr;
                    tr.begin_bottom_level(blas, rec.rootRef, stack, oorg, odir, r.TMin, r.TMax, committedT, haveHit, seedGeom, seedPrim);
                    const float tBefore = committedT;
                    const float4* __restrict__ pairs = (const float4*)rec.pairs;
                    const float4* __restrict__ btris = (const float4*)rec.tris;
                    while (!tr.done()) tr.step(stack, pairs, btris);
                    tris += tr.trisTested; boxes += tr.boxes_tested();
                    if (tr.haveHit && (tr.committedT != tBefore || tr.hitGeom != seedGeom || tr.hitPrim != seedPrim || !haveHit)) {
                        committedT = tr.committedT; hb1 = tr.hb1; hb2 = tr.hb2; hitGeom = tr.hitGeom; hitPrim = tr.hitPrim; hitInst = rec.instanceIndex; haveHit = true;
                    }
                } else {
                    const RefNode L = tnodes[nd.x & 0x3fffffffu], R = tnodes[nd.y];
                    const SlabRange a = box(L, committedT), b = box(R, committedT);
                    boxes += 2;
                    const bool lh = a.enter < a.exit, rh = b.enter < b.exit;
                    const uint2 le = make_uint2(L.flags, L.right), re = make_uint2(R.flags, R.right);
                    if (lh && rh) { const bool rightFirst = b.enter < a.enter; tstack[tsp++] = rightFirst ? le : re; tstack[tsp++] = rightFirst ? re : le; }
                    else if (lh || rh) tstack[tsp++] = rh ? re : le;
                }
            }
            o.TrianglesTested = tris; o.BoxesTested = boxes;
            if (haveHit && committedT < r.TMax) { o.t = committedT; o.b1 = hb1; o.b2 = hb2; o.PrimitiveIndex = hitPrim; o.GeometryIndex = hitGeom; o.InstanceIndex = hitInst; }
        }
        hits[i] = o;
    }
}


// The same query with the warp scheduled like k_extend: every lane owns one ray and is, at any moment, waiting for one of
// three steps - a top-level step (commit a finished instance, pop a waiting top-level node: children boxes or instance
// entry), a bottom-level box-pair step, a bottom-level triangle step. Each iteration the whole warp executes the step
// most lanes wait for, so the three code paths never serialise inside an iteration. Per ray the sequence of steps (and
// with it every result and counter) is exactly k_trace_rays_tlas's.
__global__ void __launch_bounds__(128, 4) k_trace_rays_tlas_warp(const uint8_t* __restrict__ tlas, const TlasInstanceRecord* __restrict__ records,
                                                                 const TbRay* __restrict__ rays, uint64_t n, TbHit* __restrict__ hits) {
    const RefNode* __restrict__ tnodes = (const RefNode*)(tlas + 16);
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5), warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    for (uint64_t base = warp * 32; base < n; base += warps * 32) {
        const uint64_t i = base + lane;
        const bool valid = i < n;
        TbRay r;
        if (valid) r = rays[i];
        else { r.Origin[0] = r.Origin[1] = r.Origin[2] = 0.0f; r.Direction[0] = r.Direction[1] = 0.0f; r.Direction[2] = 1.0f; r.TMin = 0.0f; r.TMax = 0.0f; }
        const f3 worg = mk3(r.Origin[0], r.Origin[1], r.Origin[2]), wdir = mk3(r.Direction[0], r.Direction[1], r.Direction[2]);
        const bool nanRay = worg.x != worg.x || worg.y != worg.y || worg.z != worg.z || wdir.x != wdir.x || wdir.y != wdir.y || wdir.z != wdir.z;
        const f3 inv = mk3(1.0f / wdir.x, 1.0f / wdir.y, 1.0f / wdir.z), oinv = worg * inv, ainv = abs3(inv);
        const int zmask = (wdir.x == 0.0f ? 1 : 0) | (wdir.y == 0.0f ? 2 : 0) | (wdir.z == 0.0f ? 4 : 0); // D6
        auto box = [&](const RefNode& b, float closest) {
            return zmask ? slab_zero(closest, worg, zmask, oinv, inv, ainv, b.c[0], b.c[1], b.c[2], b.h[0], b.h[1], b.h[2])
                         : slab(closest, oinv, inv, ainv, b.c[0], b.c[1], b.c[2], b.h[0], b.h[1], b.h[2]);
        };
        float committedT = r.TMax, hb1 = 0.0f, hb2 = 0.0f;
        bool haveHit = false;
        uint32_t hitInst = 0, hitGeom = 0, hitPrim = 0, tris = 0, boxes = 0;
        uint2 tstack[TB_TLAS_STACK_DEPTH + 1];
        int tsp = 0;
        if (valid && !nanRay) { // D7
            const RefNode root = tnodes[0];
            const SlabRange rr = box(root, committedT);
            if (rr.enter < rr.exit) tstack[tsp++] = make_uint2(root.flags, root.right);
        }
        uint32_t stack[TB_STACK_WORDS];
        Traversal tr;
        tr.idle(stack);
        // the instance this lane is inside of
        bool inBlas = false;
        uint32_t curInst = 0, seedGeom = 0, seedPrim = 0;
        float tBefore = 0.0f;
        const float4* __restrict__ pairs = nullptr;
        const float4* __restrict__ btris = nullptr;
        while (true) {
            const bool busyBlas = inBlas && !tr.done();
            const bool wantLeaf = busyBlas && tr.at_leaf();
            const bool wantInt = busyBlas && !tr.at_leaf();
            const bool wantTop = !busyBlas && (inBlas || tsp > 0);
            const uint32_t nL = __popc(__ballot_sync(0xffffffffu, wantLeaf)), nI = __popc(__ballot_sync(0xffffffffu, wantInt));
            const uint32_t nT = __popc(__ballot_sync(0xffffffffu, wantTop));
            if (nL + nI + nT == 0) break;
            if (nT >= nL && nT >= nI) {
                if (wantTop) {
                    if (inBlas) { // the instance's bottom level is done: commit what it found
                        tris += tr.trisTested; boxes += tr.boxes_tested();
                        if (tr.haveHit && (tr.committedT != tBefore || tr.hitGeom != seedGeom || tr.hitPrim != seedPrim || !haveHit)) {
                            committedT = tr.committedT; hb1 = tr.hb1; hb2 = tr.hb2; hitGeom = tr.hitGeom; hitPrim = tr.hitPrim; hitInst = curInst; haveHit = true;
                        }
                        inBlas = false;
                        tr.cur = TB_NO_NODE;
                    }
                    if (tsp > 0) {
                        const uint2 nd = tstack[--tsp];
                        if (nd.x & 0x80000000u) {
                            const TlasInstanceRecord rec = records[nd.x & 0x3fffffffu];
                            if (rec.mask != 0) { // GetInstanceMask & InstanceInclusionMask
                                const float* w = rec.worldToObject;
                                const f3 oorg = mk3(((w[0] * worg.x + w[1] * worg.y) + w[2] * worg.z) + w[3] * 1.0f,
                                                    ((w[4] * worg.x + w[5] * worg.y) + w[6] * worg.z) + w[7] * 1.0f,
                                                    ((w[8] * worg.x + w[9] * worg.y) + w[10] * worg.z) + w[11] * 1.0f);
                                const f3 odir = mk3(((w[0] * wdir.x + w[1] * wdir.y) + w[2] * wdir.z) + w[3] * 0.0f,
                                                    ((w[4] * wdir.x + w[5] * wdir.y) + w[6] * wdir.z) + w[7] * 0.0f,
                                                    ((w[8] * wdir.x + w[9] * wdir.y) + w[10] * wdir.z) + w[11] * 0.0f);
                                seedGeom = hitGeom; seedPrim = hitPrim; // see k_trace_rays_tlas
                                if (haveHit && rec.instanceIndex < hitInst) seedGeom = seedPrim = 0xffffffffu;
                                else if (haveHit && rec.instanceIndex > hitInst) seedGeom = seedPrim = 0u;
                                DeviceBvh blas;
                                blas.root.c[0] = blas.root.c[1] = blas.root.c[2] = 0.0f; blas.root.h[0] = blas.root.h[1] = blas.root.h[2] = 0.0f; blas.root.flags = 0; blas.root.right = 0;
                                tr.begin_bottom_level(blas, rec.rootRef, stack, oorg, odir, r.TMin, r.TMax, committedT, haveHit, seedGeom, seedPrim);
                                tBefore = committedT;
                                curInst = rec.instanceIndex;
                                pairs = (const float4*)rec.pairs;
                                btris = (const float4*)rec.tris;
                                inBlas = true;
                            }
                        } else {
                            const RefNode L = tnodes[nd.x & 0x3fffffffu], R = tnodes[nd.y];
                            const SlabRange a = box(L, committedT), b = box(R, committedT);
                            boxes += 2;
                            const bool lh = a.enter < a.exit, rh = b.enter < b.exit;
                            const uint2 le = make_uint2(L.flags, L.right), re = make_uint2(R.flags, R.right);
                            if (lh && rh) { const bool rightFirst = b.enter < a.enter; tstack[tsp++] = rightFirst ? le : re; tstack[tsp++] = rightFirst ? re : le; }
                            else if (lh || rh) tstack[tsp++] = rh ? re : le;
                        }
                    }
                }
            } else if (nL >= nI) {
                if (wantLeaf) tr.step_leaf(stack, btris);
            } else {
                if (wantInt) tr.step_internal(stack, pairs);
            }
        }
        if (valid) {
            TbHit o;
            o.t = -1.0f; o.b1 = o.b2 = 0.0f; o.PrimitiveIndex = o.GeometryIndex = 0xffffffffu; o.InstanceIndex = 0;
            o.TrianglesTested = tris; o.BoxesTested = boxes;
            if (haveHit && committedT < r.TMax) { o.t = committedT; o.b1 = hb1; o.b2 = hb2; o.PrimitiveIndex = hitPrim; o.GeometryIndex = hitGeom; o.InstanceIndex = hitInst; }
            hits[i] = o;
        }
    }
}

} // namespace

// ------------------------------------------------------------------ launchers

// Per-frame constants live in device memory (one copy per frame slot) so that the frame's kernel sequence can be
// captured once into a CUDA graph and replayed: nothing baked into the graph changes from frame to frame.
__global__ void k_set_frame(FrameConstants fc, FrameConstants* dst) { *dst = fc; }

// Scheduling knobs from the environment (results never depend on them). Read once per process, by
// whichever thread renders first; -1 = not set, the automatic policy applies.
struct Tuning {
    int suspend = -1, sort = -1, classSort = -1, walkRounds = -1, walkService = -1, graphs = 1;
    uint32_t budgetMain = EXTEND_BUDGET_MAIN, budgets[EXTEND_RESUME_ROUNDS] = {384u, 1536u, 0u}, refillBelow = REFILL_THRESHOLD;
    Tuning() {
        auto num = [](const char* name, int unset) { const char* e = getenv(name); return e ? atoi(e) : unset; };
        suspend = num("TB_SUSPEND", -1); sort = num("TB_SORT", -1); classSort = num("TB_CLASS_SORT", -1);
        walkRounds = num("TB_WALK_ROUNDS", -1); walkService = num("TB_WALK_SERVICE", -1); graphs = num("TB_GRAPHS", 1);
        if (const char* bs = getenv("TB_BUDGETS")) sscanf(bs, "%u,%u,%u", &budgetMain, &budgets[0], &budgets[1]);
        if (const char* rs = getenv("TB_REFILL")) refillBelow = (uint32_t)atoi(rs);
    }
};
const Tuning& tuning() {
    static const Tuning t;
    return t;
}

// Sorting the bounce queues pays when node fetches go to DRAM: traversal layout (112 B per triangle) well beyond the
// 126 MB L2. Measured (bounce + shadow queues): 20.8 M triangles (2.3 GB) +10.6 %; 875 k triangles (98 MB) -2 %,
// Teapot (14 MB) -5 %: when the nodes already come from L1 / L2 the three extra launches per queue are pure cost.
static bool sort_pays(const DeviceBvh& bvh) { return (uint64_t)bvh.numPrims * 112ull > (256ull << 20); }

static cudaError_t launch_frame(const DeviceBvh& bvh, const DeviceScene& sc, const FrameConstants& fc, const FrameConstants* fcDev,
                                PathState& st, cudaStream_t stream, uint64_t& launches, KernelTimers* timers, const RenderOptions& opts) {
    const uint32_t n = fc.width * fc.height;
    cudaMemsetAsync(st.queueCount, 0, 4 * TB_QUEUE_COUNT_WORDS, stream);
    cudaMemsetAsync(st.susCount, 0, 16, stream);
    k_raygen<<<(n + 255) / 256, 256, 0, stream>>>(sc, fcDev, st); launches++;
    // persistent grids: a multiple of the SM count, capped by the work available
    const uint32_t sms = (uint32_t)(opts.numSMs > 0 ? opts.numSMs : 148);
    const uint32_t maxBlocks = sms * 16;
    uint32_t blocks = (n + 127) / 128;
    if (blocks > maxBlocks) blocks = maxBlocks;
    const int maxBounces = fc.settings.MaxBounces;
    const uint32_t heat = fc.settings.OutputType == TB_OUTPUT_HEATMAP;
    for (int b = 0; b < maxBounces; b++) {
        int qi = b & 1;
        if (timers) cudaEventRecord(timers->next(KernelTimers::EXTEND, b), stream);
        const Tuning& tune = tuning();
        const uint32_t budgetMain = tune.budgetMain, refillBelow = tune.refillBelow;
        const uint32_t* budgets = tune.budgets;
        // ray suspension pays when few frames are in flight (their tails have nothing to overlap with): measured with 16
        // slots it costs 1-2 % (three mostly empty launches per bounce), with one slot it is worth 2.2x on Teapot
        const int suspendMode = tune.suspend >= 0 ? tune.suspend : (opts.suspendRays ? 1 : 0);
        // bounce queues of large scenes are sorted by origin cell first (primary rays already come in 8x4 pixel tiles)
        // bit 0: the bounce queue, bit 1: the shadow queue
        const int sortMode = tune.sort >= 0 ? tune.sort : (opts.sortRays == 2 ? (sort_pays(bvh) ? 3 : 0) : opts.sortRays);
        PathState stx = st; // what k_extend<EXT_MAIN> reads its queue from
        auto sort_queue = [&](const uint32_t* queue, const uint32_t* countPtr, const float4* rayO) {
            cudaMemsetAsync(st.sortHist, 0, 4 * (TB_SORT_CELLS + 1), stream);
            k_sort_keys<<<blocks / 2 ? blocks / 2 : 1, 256, 0, stream>>>(bvh, st, queue, countPtr, rayO); launches++;
            k_sort_scan<<<1, 1024, 0, stream>>>(st); launches++;
            k_sort_scatter<<<blocks / 2 ? blocks / 2 : 1, 256, 0, stream>>>(st, queue, countPtr); launches++;
        };
        if (sortMode && b > 0) {
            sort_queue(st.queue[qi], &st.queueCount[qi], st.rayO);
            stx.queue[qi] = st.sortTmp;
        }
        // hits are binned by material class when the scene's reachable materials span more than one class
        // automatic: from four classes on. Measured on B200 (profiles/r2_class_sort_sweep.log): vw-van (glass, metal, mix,
        // textured, matte) +3 %, 20.8 M triangles +2 %, Teapot +0.4 %; scenes with two classes lose 3-4 % to the extra
        // launch per bounce (dragon, cornell)
        const bool classSort = tune.classSort >= 0 ? tune.classSort != 0 : (opts.materialSort == 2 ? opts.sceneMaterialClasses >= 4 : opts.materialSort != 0);
        const uint8_t* classOfGeom = classSort ? sc.geomClass : nullptr;
        k_extend<EXT_MAIN><<<blocks, 128, 0, stream>>>(bvh, stx, qi, b, heat, fcDev, suspendMode ? budgetMain : 0xffffffffu, refillBelow, classOfGeom); launches++;
        if (suspendMode) {
            uint32_t rblocks = (st.susCapacity + 127) / 128;
            if (rblocks > sms * 4) rblocks = sms * 4;
            if (timers) cudaEventRecord(timers->next(KernelTimers::RESUME, b), stream);
            for (int r = 1; r <= EXTEND_RESUME_ROUNDS; r++) {
                k_extend_resume<<<rblocks, 128, 0, stream>>>(bvh, st, qi, r, budgets[r - 1], b, heat, fcDev, classOfGeom); launches++;
                rblocks = (rblocks + 3) / 4 > sms ? (rblocks + 3) / 4 : sms;
            }
        }
        if (timers) cudaEventRecord(timers->next(KernelTimers::SHADE, b), stream);
        k_shade_miss<<<blocks / 2 ? blocks / 2 : 1, 256, 0, stream>>>(sc, fcDev, st, qi); launches++;
        if (classSort) { k_class_scatter<<<(n + 2047) / 2048 < sms * 8 ? (n + 2047) / 2048 : sms * 8, 256, 0, stream>>>(st, qi); launches++; }
        // next-event shadow rays exist only with lights and NEE on; then shading runs as two stages
        // around a traversal kernel for the shadow queue
        const bool nee = sc.numLights > 0 && fc.settings.EnableNextEventEstimation;
        // 0 = inline, 1 = queue, 2 = automatic: the two-stage form re-reads the path state and redoes the
        // part of the bounce before the shadow ray, which only pays off when traversal is expensive
        // (measured: 874 k triangles +9 %, 36 triangles -26 %)
        const int shadowMode = opts.shadowMode == 2 ? (bvh.numPrims >= 32768u ? 1 : 0) : opts.shadowMode;
        const bool sss = opts.sceneHasSSS;
#define TB_LAUNCH_SHADE(STG) do { if (sss) k_shade<STG, true><<<blocks, 128, 0, stream>>>(bvh, sc, fcDev, st, qi, b, classSort ? 1 : 0); \
                                  else k_shade<STG, false><<<blocks, 128, 0, stream>>>(bvh, sc, fcDev, st, qi, b, classSort ? 1 : 0); launches++; } while (0)
        if (nee && shadowMode) {
            TB_LAUNCH_SHADE(0);
            PathState sts = st; // what k_extend<EXT_SHADOW> reads its queue from
            if (sortMode & 2) { // the bounce's shadow feelers, by origin cell (sortTmp is free again: k_extend<EXT_MAIN> is done)
                sort_queue(st.shadowQueue, &st.queueCount[4], st.shRayO);
                sts.shadowQueue = st.sortTmp;
            }
            k_extend<EXT_SHADOW><<<blocks, 128, 0, stream>>>(bvh, sts, qi, b, 0, fcDev, 0xffffffffu, refillBelow, nullptr); launches++;
            TB_LAUNCH_SHADE(1);
        } else if (nee) {
            TB_LAUNCH_SHADE(2);
        } else {
            TB_LAUNCH_SHADE(3);
        }
#undef TB_LAUNCH_SHADE
        if (sss) { // the bounce's glass / subsurface walkers (queued by the k_shade launches above into walk queue 0)
            const int rounds = tune.walkRounds >= 0 ? tune.walkRounds : opts.walkRounds; // wavefront rounds before the persistent tail (which reads queue rounds & 1)
            for (int r = 0; r < rounds; r++) {
                k_extend<EXT_WALK><<<blocks, 128, 0, stream>>>(bvh, st, r & 1, b, 0, fcDev, 0xffffffffu, REFILL_THRESHOLD, nullptr); launches++;
                k_walk_step<<<blocks / 2 ? blocks / 2 : 1, 256, 0, stream>>>(sc, fcDev, st, qi, r & 1); launches++;
            }
            uint32_t wblocks = blocks > sms * 4 ? sms * 4 : blocks;
            // finished rays that end a traversal phase: a walk step is long divergent code, so on a cache-resident tree
            // (cheap traversal) it pays to batch more of them; on an HBM-resident tree idle lanes cost more than that.
            // Measured (profiles/r2_walk_service_sweep.log): vw-van 8 -> 24 +3.8 %, 20.8 M triangles 8 -> 16 -5 %
            const uint32_t serviceAt = tune.walkService > 0 ? (uint32_t)tune.walkService : (bvh.numPrims < (1u << 22) ? 24u : 8u);
            k_walk<<<wblocks, 128, 0, stream>>>(bvh, sc, fcDev, st, qi, rounds & 1, b, serviceAt); launches++;
        }
        if (timers) cudaEventRecord(timers->next(KernelTimers::END, b), stream);
    }
    return cudaGetLastError();
}

void FrameGraph::reset() {
    if (exec) cudaGraphExecDestroy(exec);
    exec = nullptr;
    launches = 0;
}

// One frame (one sample per pixel) on `stream`. With a FrameGraph the kernel sequence is captured the first time and
// replayed afterwards: one k_set_frame launch + one graph launch per frame instead of ~35 launches, which is what a
// small frame (cornell 512^2: 8 us of GPU work per kernel) is bound by. The key lists everything that is baked into
// the captured launches; `epoch` is bumped by the host whenever a device pointer baked into them may have changed.
cudaError_t render_frame(const DeviceBvh& bvh, const DeviceScene& sc, const FrameConstants& fc, FrameConstants* fcDev, PathState& st,
                         cudaStream_t stream, LaunchCounter& lc, KernelTimers* timers, const RenderOptions& opts, FrameGraph* graph) {
    k_set_frame<<<1, 1, 0, stream>>>(fc, fcDev); lc.count++;
    if (!graph || timers || !tuning().graphs) return launch_frame(bvh, sc, fc, fcDev, st, stream, lc.count, timers, opts);
    FrameGraph::Key key = {opts.epoch, fc.width, fc.height, (uint32_t)fc.settings.MaxBounces, fc.settings.OutputType == TB_OUTPUT_HEATMAP ? 1u : 0u,
                           fc.settings.EnableNextEventEstimation ? 1u : 0u, (uint32_t)opts.shadowMode, (uint32_t)opts.walkRounds,
                           opts.sceneHasSSS ? 1u : 0u, opts.suspendRays ? 1u : 0u, (uint32_t)opts.sortRays,
                           (uint32_t)opts.materialSort * 256u + opts.sceneMaterialClasses};
    if (!graph->exec || memcmp(&key, &graph->key, sizeof(key)) != 0) {
        graph->reset();
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) return e;
        uint64_t launches = 0;
        cudaError_t le = launch_frame(bvh, sc, fc, fcDev, st, stream, launches, nullptr, opts);
        e = cudaStreamEndCapture(stream, &g);
        if (le != cudaSuccess) { if (g) cudaGraphDestroy(g); return le; }
        if (e != cudaSuccess) return e;
        e = cudaGraphInstantiate(&graph->exec, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { graph->exec = nullptr; return e; }
        graph->key = key;
        graph->launches = launches;
    }
    lc.count += graph->launches;
    return cudaGraphLaunch(graph->exec, stream);
}

cudaError_t accumulate_frame(const FrameConstants& fc, PathState& st, cudaStream_t stream, LaunchCounter& lc) {
    const uint32_t n = fc.width * fc.height;
    k_accumulate<<<(n + 255) / 256, 256, 0, stream>>>(fc, st); lc.count++;
    return cudaGetLastError();
}

cudaError_t resolve_rgb(const float4* accum, float* rgb, uint32_t n, cudaStream_t stream, LaunchCounter& lc) {
    k_resolve<<<(n + 255) / 256, 256, 0, stream>>>(accum, rgb, n); lc.count++;
    return cudaGetLastError();
}

cudaError_t trace_rays(const DeviceBvh& bvh, const TbRay* d_rays, uint64_t n, TbHit* d_hits, int numSMs, cudaStream_t stream, LaunchCounter& lc) {
    uint64_t blocks = (n + 127) / 128;
    uint64_t cap = (uint64_t)(numSMs > 0 ? numSMs : 148) * 16;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) return cudaSuccess;
    k_trace_rays<<<(uint32_t)blocks, 128, 0, stream>>>(bvh, d_rays, n, d_hits); lc.count++;
    return cudaGetLastError();
}

cudaError_t trace_rays_tlas(const uint8_t* tlasRef, const TlasInstanceRecord* records, uint32_t numInstances, const TbRay* d_rays, uint64_t n,
                            TbHit* d_hits, int numSMs, cudaStream_t stream, LaunchCounter& lc) {
    (void)numInstances;
    uint64_t blocks = (n + 127) / 128, cap = (uint64_t)(numSMs > 0 ? numSMs : 148) * 16;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) return cudaSuccess;
    // TB_TLAS_QUERY=thread selects the one-thread-per-ray form (kept as the plain statement of the loop; same results)
    static const bool perThread = [] { const char* e = getenv("TB_TLAS_QUERY"); return e && strcmp(e, "thread") == 0; }();
    if (perThread) k_trace_rays_tlas<<<(uint32_t)blocks, 128, 0, stream>>>(tlasRef, records, d_rays, n, d_hits);
    else k_trace_rays_tlas_warp<<<(uint32_t)blocks, 128, 0, stream>>>(tlasRef, records, d_rays, n, d_hits);
    lc.count++;
    return cudaGetLastError();
}

} // namespace tbd
