// core.cpp — CPU ORACLE (test infrastructure): restatement of the reference's per-pixel
// path-tracing core as host C++, OpenMP-parallel over pixels.
//
// Follows, in this order (file:line in /root/reference/TracerBoy unless noted):
//   SoftwareRayTraceCS.hlsl:36-50        seed + RayTraceCommon
//   RayGenCommon.h:690-728               RayTraceCommon (accumulate, jittered buffer, AOV world pos)
//   kernel.glsl:1805-1921                PathTrace (camera, filter, DOF, firefly clamp)
//   kernel.glsl:1278-1776                Trace (bounce loop)
//   RayGenCommon.h:21-44, 49-135, 137-261, 273-341, 364-414, 524-654, 662-667   glue
//   SharedHitGroup.h:38-151, SharedRaytracing.h:55-137, Tonemap.h:12-15, 208-211
//   kernel.glsl:39-40, 152-199, 308-311, 441-556, 991-1099, 1186-1269          helpers
// HLSL semantics are pinned by tracerboy_b200/csrc/common/tb_math.h and tb_vec.h; all
// rand() draws inside one argument list are hoisted left-to-right (SURVEY §8c trap 17).
#include <cmath>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../tracerboy_b200/csrc/common/tb_vec.h"
#include "glue.h"
#include "oracle.h"

using namespace tbm;

namespace oracle {

void FrameBuffers::resize(uint32_t w, uint32_t h) {
    width = w; height = h;
    size_t n = (size_t)w * h;
    TbFloat4 z = {0, 0, 0, 0};
    accum.assign(n, z); jittered.assign(n, z); aovNormal.assign(n, z); aovWorldPos[0].assign(n, z);
    aovWorldPos[1].assign(n, z); aovAlbedo.assign(n, z); aovEmissive.assign(n, z);
    aovDepth.assign(n, 0.0f);
    primaryHit.assign(2 * n, 0xffffffffu);
    counters.assign(2 * n, 0);
    raysTraced = boxesTested = trianglesTested = 0;
    memset(&stats, 0, sizeof(stats));
}






// ---------------------------------------------------------------- textures
// The sampler's sRGB -> linear conversion of an _SRGB format (D3D11.3 functional spec 7.2.1.2), pinned intrinsics.
inline float srgb_to_linear(float c) { return c <= 0.04045f ? c / 12.92f : pow_((c + 0.055f) / 1.055f, 2.4f); }
f4 fetch_texel(const Image& im, int x, int y) {
    if (im.format == 0) {
        const float* p = (const float*)im.data.data() + 4 * ((size_t)y * im.width + x);
        return mk4(p[0], p[1], p[2], p[3]);
    }
    const uint8_t* p = im.data.data() + 4 * ((size_t)y * im.width + x);
    f4 c = mk4((float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f, (float)p[3] / 255.0f);
    if (im.format == 2) { c.x = srgb_to_linear(c.x); c.y = srgb_to_linear(c.y); c.z = srgb_to_linear(c.z); } // R8G8B8A8_UNORM_SRGB: per texel, before filtering
    return c;
}
inline int wrapi(int i, int n) { i %= n; return i < 0 ? i + n : i; }
inline f4 lerp4(f4 a, f4 b, float s) { return mk4(lerp(a.x, b.x, s), lerp(a.y, b.y, s), lerp(a.z, b.z, s), lerp(a.w, b.w, s)); }
// SampleLevel(BilinearSampler, uv, 0): bilinear, wrap, mip 0; pinned with exact float weights.
f4 sample_bilinear_wrap(const Image& im, f2 uv) {
    const float nanv = as_float(0x7fc00000u);
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

// SharedRaytracing.h:80-137
f4 texture_nonrecursive(const Scene& sc, const TbTextureData& t, f2 uv) {
    f4 data = mk4(0, 0, 0, 0);
    switch (t.TextureType) {
    case TB_IMAGE_TEXTURE_TYPE:
        data = sample_bilinear_wrap(sc.images[t.DescriptorHeapIndex], uv);
        break;
    case TB_CHECKER_TEXTURE_TYPE: {
        f2 scaled = uv * mk2(t.UScale, t.VScale);
        data = mk4(F3(t.CheckerColor1), 1.0f);
        if ((((int)scaled.x + (int)scaled.y) % 2) == 0) data = mk4(F3(t.CheckerColor2), 1.0f);
        break;
    }
    default: break;
    }
    if (t.TextureFlags & TB_NEEDS_GAMMA_CORRECTION_TEXTURE_FLAG) { // Tonemap.h:208-211
        data.x = pow_(data.x, 2.2f); data.y = pow_(data.y, 2.2f); data.z = pow_(data.z, 2.2f);
    }
    return data;
}
f4 get_texture_data(const Scene& sc, uint32_t index, f2 uv) { // SharedRaytracing.h:67-78, 117-137
    if (index == TB_INVALID_TEXTURE) return mk4(0, 0, 0, 0);
    if (sc.flipTextureUVs) uv = mk2(0.0f, 1.0f) + uv * mk2(1.0f, -1.0f);
    const TbTextureData& t = sc.textures[index];
    if (t.TextureType == TB_SCALE_TEXTURE_TYPE) {
        f4 c1 = texture_nonrecursive(sc, sc.textures[t.TextureIndex1], uv);
        f4 c2 = texture_nonrecursive(sc, sc.textures[t.TextureIndex2], uv);
        return c1 * mk4(F3(t.ScaleColor1), 1.0f) + c2 * mk4(F3(t.ScaleColor2), 1.0f);
    }
    return texture_nonrecursive(sc, t, uv);
}

// RayGenCommon.h:21-44
f3 sample_environment_map(const Scene& sc, f3 v) {
    if (sc.envImage < 0) return mk3(0.0f); // 1x1 black texture (TracerBoy.cpp:1918-1934)
    v = mk3(dot(v, mk3(sc.envTransform[0].x, sc.envTransform[0].y, sc.envTransform[0].z)),
            dot(v, mk3(sc.envTransform[1].x, sc.envTransform[1].y, sc.envTransform[1].z)),
            dot(v, mk3(sc.envTransform[2].x, sc.envTransform[2].y, sc.envTransform[2].z)));
    f3 viewDir = normalize(v);
    float p = atan2_(viewDir.y, viewDir.x);
    p = p > 0.0f ? p : p + 2.0f * 3.14f;
    f2 uv;
    uv.x = p / (2.0f * 3.14f);
    uv.y = acos_(viewDir.z) / 3.14f;
    f4 t = sample_bilinear_wrap(sc.images[sc.envImage], uv);
    return mk3(t.x, t.y, t.z) * F3(sc.envColorScale);
}

// ----------------------------------------------------------- RNG / noise
float hash13(f3 p3) { // RayGenCommon.h:662-667
    p3 = frac3(p3 * 0.1031f);
    float d = dot(p3, mk3(p3.y, p3.z, p3.x) + 33.33f);
    p3 = p3 + d;
    return frac((p3.x + p3.y) * p3.z);
}
float halton(int b, int i) { // RayGenCommon.h:49-60
    float r = 0.0f, f = 1.0f;
    while (i > 0) {
        f = f / (float)b;
        r = r + f * (float)(i % b);
        i = (int)floorf((float)i / (float)b);
    }
    return r;
}
BlueNoiseData get_blue_noise(Ctx& c) { // RayGenCommon.h:104-122
    BlueNoiseData d;
    if (!c.rp.settings.EnableBlueNoise) {
        float a, b;
        a = c.rand(); b = c.rand(); d.PrimaryJitter = mk2(a, b);
        a = c.rand(); b = c.rand(); d.SecondaryRayDirection = mk2(a, b);
        a = c.rand(); b = c.rand(); d.AreaLightJitter = mk2(a, b);
        a = c.rand(); b = c.rand(); d.DOFJitter = mk2(a, b);
    } else {
        f2 h = mk2(halton(2, (int)c.rp.frame), halton(3, (int)c.rp.frame));
        const uint8_t* t0 = c.sc.blueNoise.data() + 4 * ((size_t)(c.py % 256) * 256 + (c.px % 256));
        const uint8_t* t1 = t0 + 256 * 256 * 4;
        auto un = [](uint8_t b) { return (float)b / 255.0f; };
        d.PrimaryJitter = frac2(mk2(un(t0[0]), un(t0[1])) + h);
        d.SecondaryRayDirection = frac2(mk2(un(t0[2]), un(t0[3])) + h);
        d.AreaLightJitter = frac2(mk2(un(t1[0]), un(t1[1])) + h);
        d.DOFJitter = frac2(mk2(un(t1[2]), un(t1[3])) + h);
    }
    return d;
}

// ------------------------------------------------------------- intersect

// IntersectWithMaxDistance (SW branch), RayGenCommon.h:365-414 + SharedHitGroup.h:48-151
HitResult intersect(Ctx& c, const Ray& ray, float maxT) {
    TbRay r;
    r.Origin[0] = ray.origin.x; r.Origin[1] = ray.origin.y; r.Origin[2] = ray.origin.z; r.TMin = MIN_T;
    r.Direction[0] = ray.direction.x; r.Direction[1] = ray.direction.y; r.Direction[2] = ray.direction.z; r.TMax = maxT;
    TbHit h;
    trace_ray(c.sc, r, h);
    c.rays++; c.tris += h.TrianglesTested; c.boxes += h.BoxesTested;
    if (c.firstIntersect) { c.firstIntersect = false; c.primGeom = h.GeometryIndex; c.primPrim = h.PrimitiveIndex; }
    if (c.rp.settings.OutputType == TB_OUTPUT_HEATMAP) // OutputRayStats, RayGenCommon.h:537-543
        c.aovAlbedo = mk4((float)h.TrianglesTested, (float)h.BoxesTested, 0.0f, 0.0f);
    HitResult res;
    if (h.t >= 0.0f) {
        const TbGeometryRecord& G = c.sc.geoms[h.GeometryIndex];
        f3 bary = mk3(1.0f - h.b1 - h.b2, h.b1, h.b2);
        const uint32_t* idx = &c.sc.indices[G.IndexFirst + 3 * (size_t)h.PrimitiveIndex];
        const TbVertex& a = c.sc.vertices[G.VertexFirst + idx[0]];
        const TbVertex& b = c.sc.vertices[G.VertexFirst + idx[1]];
        const TbVertex& d = c.sc.vertices[G.VertexFirst + idx[2]];
        f2 uv0 = mk2(a.UV.x, a.UV.y), uv1 = mk2(b.UV.x, b.UV.y), uv2 = mk2(d.UV.x, d.UV.y);
        res.uv = (bary.x * uv0 + bary.y * uv1) + bary.z * uv2;
        res.normal = normalize((bary.x * F3(a.Normal) + bary.y * F3(b.Normal)) + bary.z * F3(d.Normal));
        res.tangent = normalize((bary.x * F3(a.Tangent) + bary.y * F3(b.Tangent)) + bary.z * F3(d.Tangent));
        res.t = h.t;
        res.material = (int)G.MaterialIndex;
    } else {
        res.t = -1.0f; res.material = -1; res.normal = mk3(0.0f); res.tangent = mk3(0.0f); res.uv = mk2(0, 0);
    }
    return res;
}

// ---------------------------------------------------------------- materials
inline bool AllowsSpecular(const Material& m) { return (m.Flags & TB_NO_SPECULAR_MATERIAL_FLAG) == 0; }
inline bool IsMetallic(const Material& m) { return (m.Flags & TB_METALLIC_MATERIAL_FLAG) != 0; }
inline bool IsSubsurfaceScattering(const Material& m) { return (m.Flags & TB_SUBSURFACE_SCATTER_MATERIAL_FLAG) != 0; }
inline bool IsHairMaterial(const Material& m) { return (m.Flags & TB_HAIR_MATERIAL_FLAG) != 0; }
inline bool IsLight(const Material& m) { return (m.Flags & TB_LIGHT_MATERIAL_FLAG) != 0; }
inline bool UsePerfectSpecularOptimization(float roughness) { return roughness < 0.05f; }

// GetMaterialInternal, RayGenCommon.h:298-341
Material get_material_internal(Ctx& c, int id, f2 uv, bool backside) {
    Material mat = load_material(c.sc.materials[id]);
    bool ignoreEmissive = backside;
    if (ignoreEmissive) mat.emissive = mk3(0.0f);
    if ((mat.Flags & TB_MIX_MATERIAL_FLAG) != 0) {
        if (c.rand() < mat.albedo.z) return load_material(c.sc.materials[(uint32_t)mat.albedo.x]);
        else return load_material(c.sc.materials[(uint32_t)mat.albedo.y]);
    }
    if (mat.albedoIndex != TB_INVALID_TEXTURE) { f4 t = get_texture_data(c.sc, mat.albedoIndex, uv); mat.albedo = mk3(t.x, t.y, t.z); }
    if (mat.emissiveIndex != TB_INVALID_TEXTURE && !ignoreEmissive) { f4 t = get_texture_data(c.sc, mat.emissiveIndex, uv); mat.emissive = mk3(t.x, t.y, t.z); }
    if (mat.specularMapIndex != TB_INVALID_TEXTURE) {
        f4 t = get_texture_data(c.sc, mat.specularMapIndex, uv);
        mat.roughness = t.y;
        if (t.z > 0.5f) mat.Flags |= TB_METALLIC_MATERIAL_FLAG;
    }
    return mat;
}
// kernel.glsl:1224-1246
Material get_material(Ctx& c, int id, f2 uv, bool backside) {
    Material m = get_material_internal(c, id, uv, backside);
    if (IsSubsurfaceScattering(m) && (m.albedo.x != 0.0f || m.albedo.y != 0.0f || m.albedo.z != 0.0f)) {
        f3 color = m.albedo;
        f3 mfp = mk3(1.0f) / m.scattering;
        f3 alpha = mk3(1.0f) - exp3(((-5.09406f * color) + ((2.61188f * color) * color)) - (((4.31805f * color) * color) * color));
        f3 s = (mk3(1.9f) - color) + ((3.5f * (color - 0.8f)) * (color - 0.8f));
        f3 transmission = mk3(1.0f) / (s * mfp);
        m.scattering = transmission * alpha;
        m.absorption = transmission - m.scattering;
        m.albedo = mk3(0.0f);
    }
    return m;
}
// GetDetailNormal, RayGenCommon.h:273-295
f3 get_detail_normal(Ctx& c, const Material& mat, f3 normal, f3 tangent, f2 uv) {
    if (mat.normalMapIndex != TB_INVALID_TEXTURE && c.rp.settings.EnableNormalMaps) {
        f3 bitangent = cross(tangent, normal);
        f4 nm = get_texture_data(c.sc, mat.normalMapIndex, uv);
        f3 tbn = mk3((0.5f - nm.x) * 2.0f, (0.5f - nm.y) * 2.0f, 0.0f);
        tbn.z = sqrtf(1.0f - (tbn.x * tbn.x + tbn.y * tbn.y));
        const float normalYClamp = 0.02f;
        return normalize((tangent * tbn.x + bitangent * tbn.y) + normal * fmaxf(tbn.z, normalYClamp));
    }
    return normal;
}

// ------------------------------------------------------------------ lights
f3 random_barycentric(Ctx& c) { // RayGenCommon.h:124-135
    float u = c.rand();
    float v = c.rand();
    if (u + v > 1.0f) { u = 1.0f - u; v = 1.0f - v; }
    return mk3(u, v, 1.0f - u - v);
}
float color_to_luma(f3 col) { return dot(col, mk3(0.212671f, 0.715160f, 0.072169f)); }
float light_target_pdf(const TbLight& l, f3 bary, f3 pos) { // RayGenCommon.h:163-168 (precedence slip kept)
    f3 lp = (F3(l.P0) * bary.x + F3(l.P1) * bary.y) + F3(l.P2) * bary.z;
    float d = length(lp - pos);
    return (l.SurfaceArea * color_to_luma(F3(l.LightColor))) / d * d;
}
// GetOneLightSample, RayGenCommon.h:170-261
void get_one_light_sample(Ctx& c, f3 pos, f3& LightDirection, f3& LightColor, float& PDFValue, f3& LightNormal, float& LightAttenuation) {
    LightDirection = LightColor = LightNormal = mk3(0.0f);
    LightAttenuation = 0.0f;
    PDFValue = 0.0f;
    const uint32_t lightCount = (uint32_t)c.sc.lights.size();
    if (lightCount > 0 && c.rp.settings.EnableNextEventEstimation) {
        if (c.rp.settings.EnableSamplingImportanceResampling) {
            uint32_t selIndex = 0; f3 selBary = mk3(0.0f); float weightSum = 0.0f;
            const uint32_t N = 16;
            for (uint32_t i = 0; i < N; i++) {
                uint32_t li = (uint32_t)(c.rand() * (float)lightCount);
                const TbLight& light = c.sc.lights[li];
                f3 bary = random_barycentric(c);
                float target = light_target_pdf(light, bary, pos);
                float proposal = 1.0f / (float)lightCount;
                float w = target / (proposal * (float)N);
                weightSum += w;
                if (c.rand() < w / weightSum) { selIndex = li; selBary = bary; }
            }
            const TbLight& light = c.sc.lights[selIndex];
            float sirPDF = light_target_pdf(light, selBary, pos) / weightSum;
            PDFValue = sirPDF / light.SurfaceArea;
            f3 lp = (F3(light.P0) * selBary.x + F3(light.P1) * selBary.y) + F3(light.P2) * selBary.z;
            LightDirection = lp - pos;
            LightNormal = (F3(light.N0) * selBary.x + F3(light.N1) * selBary.y) + F3(light.N2) * selBary.z;
            LightColor = F3(light.LightColor);
        } else {
            uint32_t li = (uint32_t)(c.rand() * (float)lightCount);
            const TbLight& light = c.sc.lights[li];
            f3 bary = random_barycentric(c);
            switch (light.LightType) {
            case TB_LIGHT_TYPE_AREA: {
                f3 lp = (F3(light.P0) * bary.x + F3(light.P1) * bary.y) + F3(light.P2) * bary.z;
                LightDirection = lp - pos;
                LightNormal = (F3(light.N0) * bary.x + F3(light.N1) * bary.y) + F3(light.N2) * bary.z;
                float d = length(LightDirection);
                LightAttenuation = 1.0f / (d * d);
                LightDirection = LightDirection / d;
                break;
            }
            case TB_LIGHT_TYPE_DIRECTIONAL: {
                LightDirection = -F3(light.Direction);
                if (c.rp.settings.DebugValue > 0.0f) { // RayGenCommon.h:236-241 (DebugValue defaults to 1!)
                    LightDirection.x = sin_(c.rp.settings.DebugValue);
                    LightDirection.y = sin_(c.rp.settings.DebugValue2);
                    LightDirection = normalize(LightDirection);
                }
                LightNormal = -LightDirection;
                LightAttenuation = 1.0f;
                break;
            }
            default: break;
            }
            LightColor = F3(light.LightColor);
            PDFValue = 1.0f / (float)lightCount;
            if (light.LightType == TB_LIGHT_TYPE_AREA) PDFValue /= light.SurfaceArea;
        }
    }
}

// ---------------------------------------------------------------- sampling
f3 reorient_around_normal(f3 v, f3 normal) { // kernel.glsl:1001-1015
    f3 tangent;
    if (fabsf(normal.x) > fabsf(normal.y)) tangent = mk3(-normal.z, 0.0f, normal.x) / sqrtf(normal.x * normal.x + normal.z * normal.z);
    else tangent = mk3(0.0f, normal.z, -normal.y) / sqrtf(normal.y * normal.y + normal.z * normal.z);
    f3 bitangent = cross(normal, tangent);
    return normalize((v.x * tangent + v.y * normal) + v.z * bitangent);
}
f3 generate_random_direction(Ctx& c) { // kernel.glsl:991-999 (uniform hemisphere about +z, "2.0 * 3.14")
    float u1 = c.rand(), u2 = c.rand();
    float r = sqrtf(1.0f - u1 * u1);
    float phi = 2.0f * 3.14f * u2;
    return mk3(cos_(phi) * r, sin_(phi) * r, u1);
}
f3 cosine_weighted_direction(f3 normal, float rand0, float rand1, float& pdf) { // kernel.glsl:1025-1041
    float r = sqrtf(rand0);
    float theta = 2.0f * PI * rand1;
    float x = r * cos_(theta);
    float y = sqrtf(fmaxf(EPSILON, 1.0f - rand0));
    float z = r * sin_(theta);
    pdf = y / PI;
    return reorient_around_normal(mk3(x, y, z), normal);
}
f3 importance_sampled_direction(f3 normal, float roughness, float rand0, float rand1, float& pdf) { // kernel.glsl:1048-1064
    float lobe = pow_(1.0f - roughness, 5.0f) * 1000.0f;
    float u1 = rand0, u2 = rand1;
    float theta = 2.0f * PI * u2;
    float phi = acos_(sqrtf(pow_(u1, 1.0f / (lobe + 1.0f))));
    f3 d = mk3(sin_(phi) * cos_(theta), cos_(phi), sin_(phi) * sin_(theta));
    pdf = (lobe + 1.0f) * pow_(cos_(phi), lobe) / (2.0f * PI);
    return reorient_around_normal(d, normal);
}
f3 random_importance_sampled_direction(Ctx& c, f3 normal, float roughness, float& pdf) { // :1096-1099
    float r0 = c.rand();
    float r1 = c.rand();
    return importance_sampled_direction(normal, roughness, r0, r1, pdf);
}
f3 importance_sample_ggx(Ctx& c, f3 incoming, f3 normal, float roughness) { // kernel.glsl:1066-1082
    roughness = fmaxf(MIN_ROUGHNESS, roughness);
    float a = roughness * roughness;
    float a2 = a * a;
    float u1 = c.rand(), u2 = c.rand();
    float theta = 2.0f * PI * u2;
    float phi = acos_(sqrtf((1.0f - u1) / ((a2 - 1.0f) * u1 + 1.0f)));
    f3 d = mk3(sin_(phi) * cos_(theta), cos_(phi), sin_(phi) * sin_(theta));
    f3 h = reorient_around_normal(d, normal);
    return reflect(incoming, h);
}
float importance_sample_ggx_pdf(f3 normal, f3 outgoing, f3 halfVector, float roughness) { // kernel.glsl:1084-1094
    roughness = fmaxf(MIN_ROUGHNESS, roughness);
    float a = roughness * roughness;
    float a2 = a * a;
    float cosTheta = fabsf(dot(normal, halfVector));
    float e = ((a2 - 1.0f) * cosTheta) * cosTheta + 1.0f;
    if (e <= 0.0f) return LARGE_NUMBER;
    float d = a2 / ((PI * e) * e);
    return d * fabsf(dot(halfVector, normal)) / (4.0f * fabsf(dot(outgoing, halfVector)));
}
float ggx_ndf(f3 normal, f3 halfVector, float roughnessSquared) { // kernel.glsl:466-478
    roughnessSquared = fmaxf(roughnessSquared, MIN_ROUGHNESS_SQUARED);
    float a2 = roughnessSquared * roughnessSquared;
    float nDotH = dot(normal, halfVector);
    float denom = PI * pow_((nDotH * nDotH) * (a2 - 1.0f) + 1.0f, 2.0f);
    return a2 / denom;
}
float diffuse_brdf(f3 lightDirection, f3 normal) { return fmaxf(dot(lightDirection, normal), 0.0f) / PI; } // :541-546
f3 half_vector_safe(f3 a, f3 b, f3 normal) { // kernel.glsl:1258-1269
    float aDotB = dot(a, b);
    if (aDotB > (-1.0f + EPSILON)) return normalize(a + b);
    return normal;
}
f3 get_ray_point(const Ray& r, float t) { return r.origin + r.direction * t; }

// refraction + rough-refraction retry shared by SSS entry (kernel.glsl:1531-1563) and exit (:1641-1677).
// Returns false when the path must stop ("Still no luck, call it quits").
enum RefractResult { REFRACTED, REFLECTED, GIVE_UP };
RefractResult refract_or_reflect(Ctx& c, Ray& ray, f3 normal, float nr, float RayDirectionDotN, bool perfectSpec,
                                 float roughness, bool& prevPerfectlySpecular) {
    float discriminant = 1.0f - (nr * nr) * (1.0f - RayDirectionDotN * RayDirectionDotN);
    if (discriminant > EPSILON) {
        f3 refr = normalize(nr * (ray.direction - normal * RayDirectionDotN) - normal * sqrtf(discriminant));
        if (perfectSpec) {
            ray.direction = refr;
            prevPerfectlySpecular = true;
        } else {
            float pdf;
            ray.direction = random_importance_sampled_direction(c, refr, roughness, pdf);
            if (pdf < EPSILON) {
                ray.direction = random_importance_sampled_direction(c, refr, roughness, pdf);
                if (pdf < EPSILON) return GIVE_UP;
            }
        }
        return REFRACTED;
    }
    ray.direction = reflect(ray.direction, normal);
    return REFLECTED;
}

// ------------------------------------------------------------------ Trace
// kernel.glsl:1278-1776
f3 trace(Ctx& c, Ray ray, Ray neighborRay) {
    const TbOutputSettings& S = c.rp.settings;
    f3 accumulatedColor = mk3(0.0f);
    f3 thr = mk3(1.0f); // accumulatedIndirectLightMultiplier
    (void)get_blue_noise(c); // :1283, burns 8 rand() when blue noise is off
    bool bPrevRayWasPerfectlySpecular = false;
    const int MaxBounces = S.MaxBounces;

    for (int i = 0; i < MaxBounces; i++) {
        if (i >= 2) { // russian roulette, :1288-1302
            float p = fmaxf(fmaxf(thr.x, thr.y), thr.z);
            p = fmaxf(p, EPSILON);
            if (p < c.rand()) break;
            else thr *= 1.0f / p;
        }
        bool bFirstRay = (i == 0);
        HitResult hr = intersect(c, ray);
        f3 normal = hr.normal, tangent = hr.tangent;
        f2 uv = hr.uv;
        if (thr.x < EPSILON && thr.y < EPSILON && thr.z < EPSILON) break;

        if (hr.material == -1) {
            accumulatedColor += thr * sample_environment_map(c.sc, ray.direction);
            if (bFirstRay) { c.aovEmissive = mk4(accumulatedColor, 1.0f); c.wroteEmissive = true; }
            break;
        }
        f3 RayPoint = get_ray_point(ray, hr.t);
        ray.origin = RayPoint + normal * EPSILON;
        float RayDirectionDotN = dot(normal, ray.direction);
        bool IsBackside = RayDirectionDotN > 0.0f;
        Material material = get_material(c, hr.material, uv, IsBackside);
        f3 detailNormal = get_detail_normal(c, material, normal, tangent, uv);
        if (i == 0) {
            f3 nrp = get_ray_point(neighborRay, hr.t);
            c.worldPosition += RayPoint;                      // OutputPrimaryWorldPosition :585-591
            c.distanceToNeighbor += length(nrp - RayPoint);
            c.aovNormal = mk4(detailNormal, 1.0f);            // OutputPrimaryNormal
            c.aovDepth = saturate(hr.t / S.MaxZ); c.wroteDepth = true; // OutputDistanceToFirstHit
            if (c.selected()) { c.statDistance = hr.t; c.statMaterial = hr.material; c.wroteStats = true; }
            if (S.OutputType == TB_OUTPUT_HEATMAP) break;     // TerminateAfterPrimaryHit
        }
        bool IsInside = IsBackside;
        float CurrentIOR = IsInside ? material.IOR : AIR_IOR;
        float NewIOR = IsInside ? AIR_IOR : material.IOR;
        if (IsInside) { normal = -normal; RayDirectionDotN = -RayDirectionDotN; detailNormal = -detailNormal; }
        float ReflectionCoefficient = material.SpecularCoef;

        bool bSpecularRay = false;
        if (AllowsSpecular(material)) {
            if (IsMetallic(material) || IsHairMaterial(material)) bSpecularRay = true;
            else bSpecularRay = c.rand() < 0.5f;
        }
        bool bPerfectSpec = bSpecularRay && UsePerfectSpecularOptimization(material.roughness);
        if (bPrevRayWasPerfectlySpecular || bFirstRay || !IsLight(material) || !S.EnableNextEventEstimation)
            accumulatedColor += thr * material.emissive;
        if (IsLight(material)) break;

        float lightPDF, lightAttenuation;
        f3 lightDirection, lightColor, lightNormal;
        get_one_light_sample(c, RayPoint, lightDirection, lightColor, lightPDF, lightNormal, lightAttenuation);
        if (!bPerfectSpec) {
            if (lightPDF > EPSILON && dot(lightDirection, lightNormal) < 0.0f) {
                f3 ShadowMultiplier = mk3(1.0f);
                Ray shadowFeeler = {RayPoint + normal * EPSILON, lightDirection};
                HitResult sh = intersect(c, shadowFeeler);
                if (sh.material != -1) {
                    float LdotN = dot(sh.normal, lightDirection);
                    bool shBack = LdotN > 0.0f;
                    Material sm = get_material(c, sh.material, sh.uv, shBack);
                    if (!IsLight(sm)) ShadowMultiplier = mk3(0.0f);
                }
                float lightMultiplier = lightAttenuation * diffuse_brdf(lightDirection, detailNormal) * fabsf(dot(lightNormal, lightDirection)) / lightPDF;
                accumulatedColor += (((thr * material.albedo) * lightMultiplier) * ShadowMultiplier) * lightColor;
            }
        }

        f3 previousDirection = ray.direction;
        bPrevRayWasPerfectlySpecular = bPerfectSpec;
        if (bSpecularRay) {
            ray.direction = importance_sample_ggx(c, ray.direction, normal, material.roughness);
        } else if (IsSubsurfaceScattering(material)) {
            float nr = CurrentIOR / NewIOR;
            if (refract_or_reflect(c, ray, normal, nr, RayDirectionDotN, bPerfectSpec, material.roughness, bPrevRayWasPerfectlySpecular) == GIVE_UP) break;
            bool noScatter = material.scattering.x < EPSILON;
            float DistancePerScatter = 1.0f / (((material.scattering.x + material.scattering.y) + material.scattering.z) / 3.0f);
            float maxTravelDistance = noScatter ? LARGE_NUMBER : DistancePerScatter;
            bool exitting = (material.Flags & TB_SINGLE_SIDED_MATERIAL_FLAG) != 0;
            const int MAX_SSS_BOUNCES = 100;
            for (int k = 0; k < MAX_SSS_BOUNCES && !exitting; k++) {
                float travelDistance = fmaxf(-log_(c.rand()), 0.1f) * maxTravelDistance;
                HitResult w = intersect(c, ray);
                normal = w.normal; tangent = w.tangent; uv = w.uv;
                if (w.material == -1) { thr = mk3(0.0f); break; }
                float tt = fminf(travelDistance, w.t);
                float distBeforeScatter = tt;
                exitting = tt < travelDistance || noScatter;
                bool lastRay = (k == MAX_SSS_BOUNCES - 1);
                if (lastRay && !exitting) thr = mk3(0.0f);
                RayPoint = get_ray_point(ray, tt);
                ray.origin = RayPoint + normal * EPSILON;
                f3 beer = exp3((-distBeforeScatter) * material.absorption);
                thr *= beer;
                if (exitting) {
                    RayDirectionDotN = dot(normal, ray.direction);
                    if (RayDirectionDotN >= 0.0f) { normal = -normal; RayDirectionDotN = -RayDirectionDotN; }
                    float nr2 = NewIOR / CurrentIOR;
                    RefractResult rr = refract_or_reflect(c, ray, normal, nr2, RayDirectionDotN, bPerfectSpec, material.roughness, bPrevRayWasPerfectlySpecular);
                    if (rr == GIVE_UP) break;           // breaks the walk loop only (:1661)
                    if (rr == REFLECTED) exitting = false;
                    previousDirection = ray.direction;
                } else {
                    f3 nd = generate_random_direction(c); // GenerateNewDirectionFromBSDF(dir, 0.0): isotropic, pdf "1.0"
                    ray.direction = nd;
                    thr /= 1.0f;
                }
            }
            continue; // :1690
        } else {
            float pdf;
            float r0 = c.rand();
            float r1 = c.rand();
            ray.direction = cosine_weighted_direction(normal, r0, r1, pdf);
        }

        float DiffusePDF = dot(ray.direction, normal) / PI;
        if (AllowsSpecular(material)) {
            f3 halfVector = half_vector_safe(-previousDirection, ray.direction, normal);
            float SpecularPDF = importance_sample_ggx_pdf(normal, ray.direction, halfVector, material.roughness);
            float PDFValue = IsMetallic(material) ? SpecularPDF : lerp(SpecularPDF, DiffusePDF, 0.5f);
            thr /= PDFValue;
        } else {
            thr /= DiffusePDF;
        }
        if (bFirstRay) { c.aovEmissive = mk4(material.emissive, 1.0f); c.wroteEmissive = true; }
        // (IsLight(material) cannot be true here: handled above)
        bool bRemoveAlbedo = (S.RenderMode == TB_RENDER_REALTIME) && bFirstRay;
        f3 albedo = bRemoveAlbedo ? mk3(1.0f) : material.albedo;
        if (IsMetallic(material)) {
            f3 halfVector = normalize(-previousDirection + ray.direction);
            float roughnessSquared = fmaxf(material.roughness * material.roughness, MIN_ROUGHNESS_SQUARED);
            float specular = ggx_ndf(detailNormal, halfVector, roughnessSquared) /
                             ((4.0f * fabsf(dot(-previousDirection, halfVector))) * fmaxf(fabsf(dot(-previousDirection, normal)), fabsf(dot(ray.direction, normal))));
            thr *= (specular * albedo) * saturate(dot(ray.direction, normal));
        } else if (AllowsSpecular(material)) {
            f3 halfVector = half_vector_safe(-previousDirection, ray.direction, normal);
            float fresnel = ReflectionCoefficient + (1.0f - ReflectionCoefficient) * pow_(fabsf(1.0f - dot(-previousDirection, halfVector)), 5.0f);
            float diffuseMultiplier = (((28.0f / (23.0f * PI)) * (1.0f - ReflectionCoefficient)) *
                                       (1.0f - pow_(1.0f - 0.5f * dot(-previousDirection, normal), 5.0f))) *
                                      (1.0f - pow_(1.0f - 0.5f * dot(ray.direction, normal), 5.0f));
            f3 diffuse = albedo * diffuseMultiplier;
            float roughnessSquared = fmaxf(material.roughness * material.roughness, MIN_ROUGHNESS_SQUARED);
            float specular = ggx_ndf(detailNormal, halfVector, roughnessSquared) /
                             ((4.0f * fabsf(dot(-previousDirection, halfVector))) * fmaxf(fabsf(dot(-previousDirection, normal)), fabsf(dot(ray.direction, normal))));
            f3 mult = (diffuse + fresnel * specular) * saturate(dot(ray.direction, normal));
            thr *= mult;
        } else {
            thr *= albedo * diffuse_brdf(ray.direction, detailNormal);
        }
        if (bFirstRay && S.OutputType != TB_OUTPUT_HEATMAP) c.aovAlbedo = mk4(material.albedo, 1.0f); // OutputPrimaryAlbedo
    }
    return accumulatedColor;
}

float gaussian(float x, float mu, float sigma) { // kernel.glsl:1800-1803
    float d = x - mu;
    return 1.0f / sqrtf((2.0f * PI) * sigma * sigma) * exp_(-(d * d) / ((2.0f * sigma) * sigma));
}

f3 lens_position(const TbCamera& cam, f2 uv, float aspect) { // kernel.glsl:1786-1798 (view matrix == identity)
    f3 p = F3(cam.Position);
    float lensWidth = cam.LensHeight * aspect;
    p += ((F3(cam.Right) * (uv.x * 2.0f - 1.0f)) * lensWidth) / 2.0f;
    p += ((F3(cam.Up) * (uv.y * 2.0f - 1.0f)) * cam.LensHeight) / 2.0f;
    return p;
}

// PathTrace, kernel.glsl:1805-1921
f4 path_trace(Ctx& c, f2 pixelCoord) {
    const TbOutputSettings& S = c.rp.settings;
    const TbCamera& cam = c.rp.camera;
    f2 res = mk2((float)c.W, (float)c.H);
    f2 pixelUVSize = mk2(1.0f / res.x, 1.0f / res.y);
    f2 uv = pixelCoord * pixelUVSize;
    BlueNoiseData bn = get_blue_noise(c);
    f2 PrimaryJitter = bn.PrimaryJitter; // FixedPixelOffset = (-1,-1): TracerBoy.cpp:2838
    f2 off = PrimaryJitter - mk2(0.5f, 0.5f);
    float pixelRadius = S.FilterWidth / 2.0f;
    float filterWeight = 1.0f;
    switch (S.FilterType) {
    case TB_FILTER_TRIANGLE: filterWeight = fmaxf(0.5f - fabsf(off.x), 0.5f - fabsf(off.y)); break;
    case TB_FILTER_GAUSSIAN: {
        float sigma = 0.8f;
        float eX = gaussian(1.0f, 0.0f, sigma), eY = gaussian(1.0f, 0.0f, sigma);
        filterWeight = fmaxf(0.0f, gaussian(off.x * 2.0f, 0.0f, sigma) - eX) * fmaxf(0.0f, gaussian(off.y * 2.0f, 0.0f, sigma) - eY);
        break;
    }
    default: filterWeight = 1.0f; break;
    }
    uv = uv + (off * pixelUVSize) * (pixelRadius * 2.0f);
    float aspect = res.x / res.y;
    f3 camPos = F3(cam.Position);
    f3 focalPoint = camPos - cam.FocalDistance * normalize(F3(cam.LookAt) - camPos);
    f3 lensPoint = lens_position(cam, uv, aspect);
    f3 neighborLensPoint = lens_position(cam, uv + pixelUVSize, aspect);
    Ray cameraRay = {focalPoint, normalize(lensPoint - focalPoint)};
    Ray neighborRay = {focalPoint, normalize(neighborLensPoint - focalPoint)};
    if (S.DOFFocalDistance > 0.0f) { // :1889-1902
        f3 FocusPoint = get_ray_point(cameraRay, S.DOFFocalDistance);
        float Radius = sqrtf(bn.DOFJitter.x) * S.ApertureWidth;
        float Theta = (bn.DOFJitter.y * 2.0f) * PI;
        f2 fj = mk2(cos_(Theta) * Radius, sin_(Theta) * Radius);
        cameraRay.origin = cameraRay.origin + (fj.x * F3(cam.Right) + fj.y * F3(cam.Up));
        cameraRay.direction = normalize(FocusPoint - cameraRay.origin);
    }
    f3 col = trace(c, cameraRay, neighborRay);
    if (S.FireflyClampValue >= EPSILON) col = min3(col, S.FireflyClampValue);
    return mk4(col * filterWeight, filterWeight);
}

static PathTraceFn g_pathTraceOverride = nullptr;
void set_path_trace_override(PathTraceFn fn) { g_pathTraceOverride = fn; }

// SoftwareRayTraceCS.hlsl:36-50 + RayTraceCommon (RayGenCommon.h:690-728)
void render_frame(const Scene& s, const RenderParams& p, FrameBuffers& fb, int numThreads) {
    const uint32_t W = fb.width, H = fb.height;
    uint64_t rays = 0, tris = 0, boxes = 0;
    (void)numThreads;
#ifdef _OPENMP
    if (numThreads > 0) omp_set_num_threads(numThreads);
#endif
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : rays, tris, boxes)
    for (int64_t y = 0; y < (int64_t)H; y++) {
        if (((uint32_t)y / 8u) % p.rowStride != p.rowOffset) continue; // another shard's band: its pixels stay untouched (zero)
        for (uint32_t x = 0; x < W; x++) {
            Ctx c(s, p);
            c.W = W; c.H = H; c.px = x; c.py = (uint32_t)y;
            c.worldPosition = mk3(0.0f); c.distanceToNeighbor = 0.0f;
            c.aovAlbedo = mk4(0, 0, 0, 1.0f); c.aovNormal = mk4(0, 0, 0, 1.0f); // ClearAOVs :650-654
            c.wroteEmissive = c.wroteDepth = c.wroteStats = false;
            c.firstIntersect = true; c.primGeom = c.primPrim = 0xffffffffu;
            c.tris = c.boxes = c.rays = 0;
            c.seed = hash13(mk3((float)x, (float)y, (float)p.frame));
            size_t pi = (size_t)y * W + x;
            f2 dispatchUV = mk2((float)x + 0.5f, (float)y + 0.5f) / mk2((float)W, (float)H);
            f2 uv = mk2(0.0f, 1.0f) + dispatchUV * mk2(1.0f, -1.0f);
            f4 color = g_pathTraceOverride ? g_pathTraceOverride(c, uv * mk2((float)W, (float)H)) : path_trace(c, uv * mk2((float)W, (float)H));
            f4 outc = mk4(0, 0, 0, 0);
            if (!isnan_(color.x) && !isnan_(color.y) && !isnan_(color.z) && !isnan_(color.w)) outc = outc + color;
            TbFloat4 wp = {c.worldPosition.x, c.worldPosition.y, c.worldPosition.z, c.distanceToNeighbor};
            fb.aovWorldPos[p.frame % 2][pi] = wp;
            bool realtime = p.settings.RenderMode == TB_RENDER_REALTIME;
            // On one process clearAccum == (frame == 0) and this is RayGenCommon.h:721-727 verbatim; a
            // sample-sharded process starts at frame > 0 and must still start from an empty buffer.
            bool clear = realtime || p.clearAccum;
            TbFloat4 prev = clear ? TbFloat4{0, 0, 0, 0} : fb.accum[pi];
            fb.accum[pi] = {outc.x + prev.x, outc.y + prev.y, outc.z + prev.z, outc.w + prev.w};
            if (!realtime) {
                bool take = p.frame == 0 || c.rand() < 0.5f;
                if (take) {
                    TbFloat4 pj = p.clearAccum ? TbFloat4{0, 0, 0, 0} : fb.jittered[pi];
                    fb.jittered[pi] = {outc.x + pj.x, outc.y + pj.y, outc.z + pj.z, outc.w + pj.w};
                } else if (p.clearAccum) fb.jittered[pi] = {0, 0, 0, 0};
            }
            fb.aovAlbedo[pi] = {c.aovAlbedo.x, c.aovAlbedo.y, c.aovAlbedo.z, c.aovAlbedo.w};
            fb.aovNormal[pi] = {c.aovNormal.x, c.aovNormal.y, c.aovNormal.z, c.aovNormal.w};
            if (c.wroteEmissive) fb.aovEmissive[pi] = {c.aovEmissive.x, c.aovEmissive.y, c.aovEmissive.z, c.aovEmissive.w};
            if (c.wroteDepth) fb.aovDepth[pi] = c.aovDepth;
            fb.primaryHit[2 * pi] = c.primGeom; fb.primaryHit[2 * pi + 1] = c.primPrim;
            fb.counters[2 * pi] = c.tris; fb.counters[2 * pi + 1] = c.boxes;
            if (c.wroteStats) { fb.stats.SelectedPixelDistance = c.statDistance; fb.stats.SelectedMaterialID = c.statMaterial; }
            rays += c.rays; tris += c.tris; boxes += c.boxes;
        }
    }
    fb.raysTraced += rays; fb.trianglesTested += tris; fb.boxesTested += boxes;
}

float hash13_public(float x, float y, float z) { return hash13(mk3(x, y, z)); }
float halton_public(int b, int i) { return halton(b, i); }

} // namespace oracle
