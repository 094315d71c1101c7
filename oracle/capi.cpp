// capi.cpp — CPU ORACLE (test infrastructure): C entry points for ctypes.
#include <chrono>
#include <cstdio>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../tracerboy_b200/csrc/common/tb_vec.h"
#include "oracle.h"
#include "glue.h"
#include "ref/synthetic_tracer.h"

using namespace oracle;

struct OracleHandle {
    Scene scene;
    FrameBuffers fb;
    TbCamera camera;
    std::string err;
    uint32_t samples = 0;
    uint32_t shardOffset = 0, shardStride = 1; // frame f = offset + local * stride (multi-process sharding tests)
    int selX = -1, selY = -1;
    uint32_t rowOffset = 0, rowStride = 1;
};

#define ORACLE_API extern "C" __attribute__((visibility("default")))

ORACLE_API OracleHandle* oracle_create() { return new OracleHandle(); }
ORACLE_API void oracle_destroy(OracleHandle* h) { delete h; }
ORACLE_API const char* oracle_last_error(OracleHandle* h) { return h->err.c_str(); }
ORACLE_API int oracle_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORACLE_API int oracle_load_scene(OracleHandle* h, const char* tbscene, const char* blueNoiseBin) {
    if (!load_tbscene(h->scene, tbscene, h->err)) return -3;
    h->camera = h->scene.camera;
    h->scene.blueNoise.assign(2 * 256 * 256 * 4, 0);
    if (blueNoiseBin && *blueNoiseBin) {
        FILE* f = fopen(blueNoiseBin, "rb");
        if (!f || fread(h->scene.blueNoise.data(), 1, h->scene.blueNoise.size(), f) != h->scene.blueNoise.size()) {
            if (f) fclose(f);
            h->err = std::string("cannot read blue noise ") + blueNoiseBin;
            return -3;
        }
        fclose(f);
    }
    h->samples = 0;
    return 0;
}
ORACLE_API int oracle_build_bvh(OracleHandle* h, int treeletPasses) { return build_bvh(h->scene, treeletPasses, h->err) ? 0 : -1; }
ORACLE_API uint64_t oracle_bvh_size(OracleHandle* h) { return h->scene.bvh.size(); }
ORACLE_API int oracle_get_bvh(OracleHandle* h, void* dst, uint64_t bytes) {
    if (bytes < h->scene.bvh.size()) return -1;
    memcpy(dst, h->scene.bvh.data(), h->scene.bvh.size());
    return 0;
}
ORACLE_API int oracle_get_hierarchy(OracleHandle* h, uint32_t* dst, uint64_t words) { // test hook (ComputeAABBs pin)
    if (words < h->scene.hierarchy.size()) return -1;
    memcpy(dst, h->scene.hierarchy.data(), 4 * h->scene.hierarchy.size());
    return 0;
}
ORACLE_API uint32_t oracle_max_treelet_climb(OracleHandle* h) { return h->scene.maxTreeletClimb; }
ORACLE_API uint32_t oracle_num_triangles(OracleHandle* h) { return h->scene.numPrims; }
ORACLE_API int oracle_trace_rays(OracleHandle* h, const TbRay* rays, uint64_t n, TbHit* hits) {
    if (h->scene.bvh.empty()) { h->err = "no bvh"; return -7; }
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t)n; i++) trace_ray(h->scene, rays[i], hits[i]);
    return 0;
}
// analysis hook (tools/simd_model.py): the visit-type sequence of each ray, concatenated; offsets[n + 1]
ORACLE_API int oracle_trace_rays_visits(OracleHandle* h, const TbRay* rays, uint64_t n, TbHit* hits, uint8_t* seq, uint64_t seqCap, uint64_t* offsets) {
    if (h->scene.bvh.empty()) { h->err = "no bvh"; return -7; }
    std::vector<uint8_t> log;
    oracle::set_visit_log(&log);
    uint64_t at = 0;
    for (uint64_t i = 0; i < n; i++) {
        log.clear();
        trace_ray(h->scene, rays[i], hits[i]);
        offsets[i] = at;
        if (at + log.size() > seqCap) { oracle::set_visit_log(nullptr); h->err = "sequence buffer too small"; return -1; }
        memcpy(seq + at, log.data(), log.size());
        at += log.size();
    }
    offsets[n] = at;
    oracle::set_visit_log(nullptr);
    return 0;
}
ORACLE_API int oracle_get_camera(OracleHandle* h, TbCamera* c) { *c = h->camera; return 0; }
ORACLE_API int oracle_set_camera(OracleHandle* h, const TbCamera* c) { h->camera = *c; h->samples = 0; return 0; }
ORACLE_API int oracle_resize(OracleHandle* h, uint32_t w, uint32_t hh) { h->fb.resize(w, hh); h->samples = 0; return 0; }
ORACLE_API int oracle_select_pixel(OracleHandle* h, int x, int y) { h->selX = x; h->selY = y; return 0; }
ORACLE_API int oracle_invalidate(OracleHandle* h) { h->samples = 0; return 0; }
ORACLE_API uint32_t oracle_samples(OracleHandle* h) { return h->samples; }
ORACLE_API int oracle_set_samples(OracleHandle* h, uint32_t s) { h->samples = s; return 0; }
// nSamples x one-sample Render; returns elapsed seconds through *seconds.
ORACLE_API int oracle_render(OracleHandle* h, const TbOutputSettings* s, uint32_t nSamples, float time, int threads, double* seconds) {
    if (h->scene.bvh.empty() || h->fb.width == 0) { h->err = "render before load/resize"; return -7; }
    auto t0 = std::chrono::high_resolution_clock::now();
    for (uint32_t i = 0; i < nSamples; i++) {
        RenderParams p;
        p.settings = *s; p.camera = h->camera; p.time = time; p.frame = h->shardOffset + h->samples * h->shardStride;
        p.clearAccum = h->samples == 0;
        p.selectedX = h->selX; p.selectedY = h->selY;
        p.rowOffset = h->rowOffset; p.rowStride = h->rowStride;
        render_frame(h->scene, p, h->fb, threads);
        h->samples++;
    }
    auto t1 = std::chrono::high_resolution_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    return 0;
}
ORACLE_API void oracle_set_literal_rcp(int on) { oracle::set_literal_rcp(on != 0); }
ORACLE_API void oracle_set_literal_mode(int mask) { oracle::set_literal_mode(mask); }
ORACLE_API int oracle_set_shard(OracleHandle* h, uint32_t offset, uint32_t stride) {
    if (stride == 0 || offset >= stride) return -1;
    h->shardOffset = offset; h->shardStride = stride; h->samples = 0;
    return 0;
}
ORACLE_API int oracle_set_row_shard(OracleHandle* h, uint32_t offset, uint32_t stride) { // tb_set_row_shard
    if (stride == 0 || offset >= stride) return -1;
    h->rowOffset = offset; h->rowStride = stride; h->samples = 0;
    return 0;
}
// pinned intrinsics and noise functions, exposed so tests can pin them against independent numpy restatements
ORACLE_API void oracle_math_eval(int fn, const float* x, const float* y, float* out, uint64_t n) {
    using namespace tbm;
    for (uint64_t i = 0; i < n; i++) {
        switch (fn) {
        case 0: out[i] = sin_(x[i]); break;
        case 1: out[i] = cos_(x[i]); break;
        case 2: out[i] = acos_(x[i]); break;
        case 3: out[i] = atan2_(y[i], x[i]); break;
        case 4: out[i] = exp_(x[i]); break;
        case 5: out[i] = log_(x[i]); break;
        case 6: out[i] = pow_(x[i], y[i]); break;
        case 7: out[i] = oracle::hash13_public(x[3 * i], x[3 * i + 1], x[3 * i + 2]); break;
        case 8: out[i] = oracle::halton_public((int)y[i], (int)x[i]); break;
        default: out[i] = 0.0f;
        }
    }
}
ORACLE_API void oracle_karras(const uint32_t* codes, uint32_t n, uint32_t* out3) { oracle::karras_public(codes, n, out3); }
ORACLE_API void oracle_treelet(uint32_t* H3, float* aabb6, uint32_t n, uint32_t root) { oracle::treelet_public(H3, aabb6, n, root); }
ORACLE_API void oracle_load_primitives(OracleHandle* h, void* prims40, void* meta12) { oracle::load_primitives_public(h->scene, prims40, meta12); }
ORACLE_API const void* oracle_scene_positions(OracleHandle* h) { return h->scene.positions.data(); }
ORACLE_API uint64_t oracle_scene_num_positions(OracleHandle* h) { return h->scene.positions.size(); }
// Top-level structure over reference-layout bottom-level byte buffers: blas[i] / blasBytes[i]; instance desc field
// AccelerationStructure = index into that list. Two calls: size query (out == nullptr), then the bytes.
ORACLE_API int64_t oracle_build_tlas(const TbInstanceDesc* inst, uint32_t n, const uint8_t* const* blas, uint32_t numBlas, uint8_t* out, uint64_t cap) {
    std::vector<const uint8_t*> list(blas, blas + numBlas);
    std::vector<uint8_t> bytes;
    std::string err;
    if (!oracle::build_tlas(inst, n, list, bytes, err)) return -1;
    if (out) { if (cap < bytes.size()) return -2; memcpy(out, bytes.data(), bytes.size()); }
    return (int64_t)bytes.size();
}
ORACLE_API int oracle_trace_rays_tlas(const uint8_t* tlas, const uint8_t* const* blas, uint32_t numBlas, const TbRay* rays, uint64_t n, TbHit* hits) {
    std::vector<const uint8_t*> list(blas, blas + numBlas);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)n; i++) oracle::trace_ray_tlas(tlas, list, rays[i], hits[i]);
    return 0;
}
ORACLE_API void oracle_inverse_affine(const float* m12, float* out12) { oracle::inverse_affine_public(m12, out12); }
ORACLE_API void oracle_transform_aabb(const float* mn, const float* mx, const float* m12, float* out6) { oracle::transform_aabb_public(mn, mx, m12, out6); }
// PERFORM_UPDATE: replace the scene's vertex positions (same count) and refit the acceleration structure
ORACLE_API int oracle_update_bvh(OracleHandle* h, const float* positions, uint64_t count) {
    if (!positions || count != h->scene.positions.size()) { h->err = "position count differs from the scene's"; return -1; }
    memcpy(h->scene.positions.data(), positions, sizeof(TbFloat3) * count);
    return oracle::update_bvh(h->scene, h->err) ? 0 : -1;
}
ORACLE_API void oracle_scene_box(const void* prims40, uint32_t n, float* out6) { oracle::scene_box_public(prims40, n, out6); }
ORACLE_API void oracle_centroid(const void* prim40, float* out3) { oracle::centroid_public(prim40, out3); }
ORACLE_API int oracle_sorts_before(uint32_t codeA, uint32_t indexA, uint32_t codeB, uint32_t indexB) { return oracle::sorts_before_public(codeA, indexA, codeB, indexB); }
ORACLE_API void oracle_treelet_pass(uint32_t* H3, const void* prims40, uint32_t n, uint32_t minTris, uint32_t* maxClimb) {
    oracle::treelet_pass_public(H3, prims40, n, minTris, maxClimb);
}
ORACLE_API void oracle_leaf_box(const float* v9, float* c3, float* h3) { oracle::leaf_box_public(v9, c3, h3); }
ORACLE_API void oracle_parent_box(const float* ac, const float* ah, const float* bc, const float* bh, float* c3, float* h3) { oracle::parent_box_public(ac, ah, bc, bh, c3, h3); }
// test hooks for oracle/ref/ref_raygen.cpp: the restated light sampling and environment lookup on caller-provided data
ORACLE_API void oracle_light_sample(const TbLight* lights, uint32_t n, uint32_t nee, uint32_t sir, float debug1, float debug2, float time,
                                    const float* pos, float* seed, float* out12) {
    Scene sc;
    sc.lights.assign(lights, lights + n);
    RenderParams rp;
    memset(&rp.settings, 0, sizeof(rp.settings));
    rp.settings.EnableNextEventEstimation = nee; rp.settings.EnableSamplingImportanceResampling = sir;
    rp.settings.DebugValue = debug1; rp.settings.DebugValue2 = debug2;
    rp.time = time;
    Ctx c(sc, rp);
    c.seed = *seed;
    tbm::f3 dir, col, nrm; float pdf, att;
    get_one_light_sample(c, tbm::mk3(pos[0], pos[1], pos[2]), dir, col, pdf, nrm, att);
    *seed = c.seed;
    float o[12] = {dir.x, dir.y, dir.z, col.x, col.y, col.z, nrm.x, nrm.y, nrm.z, pdf, att, 0.0f};
    memcpy(out12, o, sizeof(o));
}
ORACLE_API void oracle_env(const float* rgba, uint32_t w, uint32_t h, const float* transform12, const float* scale3, const float* v, float* out3) {
    Scene sc;
    sc.images.resize(1);
    sc.images[0].width = w; sc.images[0].height = h; sc.images[0].format = 0;
    sc.images[0].data.assign((const uint8_t*)rgba, (const uint8_t*)rgba + 16ull * w * h);
    sc.envImage = 0;
    for (int r = 0; r < 3; r++) sc.envTransform[r] = TbFloat4{transform12[4 * r], transform12[4 * r + 1], transform12[4 * r + 2], transform12[4 * r + 3]};
    sc.envColorScale = TbFloat3{scale3[0], scale3[1], scale3[2]};
    tbm::f3 c = sample_environment_map(sc, tbm::mk3(v[0], v[1], v[2]));
    out3[0] = c.x; out3[1] = c.y; out3[2] = c.z;
}
// test hooks for the pin of IntersectWithMaxDistance + the SharedHitGroup.h geometry fetch (tests/test_cpu_oracle.py)
ORACLE_API int oracle_scene_arrays(OracleHandle* h, const void** geoms, uint32_t* numGeoms, const void** indices, const void** vertices) {
    *geoms = h->scene.geoms.data(); *numGeoms = (uint32_t)h->scene.geoms.size();
    *indices = h->scene.indices.data(); *vertices = h->scene.vertices.data();
    return 0;
}
// out: 12 floats per ray = t, material, normal.xyz, tangent.xyz, uv.xy, TrianglesTested, BoxesTested
ORACLE_API int oracle_intersect(OracleHandle* h, const TbRay* rays, uint64_t n, float* out12) {
    RenderParams rp;
    memset(&rp.settings, 0, sizeof(rp.settings));
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        Ctx c(h->scene, rp);
        c.firstIntersect = false; c.rays = c.tris = c.boxes = 0;
        Ray ray{tbm::mk3(rays[i].Origin[0], rays[i].Origin[1], rays[i].Origin[2]), tbm::mk3(rays[i].Direction[0], rays[i].Direction[1], rays[i].Direction[2])};
        HitResult r = intersect(c, ray, rays[i].TMax);
        float* o = out12 + 12 * i;
        o[0] = r.t; o[1] = (float)r.material; o[2] = r.normal.x; o[3] = r.normal.y; o[4] = r.normal.z;
        o[5] = r.tangent.x; o[6] = r.tangent.y; o[7] = r.tangent.z; o[8] = r.uv.x; o[9] = r.uv.y; o[10] = (float)c.tris; o[11] = (float)c.boxes;
    }
    return 0;
}
// test hooks for the pin of the material / texture fetch (GetMaterialInternal, GetDetailNormal, GetTextureData)
ORACLE_API const void* oracle_scene_ptr(OracleHandle* h) { return &h->scene; }
ORACLE_API uint32_t oracle_num_materials(OracleHandle* h) { return (uint32_t)h->scene.materials.size(); }
ORACLE_API uint32_t oracle_num_textures(OracleHandle* h) { return (uint32_t)h->scene.textures.size(); }
static void pack_material(const Material& m, void* out21) {
    TbMaterial t;
    memset(&t, 0, sizeof(t));
    t.albedo = {m.albedo.x, m.albedo.y, m.albedo.z}; t.albedoIndex = m.albedoIndex; t.alphaIndex = m.alphaIndex;
    t.normalMapIndex = m.normalMapIndex; t.emissiveIndex = m.emissiveIndex; t.specularMapIndex = m.specularMapIndex;
    t.IOR = m.IOR; t.absorption = {m.absorption.x, m.absorption.y, m.absorption.z}; t.roughness = m.roughness;
    t.scattering = {m.scattering.x, m.scattering.y, m.scattering.z}; t.emissive = {m.emissive.x, m.emissive.y, m.emissive.z};
    t.Flags = m.Flags; t.SpecularCoef = m.SpecularCoef;
    memcpy(out21, &t, sizeof(t));
}
ORACLE_API void oracle_material(OracleHandle* h, float time, int materialId, const float* uv, int backside, float* seed, void* out21) {
    RenderParams rp;
    memset(&rp.settings, 0, sizeof(rp.settings));
    rp.time = time;
    Ctx c(h->scene, rp);
    c.seed = *seed;
    Material m = get_material_internal(c, materialId, tbm::mk2(uv[0], uv[1]), backside != 0);
    *seed = c.seed;
    pack_material(m, out21);
}
ORACLE_API void oracle_detail_normal(OracleHandle* h, uint32_t enableNormalMaps, int materialId, const float* normal, const float* tangent,
                                     const float* uv, float* out3) {
    RenderParams rp;
    memset(&rp.settings, 0, sizeof(rp.settings));
    rp.settings.EnableNormalMaps = enableNormalMaps;
    Ctx c(h->scene, rp);
    tbm::f3 n = get_detail_normal(c, load_material(h->scene.materials[materialId]), tbm::mk3(normal[0], normal[1], normal[2]),
                                  tbm::mk3(tangent[0], tangent[1], tangent[2]), tbm::mk2(uv[0], uv[1]));
    out3[0] = n.x; out3[1] = n.y; out3[2] = n.z;
}
ORACLE_API void oracle_texture(OracleHandle* h, uint32_t textureIndex, const float* uv, float* out4) {
    tbm::f4 t = get_texture_data(h->scene, textureIndex, tbm::mk2(uv[0], uv[1]));
    out4[0] = t.x; out4[1] = t.y; out4[2] = t.z; out4[3] = t.w;
}
// test hook for the pin of the per-pixel wrapper (RayTraceCommon, AOV writers, GetBlueNoise): render_frame driven by the
// synthetic stand-in for PathTrace that oracle/ref/ref_frame.cpp drives the reference text with.
namespace {
struct CtxSink {
    Ctx& c;
    float rand() { return c.rand(); }
    void blue_noise(float o[8]) {
        BlueNoiseData d = get_blue_noise(c);
        o[0] = d.PrimaryJitter.x; o[1] = d.PrimaryJitter.y; o[2] = d.SecondaryRayDirection.x; o[3] = d.SecondaryRayDirection.y;
        o[4] = d.AreaLightJitter.x; o[5] = d.AreaLightJitter.y; o[6] = d.DOFJitter.x; o[7] = d.DOFJitter.y;
    }
    // the writers as core.cpp's path_trace performs them on the pixel context
    void albedo(const float* a, float) { c.aovAlbedo = tbm::mk4(a[0], a[1], a[2], 1.0f); }
    void normal(const float* n) { c.aovNormal = tbm::mk4(n[0], n[1], n[2], 1.0f); }
    void world_position(const float* p, float d) { c.worldPosition += tbm::mk3(p[0], p[1], p[2]); c.distanceToNeighbor += d; }
    void distance(float D) {
        c.aovDepth = tbm::saturate(D / c.rp.settings.MaxZ); c.wroteDepth = true;
        if (c.selected()) { c.statDistance = D; c.wroteStats = true; }
    }
    void material(int id) { if (c.selected()) { c.statMaterial = id; c.wroteStats = true; } }
    void emissive(const float* e) { c.aovEmissive = tbm::mk4(e[0], e[1], e[2], 1.0f); c.wroteEmissive = true; }
};
tbm::f4 synthetic_override(Ctx& c, tbm::f2 pixelCoord) {
    CtxSink s{c};
    float out[4];
    synthetic_path(s, pixelCoord.x, pixelCoord.y, c.rp.frame, out);
    return tbm::mk4(out[0], out[1], out[2], out[3]);
}
} // namespace
ORACLE_API void oracle_enable_synthetic_tracer(int on) { set_path_trace_override(on ? synthetic_override : nullptr); }
ORACLE_API uint32_t oracle_morton(const float* centroid, const float* smin, const float* smax) {
    return oracle::morton_public(centroid, smin, smax);
}
ORACLE_API int oracle_get_counts(OracleHandle* h, uint64_t* out3) {
    out3[0] = h->fb.raysTraced; out3[1] = h->fb.boxesTested; out3[2] = h->fb.trianglesTested;
    return 0;
}
ORACLE_API int oracle_get_stats(OracleHandle* h, TbReadbackStats* s) { *s = h->fb.stats; return 0; }
ORACLE_API int oracle_readback(OracleHandle* h, uint32_t kind, void* dst, uint64_t bytes) {
    FrameBuffers& fb = h->fb;
    size_t n = (size_t)fb.width * fb.height;
    const void* src = nullptr;
    size_t sz = 0;
    std::vector<float> tmp;
    switch (kind) {
    case TB_BUF_ACCUM_RGBW: src = fb.accum.data(); sz = n * 16; break;
    case TB_BUF_JITTERED_RGBW: src = fb.jittered.data(); sz = n * 16; break;
    case TB_BUF_RESOLVED_RGB:
        tmp.resize(n * 3);
        for (size_t i = 0; i < n; i++) { // PostProcessCS.hlsl:23-27
            tmp[3 * i] = fb.accum[i].x / fb.accum[i].w; tmp[3 * i + 1] = fb.accum[i].y / fb.accum[i].w; tmp[3 * i + 2] = fb.accum[i].z / fb.accum[i].w;
        }
        src = tmp.data(); sz = n * 12; break;
    case TB_BUF_AOV_NORMAL: src = fb.aovNormal.data(); sz = n * 16; break;
    case TB_BUF_AOV_WORLDPOS: {
        uint32_t last = h->shardOffset + (h->samples ? h->samples - 1 : 0) * h->shardStride;
        src = fb.aovWorldPos[last % 2].data(); sz = n * 16; break;
    }
    case TB_BUF_AOV_DEPTH: src = fb.aovDepth.data(); sz = n * 4; break;
    case TB_BUF_AOV_ALBEDO: src = fb.aovAlbedo.data(); sz = n * 16; break;
    case TB_BUF_AOV_EMISSIVE: src = fb.aovEmissive.data(); sz = n * 16; break;
    case TB_BUF_PRIMARY_HIT_IDS: src = fb.primaryHit.data(); sz = n * 8; break;
    case TB_BUF_RAY_COUNTERS: src = fb.counters.data(); sz = n * 8; break;
    default: h->err = "bad buffer kind"; return -1;
    }
    if (bytes < sz) { h->err = "buffer too small"; return -1; }
    memcpy(dst, src, sz);
    return 0;
}
