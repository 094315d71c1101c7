// api.cpp — the C ABI (include/tracerboy_b200.h): host-side mirror of `class TracerBoy`
// (TracerBoy/TracerBoy.h:158-397) and of the fallback layer's build/trace interface
// (D3D12RaytracingFallback.h:76-173) on top of the CUDA kernels. There is no CPU
// fallback: every compute entry point needs a CUDA device and fails with TB_ERR_CUDA
// when none is present.
#include <dlfcn.h>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../cuda/pathtrace.h"
#include "../cuda/postprocess.h"
#include "handle.h"

using namespace tbd;

static std::string g_libDir;
namespace tbh { std::string& create_error() { static std::string e; return e; } }

static std::string lib_dir() {
    if (!g_libDir.empty()) return g_libDir;
    Dl_info info;
    if (dladdr((void*)&lib_dir, &info) && info.dli_fname) {
        std::string p(info.dli_fname);
        size_t s = p.find_last_of('/');
        g_libDir = s == std::string::npos ? "." : p.substr(0, s);
    } else g_libDir = ".";
    return g_libDir;
}

template <class T>
static cudaError_t upload(TbHandle* h, std::vector<void*>& owner, const T* src, size_t count, const T** dst) {
    void* p = nullptr;
    size_t bytes = sizeof(T) * (count ? count : 1);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return e;
    owner.push_back(p);
    if (count) e = cudaMemcpyAsync(p, src, sizeof(T) * count, cudaMemcpyHostToDevice, h->stream);
    *dst = (const T*)p;
    return e;
}

static void free_list(std::vector<void*>& v) { for (void* p : v) cudaFree(p); v.clear(); }

static void free_scene(TbHandle* h) {
    h->options.epoch++; // device pointers baked into captured frame graphs are about to change
    free_list(h->sceneAllocs);
    if (h->bvh.ref) cudaFree(h->bvh.ref);
    if (h->bvh.pairs) cudaFree(h->bvh.pairs);
    if (h->bvh.tris) cudaFree(h->bvh.tris);
    h->bvh = DeviceBvh();
    h->dscene = DeviceScene();
    h->sceneLoaded = false;
}

static void free_frame(TbHandle* h) {
    for (auto& sl : h->slots) {
        if (sl.stream) { cudaStreamSynchronize(sl.stream); cudaStreamDestroy(sl.stream); }
        sl.graph.reset();
        if (sl.frameDone) cudaEventDestroy(sl.frameDone);
        if (sl.accDone) cudaEventDestroy(sl.accDone);
    }
    h->slots.clear();
    tbh::comm_release_buffers(h);
    free_list(h->frameAllocs);
    h->st = PathState();
    h->resolved = nullptr;
    h->post = nullptr; h->post8 = nullptr; h->lumHist = nullptr;
    h->width = h->height = 0;
}

// Halton, RayGenCommon.h:49-60 (per-frame constant, evaluated on the host in IEEE binary32)
static float halton(int b, int i) {
    float r = 0.0f, f = 1.0f;
    while (i > 0) {
        f = f / (float)b;
        r = r + f * (float)(i % b);
        i = (int)floorf((float)i / (float)b);
    }
    return r;
}

static void set_status(TbHandle* h, uint32_t state, uint32_t loaded, uint32_t total) {
    std::lock_guard<std::mutex> g(h->statusLock);
    h->status.State = state; h->status.InstancesLoaded = loaded; h->status.TotalInstances = total;
}

static bool scene_has_sss(const tb::Scene& s) {
    // subsurface / glass materials reachable from the geometry (directly or through a mix material)
    for (const TbGeometryRecord& g : s.geoms) {
        const TbMaterial& m = s.materials[g.MaterialIndex];
        uint32_t ids[3] = {g.MaterialIndex, g.MaterialIndex, g.MaterialIndex};
        if (m.Flags & TB_MIX_MATERIAL_FLAG) { ids[1] = (uint32_t)m.albedo.x; ids[2] = (uint32_t)m.albedo.y; }
        for (uint32_t id : ids)
            if (id < s.materials.size() && (s.materials[id].Flags & TB_SUBSURFACE_SCATTER_MATERIAL_FLAG)) return true;
    }
    return false;
}

// distinct material classes (the key of the shading stage's queue, material_class() in pathtrace.cu) among the
// materials the geometry references
static uint8_t geometry_class(const tb::Scene& s, const TbGeometryRecord& g) {
    const TbMaterial& m = s.materials[g.MaterialIndex];
    return (uint8_t)(((uint32_t)m.Flags & 0x1fu) | (m.albedoIndex != TB_INVALID_TEXTURE ? 0x20u : 0u));
}
static uint32_t scene_material_classes(const tb::Scene& s) {
    uint64_t seen = 0;
    for (const TbGeometryRecord& g : s.geoms) seen |= 1ull << geometry_class(s, g);
    return (uint32_t)__builtin_popcountll(seen);
}
// (re)writes the device table of per-geometry classes; the allocation is made once per scene upload
static int upload_geometry_classes(TbHandle* h) {
    std::vector<uint8_t> cls(h->scene.geoms.size());
    for (size_t g = 0; g < cls.size(); g++) cls[g] = geometry_class(h->scene, h->scene.geoms[g]);
    CUDA_OK(h, cudaMemcpyAsync((void*)h->dscene.geomClass, cls.data(), cls.size(), cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return TB_OK;
}

// One build on caller-described device geometry into caller-provided (or library-owned) memory. Shared by the scene
// path (upload_and_build) and the SW-RT seam (tb_bvh_build_device).
static int run_build(TbHandle* h, const std::vector<BuildGeometry>& descs, const std::vector<uint32_t>& prefix, uint32_t n,
                     uint32_t flags, DeviceBvh& out, void* scratch, cudaStream_t stream) {
    void* owned[3] = {nullptr, nullptr, nullptr};
    struct Guard { void** p; ~Guard() { for (int i = 0; i < 3; i++) if (p[i]) cudaFree(p[i]); } } guard{owned};
    CUDA_OK(h, cudaMalloc(&owned[0], sizeof(BuildGeometry) * descs.size()));
    CUDA_OK(h, cudaMalloc(&owned[1], 4 * prefix.size()));
    CUDA_OK(h, cudaMemcpyAsync(owned[0], descs.data(), sizeof(BuildGeometry) * descs.size(), cudaMemcpyHostToDevice, stream));
    CUDA_OK(h, cudaMemcpyAsync(owned[1], prefix.data(), 4 * prefix.size(), cudaMemcpyHostToDevice, stream));
    if (!scratch) { // no caller scratch: one allocation of the library's own, released after the build
        CUDA_OK(h, cudaMalloc(&owned[2], bvh_scratch_bytes(n)));
        scratch = owned[2];
    }
    int passes = (flags & TB_BVH_BUILD_PREFER_FAST_BUILD) ? 0 : ((flags & TB_BVH_BUILD_PREFER_FAST_TRACE) ? 3 : 1); // TreeletReorder.cpp:63-78
    CUDA_OK(h, build_bvh((const BuildGeometry*)owned[0], (const uint32_t*)owned[1], (uint32_t)descs.size(), n, passes, out, scratch, stream, h->lc));
    if (out.depth > TB_STACK_DEPTH)
        return fail(h, TB_ERR_NOT_IMPL, "the built BVH is " + std::to_string(out.depth) + " levels deep; the traversal stack holds " +
                                         std::to_string(TB_STACK_DEPTH) + " waiting nodes (no ray is ever dropped silently)");
    return TB_OK;
}

// Uploads h->scene and builds the BVH.
static int upload_and_build(TbHandle* h, uint32_t flags) {
    tb::Scene& s = h->scene;
    CUDA_OK(h, cudaSetDevice(h->device));
    free_scene(h); // whatever happens below, the previous device scene is gone (sceneLoaded = false)
    if (s.numTriangles() == 0) return fail(h, TB_ERR_INVALID_ARG, "scene has no triangles");
    if (s.numTriangles() > tb_max_triangles()) return fail(h, TB_ERR_INVALID_ARG, "too many triangles (the reference BVH layout has 32-bit byte offsets: 116 N - 16 must stay below 4 GiB)");
    std::string why;
    if (!tb::validate_scene(s, why)) return fail(h, TB_ERR_INVALID_ARG, "invalid scene: " + why);
    set_status(h, TB_RECORDING_DEVICE_WORK, (uint32_t)s.geoms.size(), (uint32_t)s.geoms.size());
    DeviceScene& d = h->dscene;
    CUDA_OK(h, upload(h, h->sceneAllocs, s.geoms.data(), s.geoms.size(), &d.geoms));
    CUDA_OK(h, upload(h, h->sceneAllocs, (const float*)s.positions.data(), s.positions.size() * 3, &d.positions));
    CUDA_OK(h, upload(h, h->sceneAllocs, s.vertices.data(), s.vertices.size(), &d.vertices));
    CUDA_OK(h, upload(h, h->sceneAllocs, s.indices.data(), s.indices.size(), &d.indices));
    CUDA_OK(h, upload(h, h->sceneAllocs, s.materials.data(), s.materials.size(), &d.materials));
    CUDA_OK(h, upload(h, h->sceneAllocs, s.lights.data(), s.lights.size(), &d.lights));
    CUDA_OK(h, upload(h, h->sceneAllocs, s.textures.data(), s.textures.size(), &d.textures));
    std::vector<DeviceScene::ImageRef> refs(s.images.size());
    for (size_t i = 0; i < s.images.size(); i++) {
        const uint8_t* p = nullptr;
        CUDA_OK(h, upload(h, h->sceneAllocs, s.images[i].data.data(), s.images[i].data.size(), &p));
        refs[i] = {p, s.images[i].width, s.images[i].height, s.images[i].format};
    }
    CUDA_OK(h, upload(h, h->sceneAllocs, refs.data(), refs.size(), &d.images));
    CUDA_OK(h, upload(h, h->sceneAllocs, h->blueNoiseHost.data(), h->blueNoiseHost.size(), &d.blueNoise));
    {
        void* p = nullptr;
        CUDA_OK(h, cudaMalloc(&p, s.geoms.size() ? s.geoms.size() : 1));
        h->sceneAllocs.push_back(p);
        d.geomClass = (const uint8_t*)p;
        int rc = upload_geometry_classes(h);
        if (rc != TB_OK) return rc;
    }
    d.numGeoms = (uint32_t)s.geoms.size(); d.numMaterials = (uint32_t)s.materials.size();
    d.numLights = (uint32_t)s.lights.size(); d.numTextures = (uint32_t)s.textures.size(); d.numImages = (uint32_t)s.images.size();
    d.envImage = s.envImage; d.flipTextureUVs = s.flipTextureUVs;
    for (int r = 0; r < 3; r++) { d.envTransform[r][0] = s.envTransform[r].x; d.envTransform[r][1] = s.envTransform[r].y; d.envTransform[r][2] = s.envTransform[r].z; d.envTransform[r][3] = s.envTransform[r].w; }
    d.envColorScale[0] = s.envColorScale.x; d.envColorScale[1] = s.envColorScale.y; d.envColorScale[2] = s.envColorScale.z;
    // per-geometry triangle prefix (LoadPrimitivesPass.cpp:70-166 walks the descs in order); the pooled arrays are
    // described to the builder the way a D3D12 caller describes its vertex / index buffers
    std::vector<uint32_t> prefix(s.geoms.size());
    std::vector<BuildGeometry> descs(s.geoms.size());
    uint32_t n = 0;
    for (size_t g = 0; g < s.geoms.size(); g++) {
        const TbGeometryRecord& G = s.geoms[g];
        prefix[g] = n; n += G.IndexCount / 3;
        descs[g] = {(const uint8_t*)(d.positions + 3 * (size_t)G.VertexFirst), d.indices + G.IndexFirst, nullptr, 12u, 4u, G.GeometryFlags, 0u};
    }
    h->bvh.refBytes = bvh_ref_bytes(n);
    CUDA_OK(h, cudaMalloc((void**)&h->bvh.ref, h->bvh.refBytes));
    CUDA_OK(h, cudaMalloc((void**)&h->bvh.pairs, sizeof(PairNode) * (size_t)(n > 1 ? n - 1 : 1)));
    CUDA_OK(h, cudaMalloc((void**)&h->bvh.tris, sizeof(WideTri) * (size_t)n));
    // the builder's scratch is kept between builds of the same handle (reloading a scene of the same size reuses it)
    const uint64_t need = bvh_scratch_bytes(n);
    if (h->buildScratchBytes < need) {
        if (h->buildScratch) cudaFree(h->buildScratch);
        h->buildScratch = nullptr; h->buildScratchBytes = 0;
        CUDA_OK(h, cudaMalloc(&h->buildScratch, need));
        h->buildScratchBytes = need;
    }
    set_status(h, TB_WAITING_ON_GPU, (uint32_t)s.geoms.size(), (uint32_t)s.geoms.size());
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    CUDA_OK(h, cudaEventRecord(h->ev0, h->stream));
    int rc = run_build(h, descs, prefix, n, flags, h->bvh, h->buildScratch, h->stream);
    if (rc != TB_OK) return rc;
    CUDA_OK(h, cudaEventRecord(h->ev1, h->stream));
    CUDA_OK(h, cudaEventSynchronize(h->ev1));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    h->bvhBuildMs = ms;
    h->camera = s.camera;
    h->options.sceneHasSSS = scene_has_sss(s);
    h->options.sceneMaterialClasses = scene_material_classes(s);
    h->sceneLoaded = true;
    h->samplesRendered = 0;
    set_status(h, TB_LOAD_FINISHED, (uint32_t)s.geoms.size(), (uint32_t)s.geoms.size());
    return TB_OK;
}

extern "C" {

// Largest triangle count one acceleration structure can hold: the reference layout's header and leaf records carry
// 32-bit byte offsets (RayTracingHlslCompat.h:344-398), so 116 N - 16 must stay below 4 GiB (and 2N - 1 nodes below
// the 30-bit node index, which is the weaker bound).
TB_API uint64_t tb_max_triangles(void) { return ((1ull << 32) - 1 + 16) / 116; }

TB_API const char* tb_version(void) { return "tracerboy_b200 0.1 (sm_100a)"; }

TB_API const char* tb_last_error(TbHandle* h) { return h ? h->err.c_str() : tbh::create_error().c_str(); }

TB_API int tb_create(int device, TbHandle** out) {
    if (!out) return fail(nullptr, TB_ERR_INVALID_ARG, "out is null");
    *out = nullptr;
    // Frames in flight run on independent streams; the default of 8 hardware queues would
    // serialise them. Only effective if the CUDA context of this process does not exist yet.
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, TB_ERR_CUDA, std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, TB_ERR_INVALID_ARG, "device ordinal out of range");
    TbHandle* h = new TbHandle();
    h->device = device;
    auto abandon = [&]() {
        if (h->ev0) cudaEventDestroy(h->ev0);
        if (h->ev1) cudaEventDestroy(h->ev1);
        if (h->stream) cudaStreamDestroy(h->stream);
        delete h;
    };
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess) {
        abandon();
        return fail(nullptr, TB_ERR_CUDA, "cannot initialise CUDA stream/events");
    }
    if (cudaDeviceGetAttribute(&h->numSMs, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || h->numSMs <= 0) h->numSMs = 148;
    h->options.numSMs = h->numSMs;
    // The builder's temporaries come from the device's default stream-ordered pool; keep up to 16 GiB of it
    // mapped between builds instead of handing it back to the driver at every synchronise.
    {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = 16ull << 30;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    // blue noise (TracerBoy.cpp:2126-2134): LDR_RGBA_0/1 decoded to raw RGBA8
    h->blueNoiseHost.assign(2 * 256 * 256 * 4, 0);
    std::string bn = lib_dir() + "/../data/bluenoise_rgba8_256.bin";
    FILE* f = fopen(bn.c_str(), "rb");
    if (!f || fread(h->blueNoiseHost.data(), 1, h->blueNoiseHost.size(), f) != h->blueNoiseHost.size()) {
        if (f) fclose(f);
        abandon();
        return fail(nullptr, TB_ERR_IO, "cannot read blue-noise table " + bn);
    }
    fclose(f);
    *out = h;
    return TB_OK;
}

TB_API void tb_destroy(TbHandle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    tbh::comm_destroy(h);
    free_frame(h);
    free_scene(h);
    if (h->buildScratch) cudaFree(h->buildScratch);
    cudaEventDestroy(h->ev0);
    cudaEventDestroy(h->ev1);
    cudaStreamDestroy(h->stream);
    delete h;
}

static int load_host_scene(TbHandle* h, tb::Scene& s, const char* path) {
    std::string p(path), err;
    auto ends = [&](const char* suf) { size_t n = strlen(suf); return p.size() >= n && p.compare(p.size() - n, n, suf) == 0; };
    if (p.rfind("synthetic:", 0) == 0) {
        if (!tb::make_synthetic(s, p, err)) return fail(h, TB_ERR_INVALID_ARG, err);
    } else if (ends(".tbscene")) {
        if (!tb::load_tbscene(s, p, err)) return fail(h, TB_ERR_IO, err);
    } else if (ends(".pbrt") || ends(".pbf")) {
        // optional importer built from the reference's vendored pbrt-parser
        std::string so = lib_dir() + "/libtb_pbrtimport.so";
        void* lib = dlopen(so.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!lib) return fail(h, TB_ERR_NOT_IMPL, "PBRT import needs " + so + " (built only when the pbrt-parser sources are available)");
        typedef int (*ImportFn)(const char*, uint32_t, void*, char*, size_t);
        ImportFn fn = (ImportFn)dlsym(lib, "tb_pbrt_import_ex");
        char buf[1024] = {0};
        if (!fn || fn(path, h->instanceMode, &s, buf, sizeof(buf)) != 0) return fail(h, TB_ERR_IO, std::string("pbrt import failed: ") + buf);
    } else {
        // AssimpImporter (fbx/obj/...) links a Windows-only binary in the reference; not available
        return fail(h, TB_ERR_NOT_IMPL, "unsupported scene type: " + p);
    }
    return TB_OK;
}

TB_API int tb_load_scene_ex(TbHandle* h, const char* path, uint32_t flags) {
    if (!h || !path) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    set_status(h, TB_LOADING_PBRT, 0, 0);
    tb::Scene incoming; // a failed import leaves the handle's current scene (host mirror and device copy) untouched
    int rc = load_host_scene(h, incoming, path);
    if (rc != TB_OK) { set_status(h, TB_LOAD_FAILED, 0, 0); return rc; }
    h->scene = std::move(incoming);
    set_status(h, TB_LOADING_HOST, 0, (uint32_t)h->scene.geoms.size());
    rc = upload_and_build(h, flags);
    if (rc != TB_OK) set_status(h, TB_LOAD_FAILED, 0, 0);
    return rc;
}
TB_API int tb_load_scene(TbHandle* h, const char* path) { return tb_load_scene_ex(h, path, TB_BVH_BUILD_PREFER_FAST_TRACE); }

TB_API int tb_set_instance_mode(TbHandle* h, uint32_t mode) {
    if (!h) return TB_ERR_INVALID_ARG;
    if (mode > TB_INSTANCES_INSERT_INTO_BLAS) return fail(h, TB_ERR_INVALID_ARG, "instance mode must be TB_INSTANCES_SKIP or TB_INSTANCES_INSERT_INTO_BLAS");
    h->instanceMode = mode;
    return TB_OK;
}

TB_API int tb_convert_scene(const char* inPath, const char* outTbscene, char* err, size_t errCap) {
    return tb_convert_scene_ex(inPath, outTbscene, TB_INSTANCES_SKIP, err, errCap);
}
TB_API int tb_convert_scene_ex(const char* inPath, const char* outTbscene, uint32_t instanceMode, char* err, size_t errCap) {
    TbHandle tmp; // host-only use
    tmp.instanceMode = instanceMode;
    tb::Scene s;
    int rc = load_host_scene(&tmp, s, inPath);
    std::string e = tmp.err;
    if (rc == TB_OK && !tb::save_tbscene(s, outTbscene, e)) rc = TB_ERR_IO;
    if (rc != TB_OK && err && errCap) { strncpy(err, e.c_str(), errCap - 1); err[errCap - 1] = 0; }
    return rc;
}

TB_API int tb_get_load_status(TbHandle* h, TbSceneLoadStatus* out) {
    if (!h || !out) return TB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(h->statusLock);
    *out = h->status;
    return TB_OK;
}

TB_API int tb_save_scene(TbHandle* h, const char* path) {
    if (!h || !path) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    if (!h->sceneLoaded) return fail(h, TB_ERR_STATE, "no scene loaded");
    std::string err;
    if (!tb::save_tbscene(h->scene, path, err)) return fail(h, TB_ERR_IO, err);
    return TB_OK;
}

TB_API int tb_get_scene_info(TbHandle* h, TbSceneInfo* o) {
    if (!h || !o) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    if (!h->sceneLoaded) return fail(h, TB_ERR_STATE, "no scene loaded");
    o->NumGeometries = (uint32_t)h->scene.geoms.size(); o->NumTriangles = h->scene.numTriangles();
    o->NumVertices = (uint32_t)h->scene.positions.size(); o->NumMaterials = (uint32_t)h->scene.materials.size();
    o->NumLights = (uint32_t)h->scene.lights.size(); o->NumTextures = (uint32_t)h->scene.textures.size();
    o->NumImages = (uint32_t)h->scene.images.size(); o->HasEnvironmentMap = h->scene.envImage >= 0;
    return TB_OK;
}

TB_API int tb_get_bvh_size(TbHandle* h, uint64_t* bytes) {
    if (!h || !bytes) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    if (!h->sceneLoaded) return fail(h, TB_ERR_STATE, "no scene loaded");
    *bytes = h->bvh.refBytes;
    return TB_OK;
}
TB_API int tb_get_bvh(TbHandle* h, void* dst, uint64_t bytes) {
    if (!h || !dst) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    if (!h->sceneLoaded) return fail(h, TB_ERR_STATE, "no scene loaded");
    if (bytes < h->bvh.refBytes) return fail(h, TB_ERR_INVALID_ARG, "destination too small");
    CUDA_OK(h, cudaSetDevice(h->device));
    CUDA_OK(h, cudaMemcpyAsync(dst, h->bvh.ref, h->bvh.refBytes, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return TB_OK;
}
TB_API int tb_get_bvh_build_ms(TbHandle* h, double* ms) {
    if (!h || !ms) return TB_ERR_INVALID_ARG;
    *ms = h->bvhBuildMs;
    return TB_OK;
}

TB_API int tb_get_default_settings(TbOutputSettings* s) {
    if (!s) return TB_ERR_INVALID_ARG;
    memset(s, 0, sizeof(*s));
    s->OutputType = TB_OUTPUT_LIT; s->EnableNormalMaps = 0; s->RenderMode = TB_RENDER_UNBIASED;
    s->SampleLimit = 0; s->TimeLimitInSeconds = 0.0f; s->DebugValue = 1.0f; s->DebugValue2 = 1.0f;
    s->DOFFocalDistance = 0.0f; s->ApertureWidth = 0.075f; s->FilterType = TB_FILTER_BOX; s->FilterWidth = 1.0f;
    s->FireflyClampValue = 0.0f; s->MaxZ = 10000.0f; s->ConvergencePercentage = 0.001f;
    s->EnableNextEventEstimation = 1; s->EnableSamplingImportanceResampling = 0; s->EnableBlueNoise = 1; s->MaxBounces = 6;
    return TB_OK;
}

TB_API int tb_get_camera(TbHandle* h, TbCamera* c) {
    if (!h || !c) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    *c = h->camera;
    return TB_OK;
}
TB_API int tb_set_camera(TbHandle* h, const TbCamera* c) {
    if (!h || !c) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    h->camera = *c;
    h->samplesRendered = 0; // m_bInvalidateHistory
    return TB_OK;
}

// ---- TracerBoy::Update (TracerBoy.cpp:3386-3500), restated on plain float3 math.
namespace {
struct V3 { float x, y, z; };
inline V3 vadd(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 vsub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 vmul(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float vdot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 vcross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline V3 vnorm(V3 a) { float l = std::sqrt(vdot(a, a)); return l > 0.0f ? vmul(a, 1.0f / l) : a; } // XMVector3Normalize
// v * XMMatrixRotationAxis(axis, angle): row-vector product with the axis-angle matrix, i.e. Rodrigues' formula
inline V3 rotate_about(V3 v, V3 axis, float angle) {
    V3 n = vnorm(axis);
    float s = std::sin(angle), c = std::cos(angle);
    return vadd(vadd(vmul(v, c), vmul(vcross(n, v), s)), vmul(n, vdot(n, v) * (1.0f - c)));
}
} // namespace

extern "C" TB_API int tb_camera_update(TbCamera* cam, uint32_t lastMouse[2], uint32_t width, uint32_t height, int mouseX, int mouseY,
                                       const uint8_t* keys, float dt, const TbControllerState* controller,
                                       const TbCameraSettings* settings, int* movedOut) {
    if (!cam || !lastMouse) return TB_ERR_INVALID_ARG;
    const TbControllerState noController = {0, 0, 0, 0, 0, 0};
    const TbCameraSettings defaults = {1.0f, 0u};
    const TbControllerState& cs = controller ? *controller : noController;
    const TbCameraSettings& set = settings ? *settings : defaults;
    auto key = [&](int lower, int upper) { return keys && (keys[lower] || keys[upper]); };
    bool moved = false;
    float yaw = 0.0f, pitch = 0.0f;
    if (width && height && !set.IgnoreMouse) { // :3393-3399 (the products are evaluated in double there: 2.0 is a double literal)
        const float rotationScaler = 0.5f;
        yaw = (float)(rotationScaler * 2.0 * 6.28f * ((float)mouseX - (float)lastMouse[0]) / (float)width);
        pitch = rotationScaler * 3.14f * ((float)mouseY - (float)lastMouse[1]) / (float)height;
    }
    const float deadzone = 0.2f; // :3401
    if (std::fabs(cs.RightStickX) > deadzone || std::fabs(cs.RightStickY) > deadzone) {
        const float rotationScaler = 0.001f;
        yaw += cs.RightStickX * rotationScaler * dt;
        pitch += -cs.RightStickY * rotationScaler * dt;
        moved = true;
    }
    if (lastMouse[0] != (uint32_t)mouseX || lastMouse[1] != (uint32_t)mouseY) { // :3410-3418
        if (!set.IgnoreMouse) moved = true;
        lastMouse[0] = (uint32_t)mouseX;
        lastMouse[1] = (uint32_t)mouseY;
    }
    V3 right = {cam->Right.x, cam->Right.y, cam->Right.z};
    V3 position = {cam->Position.x, cam->Position.y, cam->Position.z};
    V3 lookAt = {cam->LookAt.x, cam->LookAt.y, cam->LookAt.z};
    V3 viewDir = vsub(lookAt, position);
    const V3 globalUp = {0.0f, 1.0f, 0.0f};
    const V3 xzRight = vnorm(V3{right.x, 0.0f, right.z});
    // RotationAxis(GlobalUp, yaw) * RotationAxis(XZAlignedRight, pitch), row vectors: yaw first, then pitch (:3431-3432)
    viewDir = vnorm(rotate_about(rotate_about(viewDir, globalUp, yaw), xzRight, pitch));
    right = vnorm(vcross(globalUp, viewDir));
    V3 up = vnorm(vcross(viewDir, right));
    lookAt = vadd(position, viewDir);
    const float speed = set.MovementSpeed;
    // Position += dt * speed * Axis * multiplier: the scalar dt * speed first, then the vector, then the multiplier
    auto move = [&](V3 dir, float multiplier, bool add) {
        const V3 d = vmul(vmul(dir, dt * speed), multiplier);
        position = add ? vadd(position, d) : vsub(position, d);
        lookAt = add ? vadd(lookAt, d) : vsub(lookAt, d);
        moved = true;
    };
    const bool lsy = std::fabs(cs.LeftStickY) > deadzone, lsx = std::fabs(cs.LeftStickX) > deadzone;
    const bool rtr = cs.RightTrigger > deadzone, ltr = cs.LeftTrigger > deadzone;
    if (key('w', 'W') || lsy) move(viewDir, lsy ? cs.LeftStickY : 1.0f, true);
    if (key('s', 'S')) move(viewDir, 1.0f, false);
    if (key('a', 'A')) move(right, 1.0f, false);
    if (key('d', 'D') || lsx) move(right, lsx ? cs.LeftStickX : 1.0f, true);
    if (key('q', 'Q') || rtr) move(up, rtr ? cs.RightTrigger : 1.0f, true);
    if (key('e', 'E') || ltr) move(up, ltr ? cs.LeftTrigger : 1.0f, false);
    if (moved) { // :3491-3498: the rotated frame is only stored when something moved
        cam->Position = {position.x, position.y, position.z};
        cam->LookAt = {lookAt.x, lookAt.y, lookAt.z};
        cam->Right = {right.x, right.y, right.z};
        cam->Up = {up.x, up.y, up.z};
    }
    if (movedOut) *movedOut = moved ? 1 : 0;
    return TB_OK;
}

TB_API int tb_update(TbHandle* h, int mouseX, int mouseY, const uint8_t* keys, float dt, const TbControllerState* controller,
                     const TbCameraSettings* settings) {
    if (!h) return TB_ERR_INVALID_ARG;
    int moved = 0;
    int rc = tb_camera_update(&h->camera, h->lastMouse, h->width, h->height, mouseX, mouseY, keys, dt, controller, settings, &moved);
    if (rc != TB_OK) return fail(h, rc, "tb_update: bad argument");
    if (moved) h->samplesRendered = 0; // InvalidateHistory()
    return TB_OK;
}

TB_API int tb_resize(TbHandle* h, uint32_t w, uint32_t hh) {
    if (!h || w == 0 || hh == 0 || (uint64_t)w * hh > (1ull << 26)) return fail(h, TB_ERR_INVALID_ARG, "bad resolution (at most 2^26 pixels: queue entries keep 6 bits for the material class)");
    CUDA_OK(h, cudaSetDevice(h->device));
    if (w == h->width && hh == h->height) return TB_OK;
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    free_frame(h);
    size_t n = (size_t)w * hh;
    auto alloc = [&](void** p, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(p, bytes);
        if (e == cudaSuccess) { h->frameAllocs.push_back(*p); e = cudaMemsetAsync(*p, 0, bytes, h->stream); }
        return e;
    };
    PathState& st = h->st;
    CUDA_OK(h, alloc((void**)&st.accum, 16 * n)); CUDA_OK(h, alloc((void**)&st.jittered, 16 * n));
    CUDA_OK(h, alloc((void**)&st.aovAlbedo, 16 * n)); CUDA_OK(h, alloc((void**)&st.aovNormal, 16 * n));
    CUDA_OK(h, alloc((void**)&st.aovEmissive, 16 * n));
    CUDA_OK(h, alloc((void**)&st.aovWorldPos[0], 16 * n)); CUDA_OK(h, alloc((void**)&st.aovWorldPos[1], 16 * n));
    CUDA_OK(h, alloc((void**)&st.aovDepth, 4 * n));
    CUDA_OK(h, alloc((void**)&st.primaryHit, 8 * n)); CUDA_OK(h, alloc((void**)&st.counters, 8 * n));
    CUDA_OK(h, alloc((void**)&st.stats, 8 * TB_STATS_WORDS)); CUDA_OK(h, alloc((void**)&st.readbackStats, sizeof(TbReadbackStats)));
    {
        // private state per slot: 5 float4 + hitGeom + 2 float4 + 2 queues + staging (float4+float+float4+float)
        // + walk state + 2 suspension buffers. Automatic policy: as many slots as fit a third of the free device
        // memory (at most 64 GB), between 4 and 32. Measured (profiles/README.md): 16 instead of 8 slots is +1 %
        // on Teapot, +12 % on the dragon stand-in and +22 % on vw-van at 4K, whose glass walks end in a
        // millisecond-long tail of a few sequential 100-ray walkers that only other frames' work can hide;
        // 32 instead of 16 is another +6 % on vw-van and on the 20 M-triangle scene, +1.5 % on the dragon stand-in.
        size_t perSlot = n * (80 + 4 + 32 + 8 + 40 + 56 + 8 + 40 + 8) + 2 * (n / 16 + 1024) * 448;
        uint32_t fif = h->framesInFlight;
        if (fif == 0) {
            size_t freeB = 0, totalB = 0;
            size_t budget = (size_t)64 << 30;
            if (cudaMemGetInfo(&freeB, &totalB) == cudaSuccess && freeB / 3 < budget) budget = freeB / 3;
            size_t fit = budget / perSlot;
            fif = (uint32_t)(fit < 4 ? 4 : (fit > 32 ? 32 : fit));
        }
        h->slots.resize(fif);
    }
    for (auto& sl : h->slots) {
        sl.st = st; // shared pointers
        PathState& p = sl.st;
        CUDA_OK(h, alloc((void**)&p.rayO, 16 * n)); CUDA_OK(h, alloc((void**)&p.rayD, 16 * n));
        CUDA_OK(h, alloc((void**)&p.thr, 16 * n)); CUDA_OK(h, alloc((void**)&p.col, 16 * n));
        CUDA_OK(h, alloc((void**)&p.hit, 16 * n)); CUDA_OK(h, alloc((void**)&p.hitGeom, 4 * n));
        CUDA_OK(h, alloc((void**)&p.neighbor, 16 * n)); CUDA_OK(h, alloc((void**)&p.neighborDir, 16 * n));
        CUDA_OK(h, alloc((void**)&p.queue[0], 4 * n)); CUDA_OK(h, alloc((void**)&p.queue[1], 4 * n));
        CUDA_OK(h, alloc((void**)&p.queueCount, 4 * TB_QUEUE_COUNT_WORDS));
        CUDA_OK(h, alloc((void**)&p.hitQueue, 4 * n)); CUDA_OK(h, alloc((void**)&p.hitSorted, 4 * n)); CUDA_OK(h, alloc((void**)&p.missQueue, 4 * n));
        CUDA_OK(h, alloc((void**)&p.sortKeys, 4 * n)); CUDA_OK(h, alloc((void**)&p.sortTmp, 4 * n)); CUDA_OK(h, alloc((void**)&p.sortHist, 4 * (TB_SORT_CELLS + 1)));
        CUDA_OK(h, alloc((void**)&p.shadowQueue, 4 * n)); CUDA_OK(h, alloc((void**)&p.shRayO, 16 * n)); CUDA_OK(h, alloc((void**)&p.shRayD, 16 * n));
        CUDA_OK(h, alloc((void**)&p.shHit, 16 * n)); CUDA_OK(h, alloc((void**)&p.shHitGeom, 4 * n));
        CUDA_OK(h, alloc((void**)&p.walkQueue[0], 4 * n)); CUDA_OK(h, alloc((void**)&p.walkQueue[1], 4 * n)); CUDA_OK(h, alloc((void**)&p.walkA, 16 * n)); CUDA_OK(h, alloc((void**)&p.walkB, 16 * n));
        p.susCapacity = (uint32_t)(n / 16 + 1024);
        CUDA_OK(h, alloc((void**)&p.susBuf[0], (size_t)p.susCapacity * 448)); CUDA_OK(h, alloc((void**)&p.susBuf[1], (size_t)p.susCapacity * 448));
        CUDA_OK(h, alloc((void**)&p.susCount, 16));
        CUDA_OK(h, alloc((void**)&p.sample, 16 * n)); CUDA_OK(h, alloc((void**)&p.sampleSeed, 4 * n));
        CUDA_OK(h, alloc((void**)&p.stEmissive, 16 * n)); CUDA_OK(h, alloc((void**)&p.stDepth, 4 * n));
        CUDA_OK(h, alloc((void**)&sl.fcDev, sizeof(FrameConstants)));
        CUDA_OK(h, cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        CUDA_OK(h, cudaEventCreateWithFlags(&sl.frameDone, cudaEventDisableTiming));
        CUDA_OK(h, cudaEventCreateWithFlags(&sl.accDone, cudaEventDisableTiming));
    }
    CUDA_OK(h, alloc((void**)&h->resolved, 12 * n));
    CUDA_OK(h, alloc((void**)&h->post, 16 * n)); CUDA_OK(h, alloc((void**)&h->post8, 4 * n));
    CUDA_OK(h, alloc((void**)&h->lumHist, 257 * sizeof(uint32_t)));
    CUDA_OK(h, cudaStreamSynchronize(h->stream)); // memsets done before the slot streams touch the buffers
    h->width = w; h->height = hh;
    h->samplesRendered = 0;
    return TB_OK;
}

TB_API int tb_select_pixel(TbHandle* h, int x, int y) {
    if (!h) return TB_ERR_INVALID_ARG;
    h->selX = x; h->selY = y;
    return TB_OK;
}

TB_API int tb_get_stats(TbHandle* h, TbReadbackStats* out) {
    if (!h || !out) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    if (!h->st.readbackStats) return fail(h, TB_ERR_STATE, "no frame buffers");
    CUDA_OK(h, cudaSetDevice(h->device));
    CUDA_OK(h, cudaMemcpyAsync(out, h->st.readbackStats, sizeof(*out), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    out->ActivePixels = h->width * h->height;
    out->ActiveWaves = (h->width * h->height + 63) / 64;
    return TB_OK;
}

TB_API int tb_render(TbHandle* h, const TbOutputSettings* s, uint32_t nSamples, float time) {
    if (!h || !s) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    if (!h->sceneLoaded) return fail(h, TB_ERR_STATE, "tb_render before tb_load_scene");
    if (!h->width) return fail(h, TB_ERR_STATE, "tb_render before tb_resize");
    if (s->MaxBounces < 0 || s->MaxBounces > 255) return fail(h, TB_ERR_INVALID_ARG, "MaxBounces must be in [0,255]");
    CUDA_OK(h, cudaSetDevice(h->device));
    if (h->samplesRendered == 0) { h->renderStart = std::chrono::steady_clock::now(); }
    // sample / time limits, TracerBoy.cpp:2679-2682. With a time limit the frames are issued one at
    // a time (the last frame is not known in advance and it must own the AOVs).
    uint32_t todo = nSamples;
    if (s->SampleLimit > 0) {
        uint32_t left = (int)h->samplesRendered >= s->SampleLimit ? 0u : (uint32_t)s->SampleLimit - h->samplesRendered;
        if (todo > left) todo = left;
    }
    if (h->comm && todo) h->comm->valid = false; // the job-wide image is stale from here on
    const bool timeLimited = s->TimeLimitInSeconds > 0.0f;
    const bool serial = timeLimited || h->profiling == 1;
    CUDA_OK(h, cudaEventRecord(h->ev0, h->stream));
    for (auto& sl : h->slots) CUDA_OK(h, cudaStreamWaitEvent(sl.stream, h->ev0, 0)); // slot streams start after ev0
    h->options.suspendRays = serial || h->slots.size() < 8;
    for (uint32_t i = 0; i < todo; i++) {
        if (timeLimited &&
            std::chrono::duration<float>(std::chrono::steady_clock::now() - h->renderStart).count() >= s->TimeLimitInSeconds) break;
        FrameConstants fc;
        memset(&fc, 0, sizeof(fc));
        fc.settings = *s;
        fc.camera = h->camera;
        fc.time = time;
        fc.frame = h->shardOffset + h->samplesRendered * h->shardStride; // GlobalFrameCount
        fc.width = h->width; fc.height = h->height;
        fc.selectedX = h->selX; fc.selectedY = h->selY;
        fc.halton2 = halton(2, (int)fc.frame);
        fc.halton3 = halton(3, (int)fc.frame);
        fc.clearAccum = h->samplesRendered == 0;
        fc.rowOffset = h->rowOffset; fc.rowStride = h->rowStride;
        const bool last = serial || i + 1 == todo;
        fc.aovMask = last ? 3u : (i + 2 == todo ? 2u : 0u);
        // world-position ping-pong by the LOCAL sample index: under sample sharding the global frame index advances
        // by the stride, whose parity may never change (the last two frames would then share a buffer)
        fc.worldPosSlot = h->samplesRendered & 1u;
        TbHandle::Slot& sl = h->slots[serial ? 0 : h->framesIssued % h->slots.size()];
        // the slot's previous frame must have been consumed by its k_accumulate
        CUDA_OK(h, cudaStreamWaitEvent(sl.stream, sl.accDone, 0));
        CUDA_OK(h, render_frame(h->bvh, h->dscene, fc, sl.fcDev, sl.st, sl.stream, h->lc, h->profiling ? &sl.timers : nullptr, h->options,
                                serial ? nullptr : &sl.graph));
        CUDA_OK(h, cudaEventRecord(sl.frameDone, sl.stream));
        CUDA_OK(h, cudaStreamWaitEvent(h->stream, sl.frameDone, 0));
        CUDA_OK(h, accumulate_frame(fc, sl.st, h->stream, h->lc)); // frame order == issue order on h->stream
        CUDA_OK(h, cudaEventRecord(sl.accDone, h->stream));
        h->framesIssued++;
        h->samplesRendered++;
        h->pathsStarted += (uint64_t)h->width * h->height;
        if (h->profiling && sl.timers.used > 4096) { // bound the number of live events
            CUDA_OK(h, cudaStreamSynchronize(h->stream));
            sl.timers.resolve(h->extendMs, h->shadeMs, h->resumeMs, h->extendLaunches, h->bounceMs, h->bounceExtendMs);
        }
    }
    CUDA_OK(h, cudaEventRecord(h->ev1, h->stream));
    CUDA_OK(h, cudaEventSynchronize(h->ev1));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    h->deviceMs += ms;
    if (h->profiling) {
        for (auto& sl : h->slots) { CUDA_OK(h, cudaStreamSynchronize(sl.stream)); sl.timers.resolve(h->extendMs, h->shadeMs, h->resumeMs, h->extendLaunches, h->bounceMs, h->bounceExtendMs); }
    }
    CUDA_OK(h, cudaGetLastError());
    return TB_OK;
}

TB_API int tb_samples_rendered(TbHandle* h, uint32_t* out) {
    if (!h || !out) return TB_ERR_INVALID_ARG;
    *out = h->samplesRendered;
    return TB_OK;
}
TB_API int tb_invalidate_history(TbHandle* h) {
    if (!h) return TB_ERR_INVALID_ARG;
    h->samplesRendered = 0;
    return TB_OK;
}
TB_API int tb_set_frame_shard(TbHandle* h, uint32_t offset, uint32_t stride) {
    if (!h || stride == 0 || offset >= stride) return fail(h, TB_ERR_INVALID_ARG, "need offset < stride");
    h->shardOffset = offset; h->shardStride = stride;
    h->samplesRendered = 0;
    if (h->comm) h->comm->valid = false;
    return TB_OK;
}

TB_API int tb_set_row_shard(TbHandle* h, uint32_t offset, uint32_t stride) {
    if (!h || stride == 0 || offset >= stride) return fail(h, TB_ERR_INVALID_ARG, "need offset < stride");
    h->rowOffset = offset; h->rowStride = stride;
    h->samplesRendered = 0;
    if (h->comm) h->comm->valid = false;
    if (h->width) { // pixels of other shards must read as zero
        CUDA_OK(h, cudaSetDevice(h->device));
        size_t n = (size_t)h->width * h->height;
        CUDA_OK(h, cudaMemsetAsync(h->st.accum, 0, 16 * n, h->stream));
        CUDA_OK(h, cudaMemsetAsync(h->st.jittered, 0, 16 * n, h->stream));
        CUDA_OK(h, cudaStreamSynchronize(h->stream));
    }
    return TB_OK;
}

// With a communicator the accumulation buffers a caller sees are the job-wide ones; the reduction (collective) runs
// here when frames were rendered since the last one.
static int job_wide_current(TbHandle* h) {
    if (h->comm && !h->comm->valid) return tb_comm_reduce(h);
    return TB_OK;
}
static const float4* accum_view(TbHandle* h) { return h->comm ? h->comm->reducedAccum : h->st.accum; }

static int buffer_info(TbHandle* h, uint32_t kind, void** p, uint64_t* bytes) {
    size_t n = (size_t)h->width * h->height;
    PathState& st = h->st;
    switch (kind) {
    case TB_BUF_ACCUM_RGBW: *p = h->comm ? h->comm->reducedAccum : st.accum; *bytes = 16 * n; break;
    case TB_BUF_JITTERED_RGBW: *p = h->comm ? h->comm->reducedJittered : st.jittered; *bytes = 16 * n; break;
    case TB_BUF_RESOLVED_RGB: *p = h->resolved; *bytes = 12 * n; break;
    case TB_BUF_AOV_NORMAL: *p = st.aovNormal; *bytes = 16 * n; break;
    case TB_BUF_AOV_WORLDPOS: {
        uint32_t last = h->samplesRendered ? h->samplesRendered - 1 : 0;
        *p = st.aovWorldPos[last & 1]; *bytes = 16 * n; break;
    }
    case TB_BUF_AOV_DEPTH: *p = st.aovDepth; *bytes = 4 * n; break;
    case TB_BUF_AOV_ALBEDO: *p = st.aovAlbedo; *bytes = 16 * n; break;
    case TB_BUF_AOV_EMISSIVE: *p = st.aovEmissive; *bytes = 16 * n; break;
    case TB_BUF_PRIMARY_HIT_IDS: *p = st.primaryHit; *bytes = 8 * n; break;
    case TB_BUF_RAY_COUNTERS: *p = st.counters; *bytes = 8 * n; break;
    case TB_BUF_POSTPROCESS_RGBA: *p = h->post; *bytes = 16 * n; break;
    case TB_BUF_BACKBUFFER_RGBA8: *p = h->post8; *bytes = 4 * n; break;
    case TB_BUF_LUMINANCE_HISTOGRAM: *p = h->lumHist; *bytes = 257 * sizeof(uint32_t); break;
    case TB_BUF_LOCAL_ACCUM_RGBW: *p = st.accum; *bytes = 16 * n; break;
    default: return fail(h, TB_ERR_INVALID_ARG, "unknown buffer kind");
    }
    return TB_OK;
}

TB_API int tb_buffer_size(TbHandle* h, uint32_t kind, uint64_t* bytes) {
    if (!h || !bytes) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    if (!h->width) return fail(h, TB_ERR_STATE, "no frame buffers (tb_resize)");
    void* p;
    return buffer_info(h, kind, &p, bytes);
}

TB_API int tb_device_buffer(TbHandle* h, uint32_t kind, void** devPtr, uint64_t* bytes) {
    if (!h || !devPtr || !bytes) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    if (!h->width) return fail(h, TB_ERR_STATE, "no frame buffers (tb_resize)");
    if (kind <= TB_BUF_RESOLVED_RGB) { int rr = job_wide_current(h); if (rr != TB_OK) return rr; }
    int rc = buffer_info(h, kind, devPtr, bytes);
    if (rc == TB_OK && kind == TB_BUF_RESOLVED_RGB) {
        CUDA_OK(h, cudaSetDevice(h->device));
        CUDA_OK(h, resolve_rgb(accum_view(h), h->resolved, h->width * h->height, h->stream, h->lc));
        CUDA_OK(h, cudaStreamSynchronize(h->stream));
    }
    return rc;
}

TB_API int tb_readback(TbHandle* h, uint32_t kind, void* dst, uint64_t bytes) {
    if (!h || !dst) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    if (!h->width) return fail(h, TB_ERR_STATE, "no frame buffers (tb_resize)");
    void* p; uint64_t need;
    if (kind <= TB_BUF_RESOLVED_RGB) { int rr = job_wide_current(h); if (rr != TB_OK) return rr; }
    int rc = buffer_info(h, kind, &p, &need);
    if (rc != TB_OK) return rc;
    if (bytes < need) return fail(h, TB_ERR_INVALID_ARG, "destination too small");
    CUDA_OK(h, cudaSetDevice(h->device));
    if (kind == TB_BUF_RESOLVED_RGB) CUDA_OK(h, resolve_rgb(accum_view(h), h->resolved, h->width * h->height, h->stream, h->lc));
    CUDA_OK(h, cudaMemcpyAsync(dst, p, need, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return TB_OK;
}

// ------------------------------------------------------------ post-process
TB_API int tb_get_default_postprocess_settings(TbPostProcessSettings* s) {
    if (!s) return TB_ERR_INVALID_ARG;
    s->ExposureMultiplier = 1.0f; s->TonemapType = TB_TONEMAP_AGX_PUNCHY; s->UseGammaCorrection = 1; // TracerBoy.h:308-313
    s->UseAutoExposure = 1; s->VarianceMultiplier = 1.0f;                                             // :298
    return TB_OK;
}

TB_API int tb_postprocess(TbHandle* h, uint32_t outputType, const TbPostProcessSettings* s) {
    if (!h || !s) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    if (!h->width) return fail(h, TB_ERR_STATE, "no frame buffers (tb_resize)");
    const void* in = nullptr;
    bool scalar = false;
    switch (outputType) { // GetOutputSRV, TracerBoy.cpp:2354-2383
    case TB_OUTPUT_LIT: case TB_OUTPUT_LUMINANCE: case TB_OUTPUT_LIVE_WAVES: {
        int rr = job_wide_current(h);
        if (rr != TB_OK) return rr;
        in = accum_view(h);
        break;
    }
    case TB_OUTPUT_ALBEDO: case TB_OUTPUT_LIVE_PIXELS: case TB_OUTPUT_HEATMAP: in = h->st.aovAlbedo; break;
    case TB_OUTPUT_NORMALS: in = h->st.aovNormal; break;
    case TB_OUTPUT_DEPTH: in = h->st.aovDepth; scalar = true; break;
    case TB_OUTPUT_MOTION_VECTORS: case TB_OUTPUT_LUMINANCE_VARIANCE:
        return fail(h, TB_ERR_NOT_IMPL, "motion vectors / luminance variance are not produced by this path");
    default: return fail(h, TB_ERR_INVALID_ARG, "unknown output type");
    }
    CUDA_OK(h, cudaSetDevice(h->device));
    CUDA_OK(h, postprocess(in, scalar, h->st.aovAlbedo, h->width, h->height, outputType, *s, h->lumHist, h->post, h->post8,
                           h->numSMs, h->stream, h->lc));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return TB_OK;
}

TB_API int tb_postprocess_image(TbHandle* h, const float* inRGBA, const float* auxRGBA, uint32_t width, uint32_t height,
                                uint32_t outputType, const TbPostProcessSettings* s, float* outRGBA, uint8_t* outRGBA8,
                                uint32_t* hist, float* avgLum) {
    if (!h || !inRGBA || !s || !outRGBA) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    if (width == 0 || height == 0 || (uint64_t)width * height > (1ull << 28)) return fail(h, TB_ERR_INVALID_ARG, "bad resolution");
    if (outputType > TB_OUTPUT_HEATMAP) return fail(h, TB_ERR_INVALID_ARG, "unknown output type");
    CUDA_OK(h, cudaSetDevice(h->device));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    const size_t n = (size_t)width * height;
    std::vector<void*> tmp;
    struct Guard { std::vector<void*>& v; ~Guard() { free_list(v); } } guard{tmp};
    auto alloc = [&](void** p, size_t bytes) -> cudaError_t { cudaError_t e = cudaMalloc(p, bytes); if (e == cudaSuccess) tmp.push_back(*p); return e; };
    float4 *dIn = nullptr, *dAux = nullptr, *dOut = nullptr; uchar4* dOut8 = nullptr; uint32_t* dHist = nullptr;
    CUDA_OK(h, alloc((void**)&dIn, 16 * n)); CUDA_OK(h, alloc((void**)&dOut, 16 * n)); CUDA_OK(h, alloc((void**)&dOut8, 4 * n));
    CUDA_OK(h, alloc((void**)&dHist, 257 * sizeof(uint32_t)));
    CUDA_OK(h, cudaMemcpyAsync(dIn, inRGBA, 16 * n, cudaMemcpyHostToDevice, h->stream));
    if (auxRGBA) { CUDA_OK(h, alloc((void**)&dAux, 16 * n)); CUDA_OK(h, cudaMemcpyAsync(dAux, auxRGBA, 16 * n, cudaMemcpyHostToDevice, h->stream)); }
    CUDA_OK(h, postprocess(dIn, false, dAux, width, height, outputType, *s, dHist, dOut, dOut8, sms, h->stream, h->lc));
    CUDA_OK(h, cudaMemcpyAsync(outRGBA, dOut, 16 * n, cudaMemcpyDeviceToHost, h->stream));
    if (outRGBA8) CUDA_OK(h, cudaMemcpyAsync(outRGBA8, dOut8, 4 * n, cudaMemcpyDeviceToHost, h->stream));
    uint32_t hh[257];
    CUDA_OK(h, cudaMemcpyAsync(hh, dHist, sizeof(hh), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    if (hist) memcpy(hist, hh, 256 * sizeof(uint32_t));
    if (avgLum) memcpy(avgLum, &hh[256], 4);
    return TB_OK;
}

// host-only: no device needed
TB_API int tb_load_image_file(const char* path, uint32_t* width, uint32_t* height, uint32_t* format, int* hasAlpha, void* pixels,
                              uint64_t capBytes, char* err, size_t errCap) {
    auto bad = [&](int code, const std::string& m) { if (err && errCap) { strncpy(err, m.c_str(), errCap - 1); err[errCap - 1] = 0; } return code; };
    if (!path || !width || !height || !format) return bad(TB_ERR_INVALID_ARG, "null argument");
    tb::Image img;
    std::string e;
    bool alpha = false;
    if (!tb::load_image_file(path, img, &alpha, e)) return bad(e.find("unsupported texture format") != std::string::npos ? TB_ERR_NOT_IMPL : TB_ERR_IO, e);
    *width = img.width; *height = img.height; *format = img.format;
    if (hasAlpha) *hasAlpha = alpha ? 1 : 0;
    if (pixels) {
        if (capBytes < img.data.size()) return bad(TB_ERR_INVALID_ARG, "destination too small");
        memcpy(pixels, img.data.data(), img.data.size());
    }
    return TB_OK;
}

TB_API int tb_write_image(const char* path, const void* pixels, uint32_t width, uint32_t height, uint32_t channels,
                          uint32_t bytesPerChannel, char* err, size_t errCap) {
    auto bad = [&](int code, const std::string& m) { if (err && errCap) { strncpy(err, m.c_str(), errCap - 1); err[errCap - 1] = 0; } return code; };
    if (!path || !pixels || !width || !height) return bad(TB_ERR_INVALID_ARG, "null argument");
    std::string p(path), ext = p.size() >= 4 ? p.substr(p.size() - 4) : "", e;
    bool ok;
    if (ext == ".png") {
        if (channels != 4 || bytesPerChannel != 1) return bad(TB_ERR_INVALID_ARG, ".png needs 4 x 8-bit channels");
        ok = tb::save_png_rgba8(p, (const uint8_t*)pixels, width, height, e);
    } else if (ext == ".exr" || ext == ".pfm") {
        if ((channels != 3 && channels != 4) || bytesPerChannel != 4) return bad(TB_ERR_INVALID_ARG, "float image formats need 3 or 4 x 32-bit float channels");
        ok = ext == ".exr" ? tb::save_exr_f32(p, (const float*)pixels, width, height, (int)channels, e)
                           : tb::save_pfm_rgb(p, (const float*)pixels, width, height, (int)channels, e);
    } else return bad(TB_ERR_NOT_IMPL, "unsupported image extension (use .png, .exr or .pfm)");
    return ok ? TB_OK : bad(TB_ERR_IO, e);
}

TB_API int tb_save_image(TbHandle* h, uint32_t kind, const char* path) {
    if (!h || !path) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    if (!h->width) return fail(h, TB_ERR_STATE, "no frame buffers (tb_resize)");
    uint64_t bytes = 0;
    int rc = tb_buffer_size(h, kind, &bytes);
    if (rc != TB_OK) return rc;
    const uint64_t n = (uint64_t)h->width * h->height;
    std::vector<uint8_t> host(bytes);
    rc = tb_readback(h, kind, host.data(), bytes);
    if (rc != TB_OK) return rc;
    uint32_t channels, bpc;
    if (kind == TB_BUF_BACKBUFFER_RGBA8) { channels = 4; bpc = 1; }
    else if (kind == TB_BUF_PRIMARY_HIT_IDS || kind == TB_BUF_RAY_COUNTERS || kind == TB_BUF_LUMINANCE_HISTOGRAM || (bytes != 16 * n && bytes != 12 * n))
        return fail(h, TB_ERR_INVALID_ARG, "not an image buffer kind");
    else { channels = bytes == 16 * n ? 4 : 3; bpc = 4; }
    char err[512] = {0};
    rc = tb_write_image(path, host.data(), h->width, h->height, channels, bpc, err, sizeof(err));
    if (rc != TB_OK) return fail(h, rc, err);
    return TB_OK;
}

TB_API int tb_temporal_accumulate_image(TbHandle* h, const TbTemporalAccumulationParams* p, uint32_t width, uint32_t height,
                                        const float* history, const float* current, const float* worldPos,
                                        const float* prevWorldPos, const float* normals, const float* momentHistory,
                                        float* outColor, float* outMoment) {
    if (!h || !p || !history || !current || !worldPos || !prevWorldPos || !normals || !outColor) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    if (p->OutputMomentInformation && !momentHistory) return fail(h, TB_ERR_INVALID_ARG, "moment history required");
    if (width == 0 || height == 0 || (uint64_t)width * height > (1ull << 28)) return fail(h, TB_ERR_INVALID_ARG, "bad resolution");
    CUDA_OK(h, cudaSetDevice(h->device));
    const size_t n = (size_t)width * height;
    std::vector<void*> tmp;
    struct Guard { std::vector<void*>& v; ~Guard() { free_list(v); } } guard{tmp};
    const float* src[6] = {history, current, worldPos, prevWorldPos, normals, momentHistory};
    float4* dev[8] = {};
    for (int i = 0; i < 8; i++) {
        if (i < 6 && !src[i]) continue;
        void* q = nullptr;
        CUDA_OK(h, cudaMalloc(&q, 16 * n));
        tmp.push_back(q);
        dev[i] = (float4*)q;
        if (i < 6) CUDA_OK(h, cudaMemcpyAsync(q, src[i], 16 * n, cudaMemcpyHostToDevice, h->stream));
    }
    CUDA_OK(h, temporal_accumulate(*p, width, height, dev[0], dev[1], dev[2], dev[3], dev[4], dev[5], dev[6], dev[7], h->stream, h->lc));
    CUDA_OK(h, cudaMemcpyAsync(outColor, dev[6], 16 * n, cudaMemcpyDeviceToHost, h->stream));
    if (outMoment && p->OutputMomentInformation) CUDA_OK(h, cudaMemcpyAsync(outMoment, dev[7], 16 * n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return TB_OK;
}

TB_API int tb_get_render_stats(TbHandle* h, TbRenderStats* out) {
    if (!h || !out) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    memset(out, 0, sizeof(*out));
    if (h->st.stats) {
        unsigned long long s[TB_STATS_WORDS];
        CUDA_OK(h, cudaSetDevice(h->device));
        CUDA_OK(h, cudaMemcpyAsync(s, h->st.stats, sizeof(s), cudaMemcpyDeviceToHost, h->stream));
        CUDA_OK(h, cudaStreamSynchronize(h->stream));
        out->RaysTraced = s[0] + s[3] + s[6]; out->BoxesTested = s[1] + s[4] + s[7]; out->TrianglesTested = s[2] + s[5] + s[8];
        out->ExtendRays = s[0]; out->ExtendBoxesTested = s[1]; out->ExtendTrianglesTested = s[2];
        out->ResumeRays = s[6]; out->ResumeBoxesTested = s[7]; out->ResumeTrianglesTested = s[8];
        for (int b = 0; b < 32; b++) out->RaysByBounce[b] = s[16 + b];
    }
    for (int b = 0; b < 32; b++) { out->BounceMilliseconds[b] = h->bounceMs[b]; out->BounceExtendMilliseconds[b] = h->bounceExtendMs[b]; }
    out->ExtendLaunches = h->extendLaunches; out->ExtendMilliseconds = h->extendMs; out->ShadeMilliseconds = h->shadeMs; out->ResumeMilliseconds = h->resumeMs;
    out->PathsStarted = h->pathsStarted;
    out->KernelLaunches = h->lc.count;
    out->DeviceMilliseconds = h->deviceMs;
    return TB_OK;
}
TB_API int tb_reset_render_stats(TbHandle* h) {
    if (!h) return TB_ERR_INVALID_ARG;
    if (h->st.stats) { CUDA_OK(h, cudaSetDevice(h->device)); CUDA_OK(h, cudaMemsetAsync(h->st.stats, 0, 8 * TB_STATS_WORDS, h->stream)); }
    memset(h->bounceMs, 0, sizeof(h->bounceMs)); memset(h->bounceExtendMs, 0, sizeof(h->bounceExtendMs));
    h->pathsStarted = 0; h->lc.count = 0; h->deviceMs = 0.0;
    h->extendMs = h->shadeMs = h->resumeMs = 0.0; h->extendLaunches = 0;
    return TB_OK;
}
TB_API int tb_set_frames_in_flight(TbHandle* h, uint32_t n) {
    if (!h || n > 64) return fail(h, TB_ERR_INVALID_ARG, "frames in flight must be in [0,64] (0 = automatic)");
    if (n == h->framesInFlight && n == h->slots.size()) return TB_OK;
    h->framesInFlight = n;
    if (h->width) { // re-create the frame buffers with the new slot count
        uint32_t w = h->width, hh = h->height;
        CUDA_OK(h, cudaSetDevice(h->device));
        CUDA_OK(h, cudaStreamSynchronize(h->stream));
        free_frame(h);
        return tb_resize(h, w, hh);
    }
    return TB_OK;
}
TB_API int tb_set_shadow_mode(TbHandle* h, int mode) {
    if (!h || mode < 0 || mode > 2) return fail(h, TB_ERR_INVALID_ARG, "shadow mode must be 0 (inline), 1 (queue) or 2 (auto)");
    h->options.shadowMode = mode;
    return TB_OK;
}
TB_API int tb_set_ray_sort(TbHandle* h, int mode) {
    if (!h || mode < 0 || mode > 4 || mode == 2) return fail(h, TB_ERR_INVALID_ARG, "ray sort must be 0 (off), 1 (bounce queue), 3 (bounce + shadow queues) or 4 (auto)");
    h->options.sortRays = mode == 4 ? 2 : mode;
    return TB_OK;
}
TB_API int tb_set_material_sort(TbHandle* h, int mode) {
    if (!h || mode < 0 || mode > 2) return fail(h, TB_ERR_INVALID_ARG, "material sort must be 0 (off), 1 (on) or 2 (auto)");
    h->options.materialSort = mode;
    return TB_OK;
}
TB_API int tb_set_profiling(TbHandle* h, int enable) {
    if (!h) return TB_ERR_INVALID_ARG;
    if (enable < 0 || enable > 2) return fail(h, TB_ERR_INVALID_ARG, "profiling mode must be 0, 1 or 2");
    h->profiling = enable;
    return TB_OK;
}
TB_API int tb_synchronize(TbHandle* h) {
    if (!h) return TB_ERR_INVALID_ARG;
    CUDA_OK(h, cudaSetDevice(h->device));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return TB_OK;
}

TB_API int tb_is_material_id_valid(TbHandle* h, int id) { return h && h->sceneLoaded && id >= 0 && id < (int)h->scene.materials.size(); }
TB_API int tb_get_material(TbHandle* h, int id, TbMaterial* out, char* name, uint32_t cap) {
    if (!h || !out) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    if (!tb_is_material_id_valid(h, id)) return fail(h, TB_ERR_INVALID_ARG, "invalid material id");
    *out = h->scene.materials[id];
    if (name && cap) { strncpy(name, id < (int)h->scene.materialNames.size() ? h->scene.materialNames[id].c_str() : "", cap - 1); name[cap - 1] = 0; }
    return TB_OK;
}
TB_API int tb_set_material(TbHandle* h, int id, const TbMaterial* m) {
    if (!h || !m) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    if (!tb_is_material_id_valid(h, id)) return fail(h, TB_ERR_INVALID_ARG, "invalid material id");
    std::string why;
    if (!tb::validate_material(h->scene, *m, why)) return fail(h, TB_ERR_INVALID_ARG, why); // the kernels index textures / mixed materials unchecked
    h->scene.materials[id] = *m;
    h->options.sceneHasSSS = scene_has_sss(h->scene); // also through mix materials that reference the edited one
    h->options.sceneMaterialClasses = scene_material_classes(h->scene);
    CUDA_OK(h, cudaSetDevice(h->device));
    { int rc = upload_geometry_classes(h); if (rc != TB_OK) return rc; }
    CUDA_OK(h, cudaMemcpyAsync((void*)(h->dscene.materials + id), m, sizeof(*m), cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    h->samplesRendered = 0;
    return TB_OK;
}

// ---- SW-RT seam. A caller-owned acceleration structure (`dst` of tb_bvh_build_device) is laid out as
//   [0, 116 N - 16)          the reference layout (RayTracingHlslCompat.h:344-398)
//   [A, A + 256)             AsTrailer, A = the reference layout's size rounded up to 256
//   [A + 256, ...)           PairNode[max(N - 1, 1)], then (256-aligned) WideTri[N]: the traversal layout
// so the buffer is self-describing: any handle can trace against it.
namespace {
struct AsTrailer { char magic[8]; uint32_t numPrims, depth; RefNode root; };
inline uint64_t up256(uint64_t v) { return (v + 255) & ~255ull; }
struct AsLayout { uint64_t trailer, pairs, tris, end; };
AsLayout as_layout(uint32_t n) {
    AsLayout L;
    L.trailer = up256(bvh_ref_bytes(n));
    L.pairs = L.trailer + 256;
    L.tris = up256(L.pairs + sizeof(PairNode) * (uint64_t)(n > 1 ? n - 1 : 1));
    L.end = L.tris + sizeof(WideTri) * (uint64_t)n;
    return L;
}
int count_triangles(const TbGeometryDesc* geoms, uint32_t n, uint64_t* tris) {
    *tris = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (geoms[i].Indices == nullptr && geoms[i].IndexFormat != 0) return TB_ERR_INVALID_ARG; // LoadPrimitivesPass.cpp:73-76
        if (geoms[i].IndexFormat != 0 && geoms[i].IndexFormat != 2 && geoms[i].IndexFormat != 4) return TB_ERR_INVALID_ARG;
        uint32_t vc = geoms[i].IndexFormat == 0 ? geoms[i].VertexCount : geoms[i].IndexCount;
        *tris += vc / 3;
    }
    return TB_OK;
}
} // namespace

TB_API int tb_bvh_prebuild_info(const TbGeometryDesc* geoms, uint32_t n, TbPrebuildInfo* out) {
    if (!geoms || !out) return TB_ERR_INVALID_ARG;
    uint64_t tris = 0;
    int rc = count_triangles(geoms, n, &tris);
    if (rc != TB_OK) return rc;
    memset(out, 0, sizeof(*out));
    if (tris == 0) return TB_OK;
    if (tris > tb_max_triangles()) return TB_ERR_INVALID_ARG; // E_INVALIDARG: the result would not fit the layout's 32-bit offsets
    out->ReferenceLayoutSizeInBytes = bvh_ref_bytes((uint32_t)tris);
    out->ResultDataMaxSizeInBytes = as_layout((uint32_t)tris).end;
    out->ScratchDataSizeInBytes = bvh_scratch_bytes((uint32_t)tris);
    out->UpdateScratchDataSizeInBytes = bvh_update_scratch_bytes((uint32_t)tris);
    return TB_OK;
}

TB_API int tb_bvh_build_device(TbHandle* h, const TbGeometryDesc* geoms, uint32_t n, uint32_t flags, void* dst, uint64_t dstBytes,
                               void* scratch, uint64_t scratchBytes, void* cudaStream) {
    if (!h || !geoms || n == 0 || !dst) return fail(h, TB_ERR_INVALID_ARG, "null/empty argument");
    uint64_t tris = 0;
    if (count_triangles(geoms, n, &tris) != TB_OK) return fail(h, TB_ERR_INVALID_ARG, "bad index format (If the index buffer is null, the index format must be 0)");
    if (tris == 0) return fail(h, TB_ERR_INVALID_ARG, "no triangles");
    if (tris > tb_max_triangles()) return fail(h, TB_ERR_INVALID_ARG, "too many triangles");
    const uint32_t N = (uint32_t)tris;
    const AsLayout L = as_layout(N);
    if (dstBytes < L.end) return fail(h, TB_ERR_INVALID_ARG, "destination smaller than ResultDataMaxSizeInBytes");
    if (scratch && scratchBytes < bvh_scratch_bytes(N)) return fail(h, TB_ERR_INVALID_ARG, "scratch smaller than ScratchDataSizeInBytes");
    if (((uintptr_t)dst & 255) || ((uintptr_t)scratch & 255)) return fail(h, TB_ERR_INVALID_ARG, "dst and scratch must be 256-byte aligned (D3D12_RAYTRACING_ACCELERATION_STRUCTURE_BYTE_ALIGNMENT)");
    CUDA_OK(h, cudaSetDevice(h->device));
    cudaStream_t stream = cudaStream ? (cudaStream_t)cudaStream : h->stream;
    std::vector<BuildGeometry> descs(n);
    std::vector<uint32_t> prefix(n);
    uint32_t at = 0;
    for (uint32_t g = 0; g < n; g++) {
        const TbGeometryDesc& G = geoms[g];
        if (!G.Positions || G.PositionStrideBytes < 12 || G.PositionStrideBytes % 4) return fail(h, TB_ERR_INVALID_ARG, "bad vertex buffer");
        prefix[g] = at;
        at += (G.IndexFormat == 0 ? G.VertexCount : G.IndexCount) / 3;
        descs[g] = {(const uint8_t*)G.Positions, G.Indices, G.Transform3x4, G.PositionStrideBytes, G.IndexFormat, G.GeometryFlags, 0u};
    }
    DeviceBvh b;
    b.ref = (uint8_t*)dst; b.refBytes = bvh_ref_bytes(N);
    b.pairs = (PairNode*)((uint8_t*)dst + L.pairs); b.tris = (WideTri*)((uint8_t*)dst + L.tris);
    int rc = run_build(h, descs, prefix, N, flags, b, scratch, stream);
    if (rc != TB_OK) return rc;
    AsTrailer t;
    memset(&t, 0, sizeof(t));
    memcpy(t.magic, "TBAS0001", 8);
    t.numPrims = N; t.depth = b.depth; t.root = b.root;
    CUDA_OK(h, cudaMemcpyAsync((uint8_t*)dst + L.trailer, &t, sizeof(t), cudaMemcpyHostToDevice, stream));
    CUDA_OK(h, cudaStreamSynchronize(stream));
    h->deviceBuilds[dst] = b;
    return TB_OK;
}

// resolves a caller-owned acceleration structure to its DeviceBvh (cached, else from the structure's own trailer)
static int resolve_as(TbHandle* h, const void* as, uint64_t asBytes, cudaStream_t stream, DeviceBvh& b) {
    auto it = h->deviceBuilds.find(as);
    if (it != h->deviceBuilds.end()) { b = it->second; return TB_OK; }
    uint32_t hd[4];
    CUDA_OK(h, cudaMemcpyAsync(hd, as, sizeof(hd), cudaMemcpyDeviceToHost, stream));
    CUDA_OK(h, cudaStreamSynchronize(stream));
    // the first four words of the reference layout give its size: offsetToPrimitiveMetaData + 12 N == 116 N - 16
    if (hd[0] != 16 || hd[3] < 100 || (hd[3] + 16ull) % 116) return fail(h, TB_ERR_INVALID_ARG, "not an acceleration structure built by tb_bvh_build_device");
    const uint32_t N = (uint32_t)((hd[3] + 16ull) / 116);
    const AsLayout L = as_layout(N);
    if (asBytes < L.end) return fail(h, TB_ERR_INVALID_ARG, "acceleration structure buffer too small");
    AsTrailer t;
    CUDA_OK(h, cudaMemcpyAsync(&t, (const uint8_t*)as + L.trailer, sizeof(t), cudaMemcpyDeviceToHost, stream));
    CUDA_OK(h, cudaStreamSynchronize(stream));
    if (memcmp(t.magic, "TBAS0001", 8) != 0 || t.numPrims != N || t.depth > TB_STACK_DEPTH) return fail(h, TB_ERR_INVALID_ARG, "acceleration structure trailer is missing or corrupt");
    b.ref = (uint8_t*)as; b.refBytes = bvh_ref_bytes(N);
    b.pairs = (PairNode*)((uint8_t*)as + L.pairs); b.tris = (WideTri*)((uint8_t*)as + L.tris);
    b.root = t.root; b.numPrims = N; b.depth = t.depth;
    h->deviceBuilds[as] = b;
    return TB_OK;
}

TB_API int tb_trace_rays_device(TbHandle* h, const void* as, uint64_t asBytes, const TbRay* dRays, uint64_t n, TbHit* dHits, void* cudaStream) {
    if (!h || (n && (!dRays || !dHits))) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    CUDA_OK(h, cudaSetDevice(h->device));
    cudaStream_t stream = cudaStream ? (cudaStream_t)cudaStream : h->stream;
    DeviceBvh b;
    if (!as) {
        if (!h->sceneLoaded) return fail(h, TB_ERR_STATE, "no acceleration structure");
        b = h->bvh;
    } else {
        int rc = resolve_as(h, as, asBytes, stream, b);
        if (rc != TB_OK) return rc;
    }
    if (n == 0) return TB_OK;
    CUDA_OK(h, trace_rays(b, dRays, n, dHits, h->numSMs, stream, h->lc)); // asynchronous on `stream`, like a dispatch
    return TB_OK;
}

extern "C++" {
namespace tbh {
// a bottom-level structure named by an instance desc: the handle's cache, else the structure's own trailer
int resolve_bottom_level(TbHandle* h, const void* as, cudaStream_t stream, DeviceBvh& out) {
    if (!as) return fail(h, TB_ERR_INVALID_ARG, "instance without a bottom-level acceleration structure");
    return resolve_as(h, as, ~0ull, stream, out);
}
} // namespace tbh
} // extern "C++"

TB_API int tb_bvh_update_device(TbHandle* h, const TbGeometryDesc* geoms, uint32_t n, void* dst, uint64_t dstBytes,
                                void* scratch, uint64_t scratchBytes, void* cudaStream) {
    if (!h || !geoms || n == 0 || !dst) return fail(h, TB_ERR_INVALID_ARG, "null/empty argument");
    uint64_t tris = 0;
    if (count_triangles(geoms, n, &tris) != TB_OK) return fail(h, TB_ERR_INVALID_ARG, "bad index format");
    CUDA_OK(h, cudaSetDevice(h->device));
    cudaStream_t stream = cudaStream ? (cudaStream_t)cudaStream : h->stream;
    DeviceBvh b;
    int rc = resolve_as(h, dst, dstBytes, stream, b);
    if (rc != TB_OK) return rc;
    // "the geometry descs of an update must match the build's: same counts, same index buffers' topology" (D3D12 spec)
    if (tris != b.numPrims) return fail(h, TB_ERR_INVALID_ARG, "an update needs the triangle count of the original build");
    if (scratch && scratchBytes < bvh_update_scratch_bytes(b.numPrims)) return fail(h, TB_ERR_INVALID_ARG, "scratch smaller than UpdateScratchDataSizeInBytes");
    std::vector<BuildGeometry> descs(n);
    for (uint32_t g = 0; g < n; g++) {
        const TbGeometryDesc& G = geoms[g];
        if (!G.Positions || G.PositionStrideBytes < 12 || G.PositionStrideBytes % 4) return fail(h, TB_ERR_INVALID_ARG, "bad vertex buffer");
        descs[g] = {(const uint8_t*)G.Positions, G.Indices, G.Transform3x4, G.PositionStrideBytes, G.IndexFormat, G.GeometryFlags, 0u};
    }
    void* owned[2] = {nullptr, nullptr};
    struct Guard { void** p; ~Guard() { for (int i = 0; i < 2; i++) if (p[i]) cudaFree(p[i]); } } guard{owned};
    CUDA_OK(h, cudaMalloc(&owned[0], sizeof(BuildGeometry) * descs.size()));
    CUDA_OK(h, cudaMemcpyAsync(owned[0], descs.data(), sizeof(BuildGeometry) * descs.size(), cudaMemcpyHostToDevice, stream));
    if (!scratch) { CUDA_OK(h, cudaMalloc(&owned[1], bvh_update_scratch_bytes(b.numPrims))); scratch = owned[1]; }
    CUDA_OK(h, update_bvh((const BuildGeometry*)owned[0], b.numPrims, b, scratch, stream, h->lc));
    AsTrailer t;
    memset(&t, 0, sizeof(t));
    memcpy(t.magic, "TBAS0001", 8);
    t.numPrims = b.numPrims; t.depth = b.depth; t.root = b.root;
    CUDA_OK(h, cudaMemcpyAsync((uint8_t*)dst + as_layout(b.numPrims).trailer, &t, sizeof(t), cudaMemcpyHostToDevice, stream));
    CUDA_OK(h, cudaStreamSynchronize(stream));
    h->deviceBuilds[dst] = b;
    return TB_OK;
}

TB_API int tb_bvh_forget_device(TbHandle* h, const void* as) {
    if (!h) return TB_ERR_INVALID_ARG;
    h->deviceBuilds.erase(as);
    h->topLevelBuilds.erase(as);
    return TB_OK;
}

TB_API int tb_get_bvh_depth(TbHandle* h, uint32_t* depth) {
    if (!h || !depth) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    if (!h->sceneLoaded) return fail(h, TB_ERR_STATE, "no scene loaded");
    *depth = h->bvh.depth;
    return TB_OK;
}

TB_API int tb_bvh_build(TbHandle* h, const TbGeometryDesc* geoms, uint32_t n, uint32_t flags) {
    if (!h || !geoms || n == 0) return fail(h, TB_ERR_INVALID_ARG, "null/empty geometry list");
    tb::Scene s;
    s.materials.push_back(tb::default_material({0, 0, 0}));
    s.materials[0].albedo = {0.5f, 0.5f, 0.5f};
    s.materials[0].Flags |= TB_NO_SPECULAR_MATERIAL_FLAG | TB_NO_ALPHA_MATERIAL_FLAG;
    s.materialNames.push_back("default");
    for (uint32_t g = 0; g < n; g++) {
        const TbGeometryDesc& G = geoms[g];
        if (!G.Positions || G.PositionStrideBytes < 12 || G.PositionStrideBytes % 4) return fail(h, TB_ERR_INVALID_ARG, "bad vertex buffer");
        if (G.Indices == nullptr && G.IndexFormat != 0) return fail(h, TB_ERR_INVALID_ARG, "If the index buffer is null, the index format must be 0");
        if (G.IndexFormat != 0 && G.IndexFormat != 2 && G.IndexFormat != 4) return fail(h, TB_ERR_INVALID_ARG, "index format must be 0, 2 or 4");
        std::vector<TbFloat3> pos(G.VertexCount);
        for (uint32_t v = 0; v < G.VertexCount; v++) {
            const float* p = (const float*)((const uint8_t*)G.Positions + (size_t)v * G.PositionStrideBytes);
            float x = p[0], y = p[1], z = p[2];
            if (G.Transform3x4) { // TransformVertex, BottomLevelLoadTriangles.hlsli:83-86: mul(float3x4, float4(v,1))
                const float* m = G.Transform3x4;
                float tx = ((m[0] * x + m[1] * y) + m[2] * z) + m[3];
                float ty = ((m[4] * x + m[5] * y) + m[6] * z) + m[7];
                float tz = ((m[8] * x + m[9] * y) + m[10] * z) + m[11];
                x = tx; y = ty; z = tz;
            }
            pos[v] = {x, y, z};
        }
        uint32_t count = G.IndexFormat == 0 ? G.VertexCount : G.IndexCount;
        count -= count % 3;
        std::vector<uint32_t> idx(count);
        for (uint32_t i = 0; i < count; i++) {
            uint32_t v = G.IndexFormat == 0 ? i : (G.IndexFormat == 2 ? ((const uint16_t*)G.Indices)[i] : ((const uint32_t*)G.Indices)[i]);
            if (v >= G.VertexCount) return fail(h, TB_ERR_INVALID_ARG, "index out of range");
            idx[i] = v;
        }
        uint32_t gi = tb::append_geometry(s, pos.data(), nullptr, nullptr, nullptr, G.VertexCount, idx.data(), count, 0);
        s.geoms[gi].GeometryFlags = G.GeometryFlags;
    }
    h->scene = std::move(s);
    return upload_and_build(h, flags);
}

TB_API int tb_trace_rays(TbHandle* h, const TbRay* rays, uint64_t n, TbHit* hits) {
    if (!h || (n && (!rays || !hits))) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    if (!h->sceneLoaded) return fail(h, TB_ERR_STATE, "no acceleration structure");
    if (n == 0) return TB_OK;
    CUDA_OK(h, cudaSetDevice(h->device));
    TbRay* dr = nullptr; TbHit* dh = nullptr;
    CUDA_OK(h, cudaMalloc((void**)&dr, sizeof(TbRay) * n));
    cudaError_t e = cudaMalloc((void**)&dh, sizeof(TbHit) * n);
    if (e != cudaSuccess) { cudaFree(dr); return fail(h, TB_ERR_OOM, cudaGetErrorString(e)); }
    e = cudaMemcpyAsync(dr, rays, sizeof(TbRay) * n, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = trace_rays(h->bvh, dr, n, dh, h->numSMs, h->stream, h->lc);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hits, dh, sizeof(TbHit) * n, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(dr); cudaFree(dh);
    if (e != cudaSuccess) return fail(h, TB_ERR_CUDA, cudaGetErrorString(e));
    return TB_OK;
}

} // extern "C"
