/*
 * tracerboy_b200.h — C ABI of the B200-native TracerBoy path-tracing hot path.
 *
 * This is the drop-in boundary for the part of wallisc/TracerBoy that sits
 * behind `class TracerBoy` (TracerBoy/TracerBoy.h:158-397) and behind the
 * D3D12 Raytracing Fallback Layer interface
 * (D3D12RaytracingFallback/Include/D3D12RaytracingFallback.h:76-173).
 *
 *   - plain C, `extern "C"`, pointers + sizes only; no torch/CUDA types
 *   - every entry point returns an int (TB_OK = 0, negative = TbStatus)
 *   - no exception crosses the ABI; message via tb_last_error()
 *   - one thread at a time per handle (same contract as the reference object,
 *     D3D12App.cpp:161-168); tb_load_scene may run on a worker thread while
 *     another thread polls tb_get_load_status()
 *
 * The POD structs below reproduce the reference's host<->shader data contract
 * byte for byte (TracerBoy/SharedShaderStructs.h); static_asserts at the bottom
 * pin the sizes.
 */
#ifndef TRACERBOY_B200_H
#define TRACERBOY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define TB_API
#else
#define TB_API __attribute__((visibility("default")))
#endif

/* ------------------------------------------------------------------ status */
/* Replaces VERIFY/VERIFY_HRESULT (TracerBoy/pch.h:45-47) and the fallback
 * layer's ThrowFailure(E_INVALIDARG/E_NOTIMPL) (D3D12RaytracingFallback/src/Util.h:12-32). */
typedef enum TbStatus {
    TB_OK = 0,
    TB_ERR_INVALID_ARG = -1, /* E_INVALIDARG */
    TB_ERR_NOT_IMPL = -2,    /* E_NOTIMPL   */
    TB_ERR_IO = -3,          /* parser std::runtime_error, missing file */
    TB_ERR_OOM = -4,
    TB_ERR_CUDA = -5,
    TB_ERR_NCCL = -6,
    TB_ERR_STATE = -7        /* call out of order, e.g. render before load */
} TbStatus;

/* ------------------------------------------------------- shared data (a1) */
typedef struct TbFloat2 { float x, y; } TbFloat2;
typedef struct TbFloat3 { float x, y, z; } TbFloat3;
typedef struct TbFloat4 { float x, y, z, w; } TbFloat4;

/* SharedShaderStructs.h:116-124 */
#define TB_DEFAULT_MATERIAL_FLAG 0x0
#define TB_METALLIC_MATERIAL_FLAG 0x1
#define TB_SUBSURFACE_SCATTER_MATERIAL_FLAG 0x2
#define TB_NO_SPECULAR_MATERIAL_FLAG 0x4
#define TB_MIX_MATERIAL_FLAG 0x8
#define TB_LIGHT_MATERIAL_FLAG 0x10
#define TB_NO_ALPHA_MATERIAL_FLAG 0x20
#define TB_HAIR_MATERIAL_FLAG 0x40
#define TB_SINGLE_SIDED_MATERIAL_FLAG 0x80

#define TB_INVALID_TEXTURE 0xffffffffu /* UINT_MAX, SharedRaytracing.h:60-63 */

/* SharedShaderStructs.h:141-161, 84 bytes */
typedef struct TbMaterial {
    TbFloat3 albedo;
    uint32_t albedoIndex;
    uint32_t alphaIndex;
    uint32_t normalMapIndex;
    uint32_t emissiveIndex;
    uint32_t specularMapIndex;
    float IOR;
    TbFloat3 absorption;
    float roughness;
    TbFloat3 scattering;
    TbFloat3 emissive;
    int32_t Flags;
    float SpecularCoef;
} TbMaterial;

#define TB_LIGHT_TYPE_AREA 0
#define TB_LIGHT_TYPE_DIRECTIONAL 1

/* SharedShaderStructs.h:92-111, 104 bytes */
typedef struct TbLight {
    uint32_t LightType;
    TbFloat3 LightColor;
    float SurfaceArea;
    TbFloat3 P0, P1, P2;
    TbFloat3 N0, N1, N2;
    TbFloat3 Direction;
} TbLight;

#define TB_IMAGE_TEXTURE_TYPE 0
#define TB_CHECKER_TEXTURE_TYPE 1
#define TB_SCALE_TEXTURE_TYPE 2
#define TB_NEEDS_GAMMA_CORRECTION_TEXTURE_FLAG 0x1

/* SharedShaderStructs.h:169-190, 80 bytes. DescriptorHeapIndex is reinterpreted
 * as the index into the scene's image table. */
typedef struct TbTextureData {
    uint32_t TextureType;
    uint32_t DescriptorHeapIndex;
    uint32_t TextureFlags;
    uint32_t Padding;
    TbFloat3 CheckerColor1;
    float UScale;
    TbFloat3 CheckerColor2;
    float VScale;
    uint32_t TextureIndex1;
    TbFloat3 ScaleColor1;
    uint32_t TextureIndex2;
    TbFloat3 ScaleColor2;
} TbTextureData;

/* SharedShaderStructs.h:85-90, 32 bytes (stride 8 floats, SharedHitGroup.h:11) */
typedef struct TbVertex {
    TbFloat3 Normal;
    TbFloat2 UV;
    TbFloat3 Tangent;
} TbVertex;

/* The per-geometry record the shading stage reads. It plays the role of
 * HitGroupShaderRecord (TracerBoy.cpp:31-41 == SharedHitGroup.h:13-23): the
 * 32-byte shader identifier and the descriptor-heap indices have no meaning
 * without D3D12, so the record keeps the material index plus element offsets
 * into the scene's single pooled vertex / index arrays. */
typedef struct TbGeometryRecord {
    uint32_t MaterialIndex;
    uint32_t VertexFirst;   /* first vertex of this geometry in the pooled arrays */
    uint32_t VertexCount;
    uint32_t IndexFirst;    /* first uint32 index of this geometry */
    uint32_t IndexCount;    /* 3 * triangle count */
    uint32_t GeometryFlags; /* D3D12_RAYTRACING_GEOMETRY_FLAG_*; 1 = OPAQUE (TracerBoy.cpp:1761) */
    uint32_t GeometryIndex;
    uint32_t Padding;
} TbGeometryRecord;

/* TracerBoy.h:59-67 */
typedef struct TbCamera {
    TbFloat3 Position;
    TbFloat3 LookAt;
    TbFloat3 Right;
    TbFloat3 Up;
    float LensHeight;
    float FocalDistance;
} TbCamera;

/* TracerBoy.h:69-77 */
typedef struct TbControllerState {
    float RightStickX, RightStickY, RightTrigger;
    float LeftStickX, LeftStickY, LeftTrigger;
} TbControllerState;

/* TracerBoy::CameraSettings, TracerBoy.h:386-390 */
typedef struct TbCameraSettings {
    float MovementSpeed;
    uint32_t IgnoreMouse;
} TbCameraSettings;

/* ---------------------------------------------------------------- settings */
/* TracerBoy.h:171-197 */
typedef enum TbOutputType {
    TB_OUTPUT_LIT = 0, TB_OUTPUT_ALBEDO, TB_OUTPUT_NORMALS, TB_OUTPUT_DEPTH,
    TB_OUTPUT_MOTION_VECTORS, TB_OUTPUT_LUMINANCE, TB_OUTPUT_LUMINANCE_VARIANCE,
    TB_OUTPUT_LIVE_PIXELS, TB_OUTPUT_LIVE_WAVES, TB_OUTPUT_HEATMAP
} TbOutputType;
typedef enum TbRenderMode { TB_RENDER_UNBIASED = 0, TB_RENDER_REALTIME = 1 } TbRenderMode;
typedef enum TbFilterType { TB_FILTER_BOX = 0, TB_FILTER_TRIANGLE = 1, TB_FILTER_GAUSSIAN = 2 } TbFilterType;

/* POD mirror of TracerBoy::OutputSettings (TracerBoy.h:212-288) restricted to the
 * members that reach PerFrameConstants (TracerBoy.cpp:2810-2849). Defaults come
 * from GetDefaultOutputSettings (TracerBoy.h:290-360). */
typedef struct TbOutputSettings {
    uint32_t OutputType;           /* TbOutputType, default Lit */
    uint32_t EnableNormalMaps;     /* default 0 */
    uint32_t RenderMode;           /* TbRenderMode, default Unbiased */
    /* DebugSettings */
    int32_t SampleLimit;           /* 0 = unlimited */
    float TimeLimitInSeconds;      /* 0 = unlimited */
    float DebugValue;              /* 1.0 */
    float DebugValue2;             /* 1.0 */
    /* CameraOutputSettings */
    float DOFFocalDistance;        /* 0 = pinhole */
    float ApertureWidth;           /* 0.075 */
    uint32_t FilterType;           /* Box */
    float FilterWidth;             /* 1.0 */
    /* DenoiserSettings members that reach the tracer */
    float FireflyClampValue;       /* 0 = off */
    float MaxZ;                    /* 10000 */
    /* PerformanceSettings */
    float ConvergencePercentage;   /* 0.001 (adaptive dispatch is compiled out: RayGenCommon.h:660) */
    uint32_t EnableNextEventEstimation;          /* 1 */
    uint32_t EnableSamplingImportanceResampling; /* 0 */
    uint32_t EnableBlueNoise;      /* 1 */
    int32_t MaxBounces;            /* 6 */
} TbOutputSettings;

/* Tonemap.h:3-10 */
typedef enum TbTonemapType {
    TB_TONEMAP_REINHARD = 0, TB_TONEMAP_ACES = 1, TB_TONEMAP_CLAMP = 2, TB_TONEMAP_UNCHARTED = 3,
    TB_TONEMAP_KHRONOS_PBR_NEUTRAL = 4, TB_TONEMAP_AGX = 5, TB_TONEMAP_AGX_PUNCHY = 6, TB_TONEMAP_GT = 7
} TbTonemapType;

/* PostProcessSettings (TracerBoy.h:222-229) + DebugSettings::m_VarianceMultiplier as they reach
 * PostProcessConstants (SharedPostProcessStructs.h:3-14, TracerBoy.cpp:3172-3180). Defaults from
 * GetDefaultOutputSettings (TracerBoy.h:298, 308-313). */
typedef struct TbPostProcessSettings {
    float ExposureMultiplier;     /* 1.0 */
    uint32_t TonemapType;         /* TbTonemapType, default AgX punchy */
    uint32_t UseGammaCorrection;  /* 1 */
    uint32_t UseAutoExposure;     /* 1: luminance histogram -> averaged luminance -> exposure (TracerBoy.cpp:2948-3039) */
    float VarianceMultiplier;     /* 1.0 */
} TbPostProcessSettings;

/* TemporalAccumulationConstants (TemporalAccumulationSharedShaderStructs.h:6-34) as TemporalAccumulationPass::Run
 * fills them (TemporalAccumulationPass.cpp:87-103). */
typedef struct TbTemporalAccumulationParams {
    TbCamera Camera;                  /* current frame: LensHeight and FocalDistance are read from here */
    TbCamera PrevCamera;              /* previous frame: Position, LookAt, Right, Up */
    float HistoryWeight;              /* 0.95 at both call sites (TracerBoy.cpp:3081, 3155) */
    uint32_t IgnoreHistory;
    uint32_t OutputMomentInformation; /* 1: also update (mean luminance, mean luminance^2, sample count); alpha = variance */
} TbTemporalAccumulationParams;

/* TracerBoy.h:114-128 */
typedef enum TbSceneLoadState {
    TB_LOAD_IDLE = 0, TB_LOADING_PBRT, TB_LOADING_HOST, TB_RECORDING_DEVICE_WORK,
    TB_WAITING_ON_GPU, TB_LOAD_FINISHED, TB_LOAD_FAILED
} TbSceneLoadState;
typedef struct TbSceneLoadStatus {
    uint32_t State;
    uint32_t InstancesLoaded;
    uint32_t TotalInstances;
} TbSceneLoadStatus;

/* TracerBoy.h:362-368 */
typedef struct TbReadbackStats {
    uint32_t ActiveWaves;
    uint32_t ActivePixels;
    float SelectedPixelDistance;
    int32_t SelectedMaterialID;
} TbReadbackStats;

typedef struct TbSceneInfo {
    uint32_t NumGeometries, NumTriangles, NumVertices, NumMaterials, NumLights,
        NumTextures, NumImages, HasEnvironmentMap;
} TbSceneInfo;

/* What tb_readback can return. All float images are row-major, top row first
 * (DispatchIndex.y = 0 is the top of the image as in the reference). */
typedef enum TbBufferKind {
    TB_BUF_ACCUM_RGBW = 0,     /* float4: OutputTexture (RayGenCommon.h:721-722) */
    TB_BUF_JITTERED_RGBW = 1,  /* float4: JitteredOutputTexture (:723-727) */
    TB_BUF_RESOLVED_RGB = 2,   /* float3: rgb / w (PostProcessCS.hlsl:23-27) */
    TB_BUF_AOV_NORMAL = 3,     /* float4: AOVNormals (:575-578) */
    TB_BUF_AOV_WORLDPOS = 4,   /* float4: AOVWorldPosition of the last frame (:712-719) */
    TB_BUF_AOV_DEPTH = 5,      /* float:  AOVDepth (:632-640) */
    TB_BUF_AOV_ALBEDO = 6,     /* float4: AOVCustomOutput (:524-530) */
    TB_BUF_AOV_EMISSIVE = 7,   /* float4: AOVEmissive (:532-535) */
    TB_BUF_PRIMARY_HIT_IDS = 8,/* uint2:  (geometryIndex, primitiveIndex) of the last frame's primary hit, 0xffffffff on miss */
    TB_BUF_RAY_COUNTERS = 9,   /* uint2:  per-pixel (TrianglesTested, BoxesTested) summed over the last frame's rays */
    /* written by tb_postprocess */
    TB_BUF_POSTPROCESS_RGBA = 10,   /* float4: PostProcessCS output before the back-buffer format conversion (PostProcessCS.hlsl:195) */
    TB_BUF_BACKBUFFER_RGBA8 = 11,   /* uchar4: the same after the UNORM8 store, i.e. what lands in m_pPostProcessOutput */
    TB_BUF_LUMINANCE_HISTOGRAM = 12,/* uint32[256] LuminanceHistogram + 1 float AveragedLuminance (1028 bytes) */
    /* multi-GPU: this rank's own accumulation buffer (kinds 0-2 serve the job-wide image once a communicator is attached) */
    TB_BUF_LOCAL_ACCUM_RGBW = 13
} TbBufferKind;

/* ----------------------------------------------------------- SW-RT boundary */
/* D3D12_RAYTRACING_GEOMETRY_DESC (triangles only; fallback layer rejects other
 * vertex formats, LoadPrimitivesPass.cpp:77-84). Pointers are HOST pointers;
 * the library copies them to the device. */
typedef struct TbGeometryDesc {
    const float* Positions;     /* float3 per vertex */
    uint32_t PositionStrideBytes; /* >= 12 */
    uint32_t VertexCount;
    const void* Indices;        /* NULL => non-indexed */
    uint32_t IndexFormat;       /* 0 = none, 2 = uint16, 4 = uint32 */
    uint32_t IndexCount;
    const float* Transform3x4;  /* optional row-major 3x4, may be NULL */
    uint32_t GeometryFlags;
} TbGeometryDesc;

typedef enum TbBvhBuildFlags {
    TB_BVH_BUILD_NONE = 0,
    TB_BVH_BUILD_PREFER_FAST_TRACE = 0x4, /* 3 treelet passes (TreeletReorder.cpp:63-108) */
    TB_BVH_BUILD_PREFER_FAST_BUILD = 0x8  /* 0 treelet passes */
} TbBvhBuildFlags;

/* GetRaytracingAccelerationStructurePrebuildInfo (D3D12RaytracingFallback.h:134-136). The caller allocates
 * ResultDataMaxSizeInBytes for `dst` and ScratchDataSizeInBytes for `scratch` of tb_bvh_build_device, both
 * 256-byte aligned device memory. The result starts with the reference layout (116 N - 16 bytes,
 * GpuBVH2Builder.cpp:459 = ReferenceLayoutSizeInBytes) and continues with this library's traversal layout. */
typedef struct TbPrebuildInfo {
    uint64_t ResultDataMaxSizeInBytes;
    uint64_t ScratchDataSizeInBytes;
    uint64_t UpdateScratchDataSizeInBytes;
    uint64_t ReferenceLayoutSizeInBytes; /* 116*N - 16 */
} TbPrebuildInfo;

/* D3D12_RAYTRACING_INSTANCE_DESC (64 bytes; RayTracingHlslCompat.h:217-224 reads it as RaytracingInstanceDesc). */
typedef struct TbInstanceDesc {
    float Transform[12];                                   /* object to world, row-major 3x4 */
    uint32_t InstanceIDAndMask;                            /* InstanceID : 24 | InstanceMask << 24 (mask 0 = never hit) */
    uint32_t InstanceContributionToHitGroupIndexAndFlags;  /* contribution : 24 | D3D12_RAYTRACING_INSTANCE_FLAGS << 24 */
    uint64_t AccelerationStructure;                        /* device address of a bottom-level structure (dst of tb_bvh_build_device) */
} TbInstanceDesc;

/* RayDesc as consumed by SoftwareRayQuery::TraceRayInline (TraverseFunction.hlsli:39-132) */
typedef struct TbRay {
    float Origin[3];
    float TMin;
    float Direction[3];
    float TMax;
} TbRay;

/* SoftwareHitData (TraverseFunction.hlsli:24-34) + the two counters (:46-47).
 * t < 0 on miss. */
typedef struct TbHit {
    float t;
    float b1, b2;             /* CommittedTriangleBarycentrics */
    uint32_t PrimitiveIndex;  /* primitive index inside its geometry */
    uint32_t GeometryIndex;
    uint32_t InstanceIndex;   /* always 0 on the single-level fast path */
    uint32_t TrianglesTested;
    uint32_t BoxesTested;
} TbHit;

typedef struct TbRenderStats {
    uint64_t RaysTraced;       /* every Intersect call: extend + shadow + SSS walk */
    uint64_t BoxesTested;
    uint64_t TrianglesTested;
    uint64_t PathsStarted;
    uint64_t KernelLaunches;   /* CUDA kernels launched by the library since the last reset */
    double DeviceMilliseconds; /* CUDA-event time of tb_render calls since the last reset */
    /* extend-stage (k_extend) share of the three counters above */
    uint64_t ExtendRays, ExtendBoxesTested, ExtendTrianglesTested;
    /* profiling mode only (tb_set_profiling): summed per-launch CUDA-event durations */
    uint64_t ExtendLaunches;
    double ExtendMilliseconds;
    double ShadeMilliseconds;
    /* rays that were suspended and finished in a k_extend_resume round */
    uint64_t ResumeRays, ResumeBoxesTested, ResumeTrianglesTested;
    double ResumeMilliseconds;
    /* every Intersect call by the bounce index of its path (extension, shadow and walk rays of bounce b; 31 = 31 and
     * above): bounce 0 holds the coherent camera rays, the others are the incoherent ones */
    uint64_t RaysByBounce[32];
    /* profiling mode only: device time of all stages of bounce b, and of its k_extend launch alone */
    double BounceMilliseconds[32];
    double BounceExtendMilliseconds[32];
} TbRenderStats;

/* Multi-GPU: how the frame is partitioned over the ranks of a communicator. */
#define TB_SHARD_SAMPLES 1u /* rank r renders frames r, r+N, ...; reduction = sum over the ranks in fixed rank order */
#define TB_SHARD_ROWS 2u    /* bands of 8 rows, band b on rank b mod N; reduction = gather of the owned bands (bit-identical to one GPU) */
#define TB_COMM_ID_BYTES 128 /* sizeof(ncclUniqueId) */

typedef struct TbCommInfo {
    uint32_t Rank, NumRanks, ShardMode, NcclVersion;
    uint64_t Reductions;                 /* tb_comm_reduce calls so far */
    uint64_t BytesReceivedPerReduction;  /* over NVLink, per rank */
    double LastReductionMilliseconds;    /* CUDA events around the whole exchange + combine */
    double TotalReductionMilliseconds;   /* the same, summed over all reductions */
    uint32_t Transport;                  /* TB_COMM_TRANSPORT_* of the last reduction (NCCL until the first one) */
    uint32_t Reserved;
} TbCommInfo;
#define TB_COMM_TRANSPORT_NCCL 0u /* ncclAllGather of the buffers + local combine kernels */
#define TB_COMM_TRANSPORT_PEER 1u /* one kernel per rank over the other GPUs' memory (CUDA IPC mappings, NVLink) */

typedef struct TbHandle TbHandle; /* opaque; owns all device memory */

/* ---------------------------------------------------------------- lifecycle */
/* TracerBoy::TracerBoy(ID3D12CommandQueue*) (TracerBoy.cpp:507-960). `device` is
 * the CUDA ordinal. */
TB_API int tb_create(int device, TbHandle** out);
TB_API void tb_destroy(TbHandle* h);
TB_API const char* tb_last_error(TbHandle* h); /* h may be NULL: last create error */
TB_API const char* tb_version(void);
/* Largest triangle count of one acceleration structure (37 025 580): the reference layout stores 32-bit byte
 * offsets (RayTracingHlslCompat.h:344-398), so its 116 N - 16 bytes must stay below 4 GiB. Larger inputs are
 * rejected with TB_ERR_INVALID_ARG by tb_load_scene / tb_bvh_build / tb_bvh_prebuild_info. */
TB_API uint64_t tb_max_triangles(void);

/* ------------------------------------------------------------- scene + bvh */
/* TracerBoy::LoadScene (TracerBoy.cpp:1065-2161). Accepts
 *   *.tbscene            the library's flattened scene cache
 *   *.pbrt / *.pbf       through the optional importer (libtb_pbrtimport.so)
 *   "synthetic:<name>?k=v&..."   built-in procedural scenes
 * then builds the BVH on the device (PREFER_FAST_TRACE, TracerBoy.cpp:1970). */
TB_API int tb_load_scene(TbHandle* h, const char* path);
TB_API int tb_load_scene_ex(TbHandle* h, const char* path, uint32_t bvhBuildFlags);
/* What the PBRT import does with object instances (`ObjectInstance` directives). LoadScene has a switch for it,
 * bInsertInstancesIntoBLAS (TracerBoy.cpp:1355), compiled as false: instances get a bottom-level structure each and
 * a top-level entry (:2031-2050) that the software path then never traverses (it renders BLASList[0], :2862).
 *   TB_INSTANCES_SKIP              the reference build's software path: instances are not rendered (default)
 *   TB_INSTANCES_INSERT_INTO_BLAS  the switch set: every instance joins the global bottom-level structure as the
 *                                  first shape of its object, its transform baked into the vertex data (:1367-1376,
 *                                  1623-1651)
 * Takes effect at the next tb_load_scene of a .pbrt / .pbf path. */
#define TB_INSTANCES_SKIP 0u
#define TB_INSTANCES_INSERT_INTO_BLAS 1u
TB_API int tb_set_instance_mode(TbHandle* h, uint32_t mode);
TB_API int tb_get_load_status(TbHandle* h, TbSceneLoadStatus* out);
TB_API int tb_save_scene(TbHandle* h, const char* tbscenePath); /* .pbf-cache role, TracerBoy.cpp:1200-1223 */
/* Host-only: import any supported scene path and write the .tbscene cache (no device needed). */
TB_API int tb_convert_scene(const char* inPath, const char* outTbscenePath, char* err, size_t errCap);
TB_API int tb_convert_scene_ex(const char* inPath, const char* outTbscenePath, uint32_t instanceMode, char* err, size_t errCap);
TB_API int tb_get_scene_info(TbHandle* h, TbSceneInfo* out);
TB_API int tb_get_bvh_size(TbHandle* h, uint64_t* bytes);
/* Copies the BVH in the reference's byte layout (RayTracingHlslCompat.h:344-398). */
TB_API int tb_get_bvh(TbHandle* h, void* dst, uint64_t bytes);
TB_API int tb_get_bvh_build_ms(TbHandle* h, double* ms);

/* --------------------------------------------------------- settings/camera */
TB_API int tb_get_default_settings(TbOutputSettings* out); /* TracerBoy::GetDefaultOutputSettings */
TB_API int tb_get_camera(TbHandle* h, TbCamera* out);
TB_API int tb_set_camera(TbHandle* h, const TbCamera* cam); /* invalidates history like TracerBoy::Update */
/* TracerBoy::Update (TracerBoy.cpp:3386-3500): mouse look (yaw about +Y, pitch about the XZ-aligned right axis),
 * WASD/QE and controller motion. keyboardInput has CHAR_MAX (127) entries indexed by character, as in the
 * reference; controller and cameraSettings may be NULL (no controller; speed 1, mouse honoured). A call that
 * moves the camera invalidates the history. Mouse look needs the output size, i.e. tb_resize first
 * (the reference tests m_pPostProcessOutput). */
TB_API int tb_update(TbHandle* h, int mouseX, int mouseY, const uint8_t* keyboardInput, float dt,
                     const TbControllerState* controller, const TbCameraSettings* cameraSettings);
/* The same step on caller-owned state (host only, needs no device): lastMouse[2] is m_mouseX/m_mouseY and is
 * updated; width/height 0 disables mouse look; *moved (optional) receives bCameraMoved. */
TB_API int tb_camera_update(TbCamera* camera, uint32_t lastMouse[2], uint32_t width, uint32_t height,
                            int mouseX, int mouseY, const uint8_t* keyboardInput, float dt,
                            const TbControllerState* controller, const TbCameraSettings* cameraSettings, int* moved);
TB_API int tb_resize(TbHandle* h, uint32_t width, uint32_t height); /* ResizeBuffersIfNeeded, TracerBoy.cpp:3624-3929 */
TB_API int tb_select_pixel(TbHandle* h, int x, int y);      /* TracerBoy::SelectPixel */
TB_API int tb_get_stats(TbHandle* h, TbReadbackStats* out);

/* ------------------------------------------------------------------ render */
/* nSamples x TracerBoy::Render (TracerBoy.cpp:2677-3369): each sample is one
 * GlobalFrameCount step. `time` replaces the wall clock of TracerBoy.cpp:2817 and
 * is an explicit seed input (use 0 for reproducible renders). */
TB_API int tb_render(TbHandle* h, const TbOutputSettings* settings, uint32_t nSamples, float time);
TB_API int tb_samples_rendered(TbHandle* h, uint32_t* out); /* GetNumberOfSamplesSinceLastInvalidate */
TB_API int tb_invalidate_history(TbHandle* h);
/* Sample-index sharding for multi-GPU: this handle renders only
 * frames f with f % stride == offset. Default (0, 1). */
TB_API int tb_set_frame_shard(TbHandle* h, uint32_t offset, uint32_t stride);
/* Row-band sharding for multi-GPU (partitioning 1): the image is cut into bands of 8 rows
 * and this handle renders band b iff b % stride == offset; all other pixels stay zero in every buffer,
 * so the sum (or gather) of the shards' buffers is bit-identical to the single-GPU result. Default (0, 1). */
TB_API int tb_set_row_shard(TbHandle* h, uint32_t offset, uint32_t stride);
TB_API int tb_buffer_size(TbHandle* h, uint32_t kind, uint64_t* bytes);
TB_API int tb_readback(TbHandle* h, uint32_t kind, void* dst, uint64_t bytes);
/* Device pointer of the float4 accumulation buffer (for the NCCL reduce; the
 * pointer stays valid until tb_resize/tb_destroy). */
TB_API int tb_device_buffer(TbHandle* h, uint32_t kind, void** devPtr, uint64_t* bytes);
TB_API int tb_get_render_stats(TbHandle* h, TbRenderStats* out);
TB_API int tb_reset_render_stats(TbHandle* h);
/* Profiling mode: record CUDA events around every traversal / shading launch (small
 * overhead; off by default). Results appear in TbRenderStats.Extend/ShadeMilliseconds.
 *   1  one frame at a time: exclusive device time of every launch (what a kernel costs alone)
 *   2  frames in flight as in a normal render (no graph replay): each launch's elapsed time on its own stream while
 *      the other frames' kernels share the GPU, i.e. the kernels' SHARES of the step as it is really run */
TB_API int tb_set_profiling(TbHandle* h, int enable);
/* Number of frames (samples) kept in flight on independent CUDA streams (0 = automatic:
 * as many as fit a third of the free device memory, between 4 and 16). The result does not depend on it: samples are added to the accumulation buffer in frame order. */
TB_API int tb_set_frames_in_flight(TbHandle* h, uint32_t n);
/* Next-event shadow rays: 0 = traced inline in the shading kernel, 1 = own wavefront stage
 * (shade A -> shadow traversal -> shade B), 2 = automatic (default; by triangle count). Scheduling
 * only: results are identical. */
TB_API int tb_set_shadow_mode(TbHandle* h, int mode);
/* Spatial sort of the ray queues before traversal (counting sort by the Morton cell of the ray origin): 0 = off,
 * 1 = the bounce queue, 3 = bounce and shadow queues, 4 = automatic (default: 3 for scenes whose traversal BVH is
 * beyond 256 MB, i.e. far outside L2, else 0). Scheduling only: results are identical. */
TB_API int tb_set_ray_sort(TbHandle* h, int mode);
/* Hit queue of the shading stage grouped by material class (the material's flag bits + "albedo is textured") so that a
 * warp shades hits that take the same branches: 0 = off, 1 = on, 2 = automatic (default: on when the scene's reachable
 * materials span four or more classes, where it was measured to pay). Scheduling only: results are identical. */
TB_API int tb_set_material_sort(TbHandle* h, int mode);
TB_API int tb_synchronize(TbHandle* h);

/* --------------------------------------------------------------- multi-GPU */
/* One process per GPU, one handle per process (the reference has no multi-GPU path; SURVEY §8b asks for "one CUDA
 * stream per device + one NCCL communicator" behind the boundary). Rank 0 obtains an id (ncclGetUniqueId) and hands
 * its TB_COMM_ID_BYTES bytes to the other processes by whatever means the host has (MPI, torch.distributed, a file);
 * every process then calls tb_comm_init on its handle (ncclCommInitRank on the handle's device), which also sets the
 * handle's shard (tb_set_frame_shard / tb_set_row_shard) to (rank, nranks). Scene and BVH are replicated: every rank
 * loads the same scene; the build is deterministic. NCCL is loaded at run time (libnccl.so.2; TB_NCCL_LIB overrides).
 * On one node the ranks map each other's accumulation buffers (CUDA IPC) and the reduction is one kernel per rank over
 * peer memory (TbCommInfo::Transport == TB_COMM_TRANSPORT_PEER; TB_COMM_TRANSPORT=nccl in the environment forces the
 * all-gather transport, which is also the fallback when a mapping cannot be opened). While buffers are mapped, the calls
 * that free them - tb_resize, tb_set_frames_in_flight, tb_comm_destroy, tb_destroy - are COLLECTIVE as well: every rank
 * makes them, and a rank waits (at most 10 s) for the others before it frees memory they may have open. */
TB_API int tb_comm_get_unique_id(void* id, uint64_t bytes);
TB_API int tb_comm_init(TbHandle* h, const void* id, int rank, int nranks, uint32_t shardMode);
TB_API int tb_comm_destroy(TbHandle* h);
TB_API int tb_comm_info(TbHandle* h, TbCommInfo* out);
/* The path's only exchange step, COLLECTIVE (every rank calls it): combines the ranks' float4 accumulation buffers
 * (OutputTexture and JitteredOutputTexture) into the job-wide image on every rank, deterministically (see
 * TB_SHARD_*). Afterwards tb_readback / tb_device_buffer / tb_save_image of TB_BUF_ACCUM_RGBW, TB_BUF_JITTERED_RGBW
 * and TB_BUF_RESOLVED_RGB, and tb_postprocess of the Lit output, serve the job-wide image. On a handle with a
 * communicator those calls run the reduction themselves when frames were rendered since the last one, which makes
 * THEM collective too. */
TB_API int tb_comm_reduce(TbHandle* h);

/* ------------------------------------------------------------ post-process */
/* The step right after the path: auto exposure (GenerateHistogramCS.hlsl,
 * CalculateAveragedLuminanceCS.hlsl; host side TracerBoy.cpp:2948-3039) and PostProcessCS.hlsl
 * (:3163-3199) on the buffer GetOutputSRV(OutputType) selects (TracerBoy.cpp:2354-2383): Lit / Luminance /
 * LiveWaves read the accumulation buffer, Albedo / LivePixels / Heatmap AOVCustomOutput, Normals AOVNormals,
 * Depth AOVDepth. MotionVectors and LuminanceVariance have no producer on this path -> TB_ERR_NOT_IMPL
 * (tb_postprocess_image accepts them). Results: TB_BUF_POSTPROCESS_RGBA, TB_BUF_BACKBUFFER_RGBA8,
 * TB_BUF_LUMINANCE_HISTOGRAM. */
TB_API int tb_get_default_postprocess_settings(TbPostProcessSettings* out);
TB_API int tb_postprocess(TbHandle* h, uint32_t outputType, const TbPostProcessSettings* s);
/* The same operator on caller-provided HOST images (float4 per pixel; aux may be NULL = zeros and is only read
 * by LiveWaves). outRGBA: float4 per pixel; outRGBA8 (may be NULL): 4 bytes per pixel; hist (may be NULL):
 * 256 words; avgLum (may be NULL): 1 float. */
/* Image files (the reference captures its back buffer as PNG, D3D12App.cpp:341-363). The extension
 * selects the format: ".png" writes an 8-bit buffer (TB_BUF_BACKBUFFER_RGBA8, i.e. after tb_postprocess);
 * ".exr" (OpenEXR scanline, uncompressed, 32-bit float) and ".pfm" write a float buffer: TB_BUF_RESOLVED_RGB,
 * TB_BUF_POSTPROCESS_RGBA, TB_BUF_ACCUM_RGBW or an AOV (float4 kinds keep their 4th channel as A in .exr). */
TB_API int tb_save_image(TbHandle* h, uint32_t kind, const char* path);
/* Host-only: decode a texture file the way the reference's loaders hand it to the GPU (TracerBoy::InitializeTexture,
 * TracerBoy.cpp:2186-2246; format rules of DirectXTex's WIC / TGA / HDR loaders): .png, .tga, .hdr. *format: 0 = float4
 * (16 bytes per pixel), 1 = RGBA8 UNORM, 2 = RGBA8 UNORM sRGB (4 bytes per pixel). pixels may be NULL to query the size;
 * otherwise capBytes must cover width * height * bytes per pixel. *hasAlpha = !IsAlphaAllOpaque(). */
TB_API int tb_load_image_file(const char* path, uint32_t* width, uint32_t* height, uint32_t* format, int* hasAlpha,
                              void* pixels, uint64_t capBytes, char* err, size_t errCap);
/* Host-only: the same writers on a caller-provided image (top row first; 4 x 8 bit for .png, 3 or 4 x float for .exr / .pfm). */
TB_API int tb_write_image(const char* path, const void* pixels, uint32_t width, uint32_t height, uint32_t channels,
                          uint32_t bytesPerChannel, char* err, size_t errCap);

/* Realtime temporal accumulation: TemporalAccumulationCS.hlsl:100-235 on caller-provided HOST
 * images, all float4 per pixel (history, current frame, world position of this and of the previous frame, normals,
 * moment history — the last may be NULL when OutputMomentInformation is 0). outColor: float4 (rgb, alpha = variance or 1);
 * outMoment (may be NULL): float4 (mean luminance, mean luminance^2, sample count, 0). */
TB_API int tb_temporal_accumulate_image(TbHandle* h, const TbTemporalAccumulationParams* p, uint32_t width, uint32_t height,
                                        const float* history, const float* current, const float* worldPos,
                                        const float* prevWorldPos, const float* normals, const float* momentHistory,
                                        float* outColor, float* outMoment);
TB_API int tb_postprocess_image(TbHandle* h, const float* inRGBA, const float* auxRGBA, uint32_t width, uint32_t height,
                                uint32_t outputType, const TbPostProcessSettings* s, float* outRGBA, uint8_t* outRGBA8,
                                uint32_t* hist, float* avgLum);

/* --------------------------------------------------------------- materials */
TB_API int tb_is_material_id_valid(TbHandle* h, int id);                 /* IsMaterialIDValid */
TB_API int tb_get_material(TbHandle* h, int id, TbMaterial* out, char* name, uint32_t nameCap); /* GetMaterial */
TB_API int tb_set_material(TbHandle* h, int id, const TbMaterial* m);    /* SetMaterial (invalidates history) */

/* ------------------------------------------------- SW-RT layer equivalents */
TB_API int tb_bvh_prebuild_info(const TbGeometryDesc* geoms, uint32_t n, TbPrebuildInfo* out);
/* BuildRaytracingAccelerationStructure on caller-provided geometry; replaces
 * the handle's scene geometry (materials: one default). */
TB_API int tb_bvh_build(TbHandle* h, const TbGeometryDesc* geoms, uint32_t n, uint32_t bvhBuildFlags);
/* SoftwareRayQuery::TraceRayInline + Proceed for n rays (host arrays). */
TB_API int tb_trace_rays(TbHandle* h, const TbRay* rays, uint64_t n, TbHit* hits);
/* The same two calls with the reference's ownership model (D3D12RaytracingFallback.h:83-84, 134-136:
 * BuildRaytracingAccelerationStructure(desc) with caller-allocated DestAccelerationStructureData and
 * ScratchAccelerationStructureData GPU VAs). Every pointer inside `geoms` (Positions, Indices, Transform3x4) is a
 * DEVICE pointer on the handle's device; `dst` / `scratch` are caller-owned device memory sized by
 * tb_bvh_prebuild_info (scratch may be NULL: the library then allocates its own for the duration of the call);
 * `cudaStream` is a cudaStream_t (NULL = the handle's stream). The build runs on that stream and the call returns
 * when it has completed. The handle's scene is not touched. Indices are not range-checked (as in the reference). */
TB_API int tb_bvh_build_device(TbHandle* h, const TbGeometryDesc* geoms, uint32_t n, uint32_t bvhBuildFlags, void* dst,
                               uint64_t dstBytes, void* scratch, uint64_t scratchBytes, void* cudaStream);
/* n ray queries against a caller-owned acceleration structure (`as` = a dst of tb_bvh_build_device, also one built
 * by another handle on the same device; NULL = the handle's scene), rays and hits in DEVICE memory. Enqueued on
 * `cudaStream` and NOT synchronised: it orders with the caller's other work on that stream like a dispatch. */
TB_API int tb_trace_rays_device(TbHandle* h, const void* as, uint64_t asBytes, const TbRay* dRays, uint64_t n, TbHit* dHits,
                                void* cudaStream);
/* BuildRaytracingAccelerationStructure with PERFORM_UPDATE (GpuBVH2Builder.cpp:165-234; SURVEY 8f rank 4: animated
 * geometry): `dst` is an acceleration structure built by tb_bvh_build_device, `geoms` describes the same geometries
 * (same triangle counts, same index topology) with moved vertices. The hierarchy is kept, the triangles are reloaded
 * into their sorted slots and every box is refitted bottom-up, in place; scratch is UpdateScratchDataSizeInBytes of
 * tb_bvh_prebuild_info (NULL: library-owned for the call). No ALLOW_UPDATE flag is needed at build time: the slot of a
 * triangle is recovered from the structure's own metadata instead of the reference's sort cache. */
TB_API int tb_bvh_update_device(TbHandle* h, const TbGeometryDesc* geoms, uint32_t n, void* dst, uint64_t dstBytes,
                                void* scratch, uint64_t scratchBytes, void* cudaStream);
/* Top-level acceleration structures (D3D12_RAYTRACING_ACCELERATION_STRUCTURE_TYPE_TOP_LEVEL,
 * GpuBVH2Builder.cpp:116-146) over instances of bottom-level structures built by tb_bvh_build_device, and the two-level
 * ray query (TraverseFunction.hlsli with FAST_PATH 0: the ray is taken to object space at an instance leaf, :603-638).
 * `instances` is a HOST array (the reference takes a GPU address of an upload-heap buffer); it is uploaded and the top
 * level is built on the GPU with the bottom level's pipeline minus the treelet pass. `dst` and `scratch` are caller-owned
 * device memory of ResultDataMaxSizeInBytes / ScratchDataSizeInBytes (tb_tlas_prebuild_info), 256-byte aligned, as in
 * BuildRaytracingAccelerationStructure's DestAccelerationStructureData / ScratchAccelerationStructureData. The result starts with the reference's byte layout (header,
 * 32-byte nodes, 116-byte BVHMetadata per sorted leaf: RayTracingHlslCompat.h:226-236). Hit records carry
 * InstanceIndex; instances with InstanceMask 0 are never hit; instance flags other than 0 are not interpreted (the path
 * uses RAY_FLAG_NONE and no any-hit). TracerBoy's software path itself stays single-level (FAST_PATH 1). */
TB_API int tb_tlas_prebuild_info(uint32_t numInstances, TbPrebuildInfo* out);
TB_API int tb_tlas_build_device(TbHandle* h, const TbInstanceDesc* instances, uint32_t n, uint32_t bvhBuildFlags, void* dst,
                                uint64_t dstBytes, void* scratch, uint64_t scratchBytes, void* cudaStream);
TB_API int tb_trace_rays_tlas_device(TbHandle* h, const void* tlas, uint64_t tlasBytes, const TbRay* dRays, uint64_t n, TbHit* dHits,
                                     void* cudaStream);
/* Drop the handle's cached description of a caller-owned acceleration structure (call before freeing / reusing dst). */
TB_API int tb_bvh_forget_device(TbHandle* h, const void* as);
/* Height of the scene's BVH (0 = a single leaf). The traversal keeps at most one waiting node per level; a tree deeper
 * than its 96-entry stack is rejected at build time (TB_ERR_NOT_IMPL) rather than traversed lossily. */
TB_API int tb_get_bvh_depth(TbHandle* h, uint32_t* depth);

#ifdef __cplusplus
}
static_assert(sizeof(TbMaterial) == 84, "Material must be 84 bytes");
static_assert(sizeof(TbLight) == 104, "Light must be 104 bytes");
static_assert(sizeof(TbTextureData) == 80, "TextureData must be 80 bytes");
static_assert(sizeof(TbVertex) == 32, "Vertex must be 32 bytes");
static_assert(sizeof(TbRay) == 32, "Ray must be 32 bytes");
static_assert(sizeof(TbInstanceDesc) == 64, "D3D12_RAYTRACING_INSTANCE_DESC is 64 bytes");
static_assert(sizeof(TbHit) == 32, "Hit must be 32 bytes");
#endif
#endif /* TRACERBOY_B200_H */
