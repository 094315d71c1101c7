"""ctypes mirror of `class TracerBoy` (TracerBoy/TracerBoy.h:158-397) over the C ABI.

Method names follow the reference (LoadScene, Render, SetMaterial, ...). Everything heavy
happens in libtracerboy_b200.so; numpy is only used to hand out readback buffers.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class TracerBoyError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("tracerboy_b200 error %d: %s" % (code, message))
        self.code = code


class Float3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]

    def tuple(self):
        return (self.x, self.y, self.z)


class Material(C.Structure):  # SharedShaderStructs.h:141-161
    _fields_ = [("albedo", Float3), ("albedoIndex", C.c_uint32), ("alphaIndex", C.c_uint32),
                ("normalMapIndex", C.c_uint32), ("emissiveIndex", C.c_uint32), ("specularMapIndex", C.c_uint32),
                ("IOR", C.c_float), ("absorption", Float3), ("roughness", C.c_float), ("scattering", Float3),
                ("emissive", Float3), ("Flags", C.c_int32), ("SpecularCoef", C.c_float)]


class Camera(C.Structure):  # TracerBoy.h:59-67
    _fields_ = [("Position", Float3), ("LookAt", Float3), ("Right", Float3), ("Up", Float3),
                ("LensHeight", C.c_float), ("FocalDistance", C.c_float)]


class ControllerState(C.Structure):  # TracerBoy.h:69-77
    _fields_ = [("RightStickX", C.c_float), ("RightStickY", C.c_float), ("RightTrigger", C.c_float),
                ("LeftStickX", C.c_float), ("LeftStickY", C.c_float), ("LeftTrigger", C.c_float)]


class CameraSettings(C.Structure):  # TracerBoy::CameraSettings, TracerBoy.h:386-390
    _fields_ = [("MovementSpeed", C.c_float), ("IgnoreMouse", C.c_uint32)]


class OutputSettings(C.Structure):  # TracerBoy.h:212-288 (members that reach PerFrameConstants)
    _fields_ = [("OutputType", C.c_uint32), ("EnableNormalMaps", C.c_uint32), ("RenderMode", C.c_uint32),
                ("SampleLimit", C.c_int32), ("TimeLimitInSeconds", C.c_float), ("DebugValue", C.c_float),
                ("DebugValue2", C.c_float), ("DOFFocalDistance", C.c_float), ("ApertureWidth", C.c_float),
                ("FilterType", C.c_uint32), ("FilterWidth", C.c_float), ("FireflyClampValue", C.c_float),
                ("MaxZ", C.c_float), ("ConvergencePercentage", C.c_float), ("EnableNextEventEstimation", C.c_uint32),
                ("EnableSamplingImportanceResampling", C.c_uint32), ("EnableBlueNoise", C.c_uint32),
                ("MaxBounces", C.c_int32)]


class PostProcessSettings(C.Structure):  # TracerBoy.h:222-229 as they reach PostProcessConstants (SharedPostProcessStructs.h:3-14)
    _fields_ = [("ExposureMultiplier", C.c_float), ("TonemapType", C.c_uint32), ("UseGammaCorrection", C.c_uint32),
                ("UseAutoExposure", C.c_uint32), ("VarianceMultiplier", C.c_float)]


class TemporalAccumulationParams(C.Structure):  # TemporalAccumulationSharedShaderStructs.h:6-34
    _fields_ = [("Camera", Camera), ("PrevCamera", Camera), ("HistoryWeight", C.c_float), ("IgnoreHistory", C.c_uint32),
                ("OutputMomentInformation", C.c_uint32)]


class OutputType:  # TracerBoy.h:171-183
    LIT, ALBEDO, NORMALS, DEPTH, MOTION_VECTORS, LUMINANCE, LUMINANCE_VARIANCE, LIVE_PIXELS, LIVE_WAVES, HEATMAP = range(10)


class TonemapType:  # Tonemap.h:3-10
    REINHARD, ACES, CLAMP, UNCHARTED, KHRONOS_PBR_NEUTRAL, AGX, AGX_PUNCHY, GT = range(8)


class Ray(C.Structure):
    _fields_ = [("Origin", C.c_float * 3), ("TMin", C.c_float), ("Direction", C.c_float * 3), ("TMax", C.c_float)]


class Hit(C.Structure):
    _fields_ = [("t", C.c_float), ("b1", C.c_float), ("b2", C.c_float), ("PrimitiveIndex", C.c_uint32),
                ("GeometryIndex", C.c_uint32), ("InstanceIndex", C.c_uint32), ("TrianglesTested", C.c_uint32),
                ("BoxesTested", C.c_uint32)]


RAY_DTYPE = np.dtype([("Origin", np.float32, 3), ("TMin", np.float32), ("Direction", np.float32, 3), ("TMax", np.float32)])
HIT_DTYPE = np.dtype([("t", np.float32), ("b1", np.float32), ("b2", np.float32), ("PrimitiveIndex", np.uint32),
                      ("GeometryIndex", np.uint32), ("InstanceIndex", np.uint32), ("TrianglesTested", np.uint32),
                      ("BoxesTested", np.uint32)])


class RenderStats(C.Structure):
    _fields_ = [("RaysTraced", C.c_uint64), ("BoxesTested", C.c_uint64), ("TrianglesTested", C.c_uint64),
                ("PathsStarted", C.c_uint64), ("KernelLaunches", C.c_uint64), ("DeviceMilliseconds", C.c_double),
                ("ExtendRays", C.c_uint64), ("ExtendBoxesTested", C.c_uint64), ("ExtendTrianglesTested", C.c_uint64),
                ("ExtendLaunches", C.c_uint64), ("ExtendMilliseconds", C.c_double), ("ShadeMilliseconds", C.c_double),
                ("ResumeRays", C.c_uint64), ("ResumeBoxesTested", C.c_uint64), ("ResumeTrianglesTested", C.c_uint64),
                ("ResumeMilliseconds", C.c_double), ("RaysByBounce", C.c_uint64 * 32),
                ("BounceMilliseconds", C.c_double * 32), ("BounceExtendMilliseconds", C.c_double * 32)]


class SceneInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("NumGeometries", "NumTriangles", "NumVertices", "NumMaterials",
                                          "NumLights", "NumTextures", "NumImages", "HasEnvironmentMap")]


class ReadbackStats(C.Structure):
    _fields_ = [("ActiveWaves", C.c_uint32), ("ActivePixels", C.c_uint32), ("SelectedPixelDistance", C.c_float),
                ("SelectedMaterialID", C.c_int32)]


class LoadStatus(C.Structure):
    _fields_ = [("State", C.c_uint32), ("InstancesLoaded", C.c_uint32), ("TotalInstances", C.c_uint32)]


class GeometryDesc(C.Structure):
    _fields_ = [("Positions", C.c_void_p), ("PositionStrideBytes", C.c_uint32), ("VertexCount", C.c_uint32),
                ("Indices", C.c_void_p), ("IndexFormat", C.c_uint32), ("IndexCount", C.c_uint32),
                ("Transform3x4", C.c_void_p), ("GeometryFlags", C.c_uint32)]


class InstanceDesc(C.Structure):
    """D3D12_RAYTRACING_INSTANCE_DESC (64 bytes)."""
    _fields_ = [("Transform", C.c_float * 12), ("InstanceIDAndMask", C.c_uint32),
                ("InstanceContributionToHitGroupIndexAndFlags", C.c_uint32), ("AccelerationStructure", C.c_uint64)]


class PrebuildInfo(C.Structure):
    _fields_ = [("ResultDataMaxSizeInBytes", C.c_uint64), ("ScratchDataSizeInBytes", C.c_uint64),
                ("UpdateScratchDataSizeInBytes", C.c_uint64), ("ReferenceLayoutSizeInBytes", C.c_uint64)]


class CommInfo(C.Structure):
    _fields_ = [("Rank", C.c_uint32), ("NumRanks", C.c_uint32), ("ShardMode", C.c_uint32), ("NcclVersion", C.c_uint32),
                ("Reductions", C.c_uint64), ("BytesReceivedPerReduction", C.c_uint64), ("LastReductionMilliseconds", C.c_double),
                ("TotalReductionMilliseconds", C.c_double), ("Transport", C.c_uint32), ("Reserved", C.c_uint32)]


SHARD_SAMPLES, SHARD_ROWS = 1, 2
COMM_TRANSPORT_NCCL, COMM_TRANSPORT_PEER = 0, 1
COMM_ID_BYTES = 128


class BufferKind:
    ACCUM_RGBW, JITTERED_RGBW, RESOLVED_RGB, AOV_NORMAL, AOV_WORLDPOS, AOV_DEPTH, AOV_ALBEDO, AOV_EMISSIVE, \
        PRIMARY_HIT_IDS, RAY_COUNTERS, POSTPROCESS_RGBA, BACKBUFFER_RGBA8, LUMINANCE_HISTOGRAM, LOCAL_ACCUM_RGBW = range(14)
    _shape = {0: (np.float32, 4), 1: (np.float32, 4), 2: (np.float32, 3), 3: (np.float32, 4), 4: (np.float32, 4),
              5: (np.float32, 1), 6: (np.float32, 4), 7: (np.float32, 4), 8: (np.uint32, 2), 9: (np.uint32, 2),
              10: (np.float32, 4), 11: (np.uint8, 4), 13: (np.float32, 4)}


BVH_BUILD_PREFER_FAST_TRACE = 0x4
BVH_BUILD_PREFER_FAST_BUILD = 0x8

_lib = None


def lib_path():
    # TB_LIB: an alternative build of the same library (tuning experiments, e.g. other launch bounds)
    return os.environ.get("TB_LIB") or os.path.join(_HERE, "lib", "libtracerboy_b200.so")


def load_library():
    """Load the CUDA library. Fails loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise TracerBoyError(-5, "CUDA library %s is missing; run `python -m tracerboy_b200.build`" % p)
    lib = C.CDLL(p)
    lib.tb_last_error.restype = C.c_char_p
    lib.tb_last_error.argtypes = [C.c_void_p]
    lib.tb_version.restype = C.c_char_p
    lib.tb_destroy.restype = None
    lib.tb_destroy.argtypes = [C.c_void_p]
    lib.tb_max_triangles.restype = C.c_uint64
    lib.tb_max_triangles.argtypes = []
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    sig = {
        "tb_create": [i32, C.POINTER(vp)], "tb_load_scene": [vp, C.c_char_p], "tb_load_scene_ex": [vp, C.c_char_p, u32],
        "tb_get_load_status": [vp, C.POINTER(LoadStatus)], "tb_save_scene": [vp, C.c_char_p],
        "tb_convert_scene": [C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t],
        "tb_convert_scene_ex": [C.c_char_p, C.c_char_p, u32, C.c_char_p, C.c_size_t], "tb_set_instance_mode": [vp, u32],
        "tb_get_scene_info": [vp, C.POINTER(SceneInfo)], "tb_get_bvh_size": [vp, C.POINTER(u64)],
        "tb_get_bvh": [vp, vp, u64], "tb_get_bvh_build_ms": [vp, C.POINTER(C.c_double)],
        "tb_get_default_settings": [C.POINTER(OutputSettings)], "tb_get_camera": [vp, C.POINTER(Camera)],
        "tb_set_camera": [vp, C.POINTER(Camera)], "tb_resize": [vp, u32, u32], "tb_select_pixel": [vp, i32, i32],
        "tb_update": [vp, i32, i32, C.c_void_p, C.c_float, C.POINTER(ControllerState), C.POINTER(CameraSettings)],
        "tb_camera_update": [C.POINTER(Camera), C.POINTER(C.c_uint32), u32, u32, i32, i32, C.c_void_p, C.c_float,
                             C.POINTER(ControllerState), C.POINTER(CameraSettings), C.POINTER(C.c_int)],
        "tb_get_stats": [vp, C.POINTER(ReadbackStats)],
        "tb_render": [vp, C.POINTER(OutputSettings), u32, C.c_float], "tb_samples_rendered": [vp, C.POINTER(u32)],
        "tb_invalidate_history": [vp], "tb_set_frame_shard": [vp, u32, u32], "tb_set_row_shard": [vp, u32, u32], "tb_buffer_size": [vp, u32, C.POINTER(u64)],
        "tb_readback": [vp, u32, vp, u64], "tb_device_buffer": [vp, u32, C.POINTER(vp), C.POINTER(u64)],
        "tb_get_render_stats": [vp, C.POINTER(RenderStats)], "tb_reset_render_stats": [vp], "tb_synchronize": [vp], "tb_set_profiling": [vp, i32], "tb_set_frames_in_flight": [vp, u32], "tb_set_shadow_mode": [vp, i32], "tb_set_ray_sort": [vp, i32], "tb_set_material_sort": [vp, i32],
        "tb_is_material_id_valid": [vp, i32], "tb_get_material": [vp, i32, C.POINTER(Material), C.c_char_p, u32],
        "tb_set_material": [vp, i32, C.POINTER(Material)],
        "tb_bvh_prebuild_info": [C.POINTER(GeometryDesc), u32, C.POINTER(PrebuildInfo)],
        "tb_bvh_build": [vp, C.POINTER(GeometryDesc), u32, u32], "tb_trace_rays": [vp, vp, u64, vp],
        "tb_bvh_build_device": [vp, C.POINTER(GeometryDesc), u32, u32, vp, u64, vp, u64, vp],
        "tb_trace_rays_device": [vp, vp, u64, vp, u64, vp, vp], "tb_bvh_forget_device": [vp, vp],
        "tb_bvh_update_device": [vp, C.POINTER(GeometryDesc), u32, vp, u64, vp, u64, vp],
        "tb_tlas_prebuild_info": [u32, C.POINTER(PrebuildInfo)], "tb_tlas_build_device": [vp, C.POINTER(InstanceDesc), u32, u32, vp, u64, vp, u64, vp],
        "tb_trace_rays_tlas_device": [vp, vp, u64, vp, u64, vp, vp],
        "tb_get_bvh_depth": [vp, C.POINTER(u32)],
        "tb_comm_get_unique_id": [vp, u64], "tb_comm_init": [vp, vp, i32, i32, u32], "tb_comm_destroy": [vp],
        "tb_comm_info": [vp, C.POINTER(CommInfo)], "tb_comm_reduce": [vp],
        "tb_get_default_postprocess_settings": [C.POINTER(PostProcessSettings)],
        "tb_postprocess": [vp, u32, C.POINTER(PostProcessSettings)],
        "tb_temporal_accumulate_image": [vp, C.POINTER(TemporalAccumulationParams), u32, u32, vp, vp, vp, vp, vp, vp, vp, vp],
        "tb_save_image": [vp, u32, C.c_char_p],
        "tb_write_image": [C.c_char_p, vp, u32, u32, u32, u32, C.c_char_p, C.c_size_t],
        "tb_load_image_file": [C.c_char_p, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(C.c_int), vp, u64, C.c_char_p, C.c_size_t],
        "tb_postprocess_image": [vp, vp, vp, u32, u32, u32, C.POINTER(PostProcessSettings), vp, vp, vp, C.POINTER(C.c_float)],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _lib = lib
    return lib


EXPORTED_SYMBOLS = ["tb_create", "tb_destroy", "tb_last_error", "tb_version", "tb_load_scene", "tb_load_scene_ex",
                    "tb_get_load_status", "tb_save_scene", "tb_convert_scene", "tb_convert_scene_ex", "tb_set_instance_mode", "tb_get_scene_info", "tb_get_bvh_size",
                    "tb_get_bvh", "tb_get_bvh_build_ms", "tb_get_default_settings", "tb_get_camera", "tb_set_camera",
                    "tb_resize", "tb_select_pixel", "tb_get_stats", "tb_render", "tb_samples_rendered",
                    "tb_invalidate_history", "tb_set_frame_shard", "tb_set_row_shard", "tb_buffer_size", "tb_readback", "tb_device_buffer",
                    "tb_get_render_stats", "tb_reset_render_stats", "tb_set_profiling", "tb_set_frames_in_flight", "tb_set_shadow_mode", "tb_synchronize", "tb_is_material_id_valid",
                    "tb_get_material", "tb_set_material", "tb_bvh_prebuild_info", "tb_bvh_build", "tb_trace_rays",
                    "tb_get_default_postprocess_settings", "tb_postprocess", "tb_postprocess_image",
                    "tb_temporal_accumulate_image", "tb_save_image", "tb_write_image", "tb_update", "tb_camera_update", "tb_set_ray_sort",
                    "tb_set_material_sort", "tb_load_image_file", "tb_max_triangles", "tb_bvh_build_device", "tb_trace_rays_device", "tb_bvh_forget_device", "tb_bvh_update_device", "tb_tlas_prebuild_info", "tb_tlas_build_device", "tb_trace_rays_tlas_device", "tb_get_bvh_depth",
                    "tb_comm_get_unique_id", "tb_comm_init", "tb_comm_destroy", "tb_comm_info", "tb_comm_reduce"]


def _key_table(keyboardInput):
    """bool keyboardInput[CHAR_MAX] of TracerBoy::Update from an iterable of pressed characters (or a ready table)."""
    if keyboardInput is None:
        return None
    table = (C.c_uint8 * 127)()
    if isinstance(keyboardInput, (bytes, bytearray)) and len(keyboardInput) == 127:
        for i, v in enumerate(keyboardInput):
            table[i] = 1 if v else 0
    else:
        for ch in keyboardInput:
            table[ord(ch) if isinstance(ch, str) else int(ch)] = 1
    return C.cast(table, C.c_void_p)


def camera_update(camera, last_mouse, width, height, mouseX, mouseY, keyboardInput=None, dt=0.0, controllerState=None,
                  cameraSettings=None):
    """tb_camera_update: TracerBoy::Update on caller-owned state (host only). Returns bCameraMoved; camera and
    last_mouse (a 2-element ctypes c_uint32 array) are updated in place."""
    moved = C.c_int(0)
    rc = load_library().tb_camera_update(C.byref(camera), last_mouse, int(width), int(height), int(mouseX), int(mouseY),
                                         _key_table(keyboardInput), float(dt),
                                         C.byref(controllerState) if controllerState is not None else None,
                                         C.byref(cameraSettings) if cameraSettings is not None else None, C.byref(moved))
    if rc != 0:
        raise TracerBoyError(rc, "tb_camera_update")
    return bool(moved.value)


def get_default_output_settings():
    """TracerBoy::GetDefaultOutputSettings (TracerBoy.h:290-360)."""
    s = OutputSettings()
    load_library().tb_get_default_settings(C.byref(s))
    return s


def get_default_postprocess_settings():
    """PostProcessSettings of TracerBoy::GetDefaultOutputSettings (TracerBoy.h:308-313)."""
    s = PostProcessSettings()
    load_library().tb_get_default_postprocess_settings(C.byref(s))
    return s


def write_image(path, pixels):
    """Host-only: write an (h, w, 4) uint8 array as .png, or an (h, w, 3|4) float32 array as .exr / .pfm."""
    a = np.ascontiguousarray(pixels)
    err = C.create_string_buffer(512)
    rc = load_library().tb_write_image(str(path).encode(), a.ctypes.data, a.shape[1], a.shape[0], a.shape[2], a.dtype.itemsize, err, 512)
    if rc != 0:
        raise TracerBoyError(rc, err.value.decode())


def comm_get_unique_id():
    """ncclGetUniqueId: called by one process, the bytes go to every rank's CommInit by whatever means the host has."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    lib = load_library()
    rc = lib.tb_comm_get_unique_id(buf, COMM_ID_BYTES)
    if rc != 0:
        raise TracerBoyError(rc, lib.tb_last_error(None).decode())
    return buf.raw


def prebuild_info(descs, n):
    """GetRaytracingAccelerationStructurePrebuildInfo on a (GeometryDesc * n) array."""
    info = PrebuildInfo()
    rc = load_library().tb_bvh_prebuild_info(descs, n, C.byref(info))
    if rc != 0:
        raise TracerBoyError(rc, "tb_bvh_prebuild_info")
    return info


def load_image_file(path):
    """Host-only: decode a .png / .tga / .hdr texture as the reference's loaders would hand it to the GPU.
    Returns (pixels, format, has_alpha): float32 [h, w, 4] for format 0, uint8 [h, w, 4] for 1 (UNORM) and 2 (UNORM sRGB)."""
    lib = load_library()
    w, h, fmt, alpha = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_int()
    err = C.create_string_buffer(512)
    rc = lib.tb_load_image_file(str(path).encode(), C.byref(w), C.byref(h), C.byref(fmt), C.byref(alpha), None, 0, err, 512)
    if rc != 0:
        raise TracerBoyError(rc, err.value.decode())
    out = np.empty((h.value, w.value, 4), np.float32 if fmt.value == 0 else np.uint8)
    rc = lib.tb_load_image_file(str(path).encode(), C.byref(w), C.byref(h), C.byref(fmt), C.byref(alpha), out.ctypes.data, out.nbytes, err, 512)
    if rc != 0:
        raise TracerBoyError(rc, err.value.decode())
    return out, fmt.value, bool(alpha.value)


def tlas_prebuild_info(n):
    info = PrebuildInfo()
    rc = load_library().tb_tlas_prebuild_info(n, C.byref(info))
    if rc != 0:
        raise TracerBoyError(rc, "tb_tlas_prebuild_info")
    return info


INSTANCES_SKIP, INSTANCES_INSERT_INTO_BLAS = 0, 1


def convert_scene(src, dst, instance_mode=INSTANCES_SKIP):
    """Host-only: import a scene (.pbrt/.pbf/.tbscene/synthetic:) and write the .tbscene cache."""
    err = C.create_string_buffer(1024)
    rc = load_library().tb_convert_scene_ex(src.encode(), dst.encode(), instance_mode, err, 1024)
    if rc != 0:
        raise TracerBoyError(rc, err.value.decode())


class TracerBoy:
    """Mirror of `class TracerBoy`: ctor / LoadScene / Render / readback / materials."""

    def __init__(self, device=0):
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.tb_create(int(device), C.byref(h))
        if rc != 0:
            raise TracerBoyError(rc, self._lib.tb_last_error(None).decode())
        self._h = h
        self.width = self.height = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.tb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise TracerBoyError(rc, self._lib.tb_last_error(self._h).decode())

    # --- scene -----------------------------------------------------------
    def LoadScene(self, path, bvh_build_flags=BVH_BUILD_PREFER_FAST_TRACE):
        self._ck(self._lib.tb_load_scene_ex(self._h, path.encode(), bvh_build_flags))

    def SetInstanceMode(self, mode):
        """INSTANCES_SKIP (the reference's software path) or INSTANCES_INSERT_INTO_BLAS (LoadScene's bInsertInstancesIntoBLAS)."""
        self._ck(self._lib.tb_set_instance_mode(self._h, mode))

    def SaveScene(self, path):
        self._ck(self._lib.tb_save_scene(self._h, path.encode()))

    def GetSceneLoadStatus(self):
        s = LoadStatus()
        self._ck(self._lib.tb_get_load_status(self._h, C.byref(s)))
        return s

    def GetSceneInfo(self):
        s = SceneInfo()
        self._ck(self._lib.tb_get_scene_info(self._h, C.byref(s)))
        return s

    def GetBVH(self):
        n = C.c_uint64()
        self._ck(self._lib.tb_get_bvh_size(self._h, C.byref(n)))
        buf = np.empty(n.value, np.uint8)
        self._ck(self._lib.tb_get_bvh(self._h, buf.ctypes.data, n.value))
        return buf

    def GetBVHBuildMilliseconds(self):
        ms = C.c_double()
        self._ck(self._lib.tb_get_bvh_build_ms(self._h, C.byref(ms)))
        return ms.value

    def BuildRaytracingAccelerationStructure(self, geometries, flags=BVH_BUILD_PREFER_FAST_TRACE):
        """geometries: list of (positions float32 [V,3], indices uint32/uint16 [T*3] or None)."""
        descs = (GeometryDesc * len(geometries))()
        keep = []
        for i, (pos, idx) in enumerate(geometries):
            pos = np.ascontiguousarray(pos, np.float32)
            keep.append(pos)
            descs[i].Positions = pos.ctypes.data
            descs[i].PositionStrideBytes = 12
            descs[i].VertexCount = pos.shape[0]
            descs[i].GeometryFlags = 1
            if idx is None:
                descs[i].Indices = None
                descs[i].IndexFormat = 0
                descs[i].IndexCount = 0
            else:
                idx = np.ascontiguousarray(idx)
                if idx.dtype not in (np.uint16, np.uint32):
                    idx = idx.astype(np.uint32)
                keep.append(idx)
                descs[i].Indices = idx.ctypes.data
                descs[i].IndexFormat = idx.dtype.itemsize
                descs[i].IndexCount = idx.size
        self._ck(self._lib.tb_bvh_build(self._h, descs, len(geometries), flags))

    def GetBVHDepth(self):
        d = C.c_uint32()
        self._ck(self._lib.tb_get_bvh_depth(self._h, C.byref(d)))
        return d.value

    def BuildRaytracingAccelerationStructureDevice(self, descs, n, dst, dst_bytes, scratch=None, scratch_bytes=0, stream=None,
                                                   flags=BVH_BUILD_PREFER_FAST_TRACE):
        """BuildRaytracingAccelerationStructure with caller-allocated dest + scratch device memory; `descs` is a
        (GeometryDesc * n) array whose pointers are DEVICE pointers (D3D12RaytracingFallback.h:83-84)."""
        self._ck(self._lib.tb_bvh_build_device(self._h, descs, n, flags, dst, dst_bytes, scratch, scratch_bytes, stream))

    def UpdateRaytracingAccelerationStructureDevice(self, descs, n, dst, dst_bytes, scratch=None, scratch_bytes=0, stream=None):
        """PERFORM_UPDATE: refit a caller-owned acceleration structure to moved vertices (same topology), in place."""
        self._ck(self._lib.tb_bvh_update_device(self._h, descs, n, dst, dst_bytes, scratch, scratch_bytes, stream))

    def BuildTopLevelAccelerationStructureDevice(self, instances, n, dst, dst_bytes, scratch, scratch_bytes, stream=None, flags=0):
        """TYPE_TOP_LEVEL build over a host (InstanceDesc * n) array whose AccelerationStructure fields are device
        addresses of bottom-level structures; dst / scratch = caller-owned device memory sized by tlas_prebuild_info."""
        self._ck(self._lib.tb_tlas_build_device(self._h, instances, n, flags, dst, dst_bytes, scratch, scratch_bytes, stream))

    def TraceRaysTopLevelDevice(self, tlas, tlas_bytes, d_rays, n, d_hits, stream=None):
        self._ck(self._lib.tb_trace_rays_tlas_device(self._h, tlas, tlas_bytes, d_rays, n, d_hits, stream))

    def TraceRaysDevice(self, accel, accel_bytes, d_rays, n, d_hits, stream=None):
        """n ray queries against a caller-owned acceleration structure (None: the handle's scene); device pointers,
        enqueued on `stream` without synchronising."""
        self._ck(self._lib.tb_trace_rays_device(self._h, accel, accel_bytes, d_rays, n, d_hits, stream))

    def ForgetAccelerationStructure(self, accel):
        self._ck(self._lib.tb_bvh_forget_device(self._h, accel))

    # --- multi-GPU -------------------------------------------------------
    def CommInit(self, unique_id, rank, nranks, shard_mode=SHARD_SAMPLES):
        """ncclCommInitRank on this handle's device; also sets the handle's shard to (rank, nranks)."""
        self._ck(self._lib.tb_comm_init(self._h, unique_id, int(rank), int(nranks), int(shard_mode)))

    def CommDestroy(self):
        self._ck(self._lib.tb_comm_destroy(self._h))

    def CommReduce(self):
        """Collective: the job-wide accumulation image on every rank (deterministic; see TB_SHARD_*)."""
        self._ck(self._lib.tb_comm_reduce(self._h))

    def CommInfo(self):
        s = CommInfo()
        self._ck(self._lib.tb_comm_info(self._h, C.byref(s)))
        return s

    def TraceRays(self, rays):
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        hits = np.empty(rays.shape[0], HIT_DTYPE)
        self._ck(self._lib.tb_trace_rays(self._h, rays.ctypes.data, rays.shape[0], hits.ctypes.data))
        return hits

    # --- camera / settings ----------------------------------------------
    @staticmethod
    def GetDefaultOutputSettings():
        return get_default_output_settings()

    def GetCamera(self):
        c = Camera()
        self._ck(self._lib.tb_get_camera(self._h, C.byref(c)))
        return c

    def SetCamera(self, cam):
        self._ck(self._lib.tb_set_camera(self._h, C.byref(cam)))

    def Resize(self, width, height):
        self._ck(self._lib.tb_resize(self._h, width, height))
        self.width, self.height = width, height

    def Update(self, mouseX, mouseY, keyboardInput=None, dt=0.0, controllerState=None, cameraSettings=None):
        """TracerBoy::Update (TracerBoy.cpp:3386-3500). keyboardInput: iterable of pressed characters or a 127-entry table."""
        self._ck(self._lib.tb_update(self._h, int(mouseX), int(mouseY), _key_table(keyboardInput), float(dt),
                                     C.byref(controllerState) if controllerState is not None else None,
                                     C.byref(cameraSettings) if cameraSettings is not None else None))

    def SelectPixel(self, x, y):
        self._ck(self._lib.tb_select_pixel(self._h, x, y))

    def GetReadbackStats(self):
        s = ReadbackStats()
        self._ck(self._lib.tb_get_stats(self._h, C.byref(s)))
        return s

    # --- render ----------------------------------------------------------
    def Render(self, settings=None, samples=1, time=0.0):
        """`samples` x the reference's one-sample Render (TracerBoy.cpp:2677-3369)."""
        if settings is None:
            settings = get_default_output_settings()
        self._ck(self._lib.tb_render(self._h, C.byref(settings), int(samples), C.c_float(time)))

    def GetNumberOfSamplesSinceLastInvalidate(self):
        n = C.c_uint32()
        self._ck(self._lib.tb_samples_rendered(self._h, C.byref(n)))
        return n.value

    def InvalidateHistory(self):
        self._ck(self._lib.tb_invalidate_history(self._h))

    def SetFrameShard(self, offset, stride):
        self._ck(self._lib.tb_set_frame_shard(self._h, offset, stride))

    def SetRowShard(self, offset, stride):
        self._ck(self._lib.tb_set_row_shard(self._h, offset, stride))

    def Readback(self, kind, out=None):
        dt, ch = BufferKind._shape[kind]
        shape = (self.height, self.width, ch) if ch > 1 else (self.height, self.width)
        if out is None:
            out = np.empty(shape, dt)
        self._ck(self._lib.tb_readback(self._h, kind, out.ctypes.data, out.nbytes))
        return out

    def PostProcess(self, output_type, settings):
        """Auto exposure + PostProcessCS on the buffer the output type selects (TracerBoy.cpp:2948-3039, 3163-3199).
        Results: Readback(POSTPROCESS_RGBA / BACKBUFFER_RGBA8), GetLuminanceHistogram()."""
        self._ck(self._lib.tb_postprocess(self._h, int(output_type), C.byref(settings)))

    def GetLuminanceHistogram(self):
        """(LuminanceHistogram[256], AveragedLuminance) of the last PostProcess call."""
        raw = np.empty(257, np.uint32)
        self._ck(self._lib.tb_readback(self._h, BufferKind.LUMINANCE_HISTOGRAM, raw.ctypes.data, raw.nbytes))
        return raw[:256].copy(), float(raw[256:].view(np.float32)[0])

    def PostProcessImage(self, img, output_type, settings, aux=None):
        """The same operator on a host float4 image: returns (float4 image, rgba8 image, histogram, averaged luminance)."""
        img = np.ascontiguousarray(img, np.float32)
        h, w = img.shape[:2]
        auxp = None
        if aux is not None:
            aux = np.ascontiguousarray(aux, np.float32)
            auxp = aux.ctypes.data
        out = np.empty((h, w, 4), np.float32)
        out8 = np.empty((h, w, 4), np.uint8)
        hist = np.zeros(256, np.uint32)
        avg = C.c_float(0.0)
        self._ck(self._lib.tb_postprocess_image(self._h, img.ctypes.data, auxp, w, h, int(output_type), C.byref(settings),
                                                out.ctypes.data, out8.ctypes.data, hist.ctypes.data, C.byref(avg)))
        return out, out8, hist, avg.value

    def TemporalAccumulateImage(self, params, history, current, world_pos, prev_world_pos, normals, moment_history=None):
        """TemporalAccumulationPass::Run on host float4 images: returns (color float4, moment float4 or None)."""
        imgs = [np.ascontiguousarray(a, np.float32) for a in (history, current, world_pos, prev_world_pos, normals)]
        h, w = imgs[0].shape[:2]
        mh = np.ascontiguousarray(moment_history, np.float32) if moment_history is not None else None
        out = np.empty((h, w, 4), np.float32)
        mom = np.empty((h, w, 4), np.float32) if params.OutputMomentInformation else None
        self._ck(self._lib.tb_temporal_accumulate_image(self._h, C.byref(params), w, h, *[a.ctypes.data for a in imgs],
                                                        mh.ctypes.data if mh is not None else None, out.ctypes.data,
                                                        mom.ctypes.data if mom is not None else None))
        return out, mom

    def SaveImage(self, kind, path):
        """Write a buffer as .png (BACKBUFFER_RGBA8), .exr or .pfm (float buffers)."""
        self._ck(self._lib.tb_save_image(self._h, int(kind), str(path).encode()))

    def DeviceBuffer(self, kind):
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(self._lib.tb_device_buffer(self._h, kind, C.byref(p), C.byref(n)))
        return p.value, n.value

    def GetRenderStats(self):
        s = RenderStats()
        self._ck(self._lib.tb_get_render_stats(self._h, C.byref(s)))
        return s

    def ResetRenderStats(self):
        self._ck(self._lib.tb_reset_render_stats(self._h))

    def SetFramesInFlight(self, n):
        self._ck(self._lib.tb_set_frames_in_flight(self._h, int(n)))

    def SetShadowMode(self, mode):
        self._ck(self._lib.tb_set_shadow_mode(self._h, int(mode)))

    def SetRaySort(self, mode):
        """0 off, 1 bounce queue, 3 bounce + shadow queues, 4 automatic (scheduling only, results are identical)."""
        self._ck(self._lib.tb_set_ray_sort(self._h, int(mode)))

    def SetMaterialSort(self, mode):
        """0 off, 1 on, 2 automatic: the shading stage's hit queue grouped by material class (scheduling only)."""
        self._ck(self._lib.tb_set_material_sort(self._h, int(mode)))

    def SetProfiling(self, enable):
        self._ck(self._lib.tb_set_profiling(self._h, int(enable)))

    def Synchronize(self):
        self._ck(self._lib.tb_synchronize(self._h))

    # --- materials -------------------------------------------------------
    def IsMaterialIDValid(self, i):
        return bool(self._lib.tb_is_material_id_valid(self._h, i))

    def GetMaterial(self, i):
        m = Material()
        name = C.create_string_buffer(64)
        self._ck(self._lib.tb_get_material(self._h, i, C.byref(m), name, 64))
        return m, name.value.decode()

    def SetMaterial(self, i, m):
        self._ck(self._lib.tb_set_material(self._h, i, C.byref(m)))
