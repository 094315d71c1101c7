"""tracerboy_b200 — B200-native implementation of TracerBoy's path-tracing hot path.

The product is the CUDA library behind the C ABI in include/tracerboy_b200.h; this package
is a thin ctypes mirror of the reference's `class TracerBoy` (TracerBoy/TracerBoy.h:158-397).
There is no CPU fallback: importing works anywhere, but every compute call raises
TracerBoyError when the CUDA library or a CUDA device is missing.
"""
import os as _os

# frames in flight use independent CUDA streams: ask for more hardware queues before CUDA starts
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .api import (TracerBoy, TracerBoyError, OutputSettings, Camera, Material, Ray, Hit, RenderStats,
                  SceneInfo, BufferKind, lib_path, load_library, convert_scene, get_default_output_settings,
                  PostProcessSettings, OutputType, TonemapType, get_default_postprocess_settings,
                  TemporalAccumulationParams, write_image, ControllerState, CameraSettings, camera_update,
                  comm_get_unique_id, prebuild_info, tlas_prebuild_info, InstanceDesc, load_image_file, SHARD_SAMPLES, SHARD_ROWS, GeometryDesc, PrebuildInfo,
                  INSTANCES_SKIP, INSTANCES_INSERT_INTO_BLAS)

__all__ = ["TracerBoy", "TracerBoyError", "OutputSettings", "Camera", "Material", "Ray", "Hit", "RenderStats",
           "SceneInfo", "BufferKind", "lib_path", "load_library", "convert_scene", "get_default_output_settings",
           "PostProcessSettings", "OutputType", "TonemapType", "get_default_postprocess_settings",
           "TemporalAccumulationParams", "write_image", "ControllerState", "CameraSettings", "camera_update",
           "comm_get_unique_id", "prebuild_info", "tlas_prebuild_info", "InstanceDesc", "load_image_file", "SHARD_SAMPLES", "SHARD_ROWS", "GeometryDesc", "PrebuildInfo",
           "INSTANCES_SKIP", "INSTANCES_INSERT_INTO_BLAS"]
