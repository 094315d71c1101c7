"""Builds oracle/_ref/libref_core.so: the reference's TracerBoy/kernel.glsl compiled as host C++
(TEST INFRASTRUCTURE). Needs the reference mount; outputs only into oracle/_ref/ (git-ignored,
travels to the GPU box as a binary). See oracle/ref/prepass.py and oracle/ref/ref_glue.cpp."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
KERNEL = "/root/reference/TracerBoy/kernel.glsl"
OUT = os.path.join(HERE, "_ref")
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def lib_path():
    return os.path.join(OUT, "libref_core.so")


def build(force=False):
    if not os.path.exists(KERNEL):
        return lib_path() if os.path.exists(lib_path()) else None
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE, "ref", f) for f in ("prepass.py", "hlsl_compat.h", "ref_glue.cpp")] + \
           [os.path.join(HERE, "glue.h"), os.path.join(HERE, "oracle.h"), KERNEL]
    target = lib_path()
    if not force and os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs) and \
            os.path.getmtime(target) >= os.path.getmtime(os.path.join(HERE, "liboracle.so")):
        return target
    sys.path.insert(0, os.path.join(HERE, "ref"))
    import prepass
    prepass.run(KERNEL, os.path.join(OUT, "kernel_gen.inc"))
    cmd = [GXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-mfma", "-ffp-contract=off",
           "-fsingle-precision-constant",  # HLSL literals are float
           "-fno-fast-math", "-fvisibility=hidden", "-w", "-I" + os.path.join(ROOT, "include"), "-I" + HERE,
           os.path.join(HERE, "ref", "ref_glue.cpp"), "-o", target, "-L" + HERE, "-loracle", "-Wl,-rpath,$ORIGIN/.."]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref build failed:\n" + r.stdout)
    return target


TONEMAP = "/root/reference/TracerBoy/Tonemap.h"
POSTPROCESS = "/root/reference/TracerBoy/PostProcessCS.hlsl"


def post_lib_path():
    return os.path.join(OUT, "libref_post.so")


def build_post(force=False):
    """oracle/_ref/libref_post.so: the reference's Tonemap.h + PostProcessCS.hlsl Process* functions as host C++."""
    target = post_lib_path()
    if not (os.path.exists(TONEMAP) and os.path.exists(POSTPROCESS)):
        return target if os.path.exists(target) else None
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE, "ref", f) for f in ("prepass.py", "hlsl_compat.h", "ref_post.cpp")] + [TONEMAP, POSTPROCESS]
    if not force and os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs):
        return target
    sys.path.insert(0, os.path.join(HERE, "ref"))
    import prepass
    prepass.run_post(TONEMAP, POSTPROCESS, os.path.join(OUT, "post_gen.inc"))
    cmd = [GXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-mfma", "-ffp-contract=off", "-fsingle-precision-constant",
           "-fno-fast-math", "-fvisibility=hidden", "-w", "-I" + os.path.join(ROOT, "include"), "-I" + HERE,
           os.path.join(HERE, "ref", "ref_post.cpp"), "-o", target]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref post-process build failed:\n" + r.stdout)
    return target


def treelet_pass_lib_path():
    return os.path.join(OUT, "libref_treelet_pass.so")


def build_treelet_pass(force=False):
    """oracle/_ref/libref_treelet_pass.so: ClearBuffers + FindTreelets + the whole TreeletReorder.hlsl as host C++."""
    src = "/root/reference/D3D12RaytracingFallback/src/"
    files = [src + f for f in ("TreeletReorderBindings.h", "TreeletReorder.hlsl", "ClearBuffers.hlsl", "FindTreelets.hlsl",
                               "RayTracingHelper.hlsli", "RayTracingHlslCompat.h")]
    target = treelet_pass_lib_path()
    if not all(os.path.exists(f) for f in files):
        return target if os.path.exists(target) else None
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE, "ref", f) for f in ("prepass.py", "hlsl_compat.h", "ref_treelet_pass.cpp")] + files
    if not force and os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs):
        return target
    sys.path.insert(0, os.path.join(HERE, "ref"))
    import prepass
    prepass.run_treelet_pass(*files, os.path.join(OUT, "treelet_pass_gen.inc"))
    cmd = [GXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-mfma", "-ffp-contract=off", "-fsingle-precision-constant", "-fno-fast-math",
           "-fvisibility=hidden", "-w", "-I" + os.path.join(ROOT, "include"), "-I" + HERE, os.path.join(HERE, "ref", "ref_treelet_pass.cpp"), "-o", target]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref treelet pass build failed:\n" + r.stdout)
    return target


def load_lib_path():
    return os.path.join(OUT, "libref_load.so")


def build_load(force=False):
    """oracle/_ref/libref_load.so: BottomLevelLoadTriangles.hlsli (three index-format variants) as host C++."""
    src = "/root/reference/D3D12RaytracingFallback/src/"
    files = [src + f for f in ("BottomLevelLoadTriangles.hlsli", "LoadPrimitivesBindings.h", "RayTracingHlslCompat.h")]
    target = load_lib_path()
    if not all(os.path.exists(f) for f in files):
        return target if os.path.exists(target) else None
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE, "ref", f) for f in ("prepass.py", "hlsl_compat.h", "ref_load.cpp")] + files
    if not force and os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs):
        return target
    sys.path.insert(0, os.path.join(HERE, "ref"))
    import prepass
    prepass.run_load(files[0], files[1], files[2], os.path.join(OUT, "load_common_gen.inc"), os.path.join(OUT, "load_variant_gen.inc"))
    cmd = [GXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-mfma", "-ffp-contract=off", "-fsingle-precision-constant", "-fno-fast-math",
           "-fvisibility=hidden", "-w", "-I" + os.path.join(ROOT, "include"), "-I" + HERE, os.path.join(HERE, "ref", "ref_load.cpp"), "-o", target]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref load build failed:\n" + r.stdout)
    return target


def refit_lib_path():
    return os.path.join(OUT, "libref_refit.so")


def build_refit(force=False):
    """oracle/_ref/libref_refit.so: PrepareForComputeAABBs + ComputeAABBs (node encoding, bottom-up climb) as host C++."""
    src = "/root/reference/D3D12RaytracingFallback/src/"
    files = [src + f for f in ("RayTracingHelper.hlsli", "BottomLevelPrepareForComputeAABBs.hlsl", "BottomLevelComputeAABBs.hlsl", "ComputeAABBs.hlsli")]
    target = refit_lib_path()
    if not all(os.path.exists(f) for f in files):
        return target if os.path.exists(target) else None
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE, "ref", f) for f in ("prepass.py", "hlsl_compat.h", "ref_refit.cpp")] + files
    if not force and os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs):
        return target
    sys.path.insert(0, os.path.join(HERE, "ref"))
    import prepass
    prepass.run_refit(files[0], files[1], files[2], files[3], os.path.join(OUT, "refit_prepare_gen.inc"), os.path.join(OUT, "refit_compute_gen.inc"))
    cmd = [GXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-mfma", "-ffp-contract=off", "-fsingle-precision-constant", "-fno-fast-math",
           "-fvisibility=hidden", "-w", "-I" + os.path.join(ROOT, "include"), "-I" + HERE, os.path.join(HERE, "ref", "ref_refit.cpp"), "-o", target]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref refit build failed:\n" + r.stdout)
    return target


def flatten_lib_path():
    return os.path.join(OUT, "libref_flatten.so")


def build_flatten(force=False):
    """oracle/_ref/libref_flatten.so: CreateMaterial, MaterialTracker and the area-light rule of LoadScene compiled from the
    mount, linked against the vendored pbrt-parser objects the importer build leaves in build/pbrt."""
    cpp, hdr, structs = ("/root/reference/TracerBoy/TracerBoy.cpp", "/root/reference/TracerBoy/TracerBoy.h", "/root/reference/TracerBoy/SharedShaderStructs.h")
    target = flatten_lib_path()
    parser = "/root/reference/PBRTParser"
    objdir = os.path.join(ROOT, "build", "pbrt")
    if not all(os.path.exists(f) for f in (cpp, hdr, structs)) or not os.path.isdir(objdir):
        return target if os.path.exists(target) else None
    objs = [os.path.join(objdir, f) for f in sorted(os.listdir(objdir)) if f.endswith(".o")]
    if not objs:
        return target if os.path.exists(target) else None
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE, "ref", f) for f in ("prepass.py", "ref_flatten.cpp")] + [cpp, hdr, structs]
    if not force and os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs):
        return target
    sys.path.insert(0, os.path.join(HERE, "ref"))
    import prepass
    prepass.run_flatten(cpp, hdr, structs, os.path.join(OUT, "flatten_gen.inc"), os.path.join(OUT, "flatten_light_gen.inc"))
    cmd = [GXX, "-O2", "-std=c++14", "-fPIC", "-shared", "-mfma", "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden", "-w", "-include", "cstdint",
           "-I" + os.path.join(parser, "include"), "-I" + HERE, os.path.join(HERE, "ref", "ref_flatten.cpp")] + objs + ["-o", target]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref flatten build failed:\n" + r.stdout)
    return target


def tlas_lib_path():
    return os.path.join(OUT, "libref_tlas.so")


def build_tlas(force=False):
    """oracle/_ref/libref_tlas.so: the arithmetic of the top-level instance load (inverse transform, transformed box) as host C++."""
    helper = "/root/reference/D3D12RaytracingFallback/src/RayTracingHelper.hlsli"
    target = tlas_lib_path()
    if not os.path.exists(helper):
        return target if os.path.exists(target) else None
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE, "ref", f) for f in ("prepass.py", "hlsl_compat.h", "ref_tlas.cpp")] + [helper]
    if not force and os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs):
        return target
    sys.path.insert(0, os.path.join(HERE, "ref"))
    import prepass
    prepass.run_tlas(helper, os.path.join(OUT, "tlas_gen.inc"))
    cmd = [GXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-mfma", "-ffp-contract=off", "-fsingle-precision-constant", "-fno-fast-math",
           "-fvisibility=hidden", "-w", "-I" + os.path.join(ROOT, "include"), "-I" + HERE, os.path.join(HERE, "ref", "ref_tlas.cpp"), "-o", target]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref tlas build failed:\n" + r.stdout)
    return target


ENTRY = "/root/reference/TracerBoy/SoftwareRayTraceCS.hlsl"


def frame_lib_path():
    return os.path.join(OUT, "libref_frame.so")


def build_frame(force=False):
    """oracle/_ref/libref_frame.so: the per-pixel wrapper (GetBlueNoise, AOV writers, RayTraceCommon, the entry point's
    per-pixel part) as host C++, PathTrace = the synthetic stand-in."""
    target = frame_lib_path()
    raygen = "/root/reference/TracerBoy/RayGenCommon.h"
    if not (os.path.exists(raygen) and os.path.exists(ENTRY)):
        return target if os.path.exists(target) else None
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE, "ref", f) for f in ("prepass.py", "hlsl_compat.h", "ref_frame.cpp", "synthetic_tracer.h")] + \
           [raygen, ENTRY, os.path.join(HERE, "glue.h"), os.path.join(HERE, "liboracle.so")]
    if not force and os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs):
        return target
    sys.path.insert(0, os.path.join(HERE, "ref"))
    import prepass
    prepass.run_frame(raygen, ENTRY, os.path.join(OUT, "frame_gen.inc"))
    cmd = [GXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-mfma", "-ffp-contract=off", "-fsingle-precision-constant", "-fno-fast-math",
           "-fvisibility=hidden", "-w", "-I" + os.path.join(ROOT, "include"), "-I" + HERE, os.path.join(HERE, "ref", "ref_frame.cpp"),
           "-o", target, "-L" + HERE, "-loracle", "-Wl,-rpath,$ORIGIN/.."]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref frame wrapper build failed:\n" + r.stdout)
    return target


GENHIST = "/root/reference/TracerBoy/GenerateHistogramCS.hlsl"
AVGLUM = "/root/reference/TracerBoy/CalculateAveragedLuminanceCS.hlsl"


def hist_lib_path():
    return os.path.join(OUT, "libref_hist.so")


def build_hist(force=False):
    """oracle/_ref/libref_hist.so: the reference's auto-exposure shaders (GenerateHistogramCS.hlsl, CalculateAveragedLuminanceCS.hlsl)
    as host C++, a 16x16 group = 256 host threads."""
    target = hist_lib_path()
    if not (os.path.exists(TONEMAP) and os.path.exists(GENHIST) and os.path.exists(AVGLUM)):
        return target if os.path.exists(target) else None
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE, "ref", f) for f in ("prepass.py", "hlsl_compat.h", "ref_hist.cpp")] + [TONEMAP, GENHIST, AVGLUM]
    if not force and os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs):
        return target
    sys.path.insert(0, os.path.join(HERE, "ref"))
    import prepass
    prepass.run_hist(TONEMAP, GENHIST, AVGLUM, os.path.join(OUT, "hist_gen.inc"), os.path.join(OUT, "hist_avg.inc"))
    cmd = [GXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-mfma", "-ffp-contract=off", "-fsingle-precision-constant",
           "-fno-fast-math", "-fvisibility=hidden", "-w", "-I" + os.path.join(ROOT, "include"), "-I" + HERE,
           os.path.join(HERE, "ref", "ref_hist.cpp"), "-o", target]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref histogram build failed:\n" + r.stdout)
    return target


TEMPORAL = "/root/reference/TracerBoy/TemporalAccumulationCS.hlsl"


def temporal_lib_path():
    return os.path.join(OUT, "libref_temporal.so")


def build_temporal(force=False):
    """oracle/_ref/libref_temporal.so: the reference's TemporalAccumulationCS.hlsl main() as host C++ (resources shimmed)."""
    target = temporal_lib_path()
    if not (os.path.exists(TONEMAP) and os.path.exists(TEMPORAL)):
        return target if os.path.exists(target) else None
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE, "ref", f) for f in ("prepass.py", "hlsl_compat.h", "ref_temporal.cpp")] + [TONEMAP, TEMPORAL]
    if not force and os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs):
        return target
    sys.path.insert(0, os.path.join(HERE, "ref"))
    import prepass
    prepass.run_temporal(TONEMAP, TEMPORAL, os.path.join(OUT, "temporal_gen.inc"))
    cmd = [GXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-mfma", "-ffp-contract=off", "-fsingle-precision-constant",
           "-fno-fast-math", "-fvisibility=hidden", "-w", "-I" + os.path.join(ROOT, "include"), "-I" + HERE,
           os.path.join(HERE, "ref", "ref_temporal.cpp"), "-o", target]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref temporal accumulation build failed:\n" + r.stdout)
    return target


TRAVERSE = "/root/reference/D3D12RaytracingFallback/src/TraverseFunction.hlsli"
HELPER = "/root/reference/D3D12RaytracingFallback/src/RayTracingHelper.hlsli"  # also used by build_boxes


def traverse_lib_path():
    return os.path.join(OUT, "libref_traverse.so")


def build_traverse(force=False):
    """oracle/_ref/libref_traverse.so: the reference's RayBoxTest (contraction on), GetRayData and RayTriangleIntersect
    (contraction off) as host C++."""
    target = traverse_lib_path()
    if not os.path.exists(TRAVERSE):
        return target if os.path.exists(target) else None
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE, "ref", f) for f in ("prepass.py", "hlsl_compat.h", "ref_traverse_box.cpp", "ref_traverse_rest.cpp",
                                                     "ref_traverse_loop.cpp", "ref_traverse_loop2.cpp")] + [TRAVERSE, HELPER]
    if not force and os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs):
        return target
    sys.path.insert(0, os.path.join(HERE, "ref"))
    import prepass
    prepass.run_traverse(TRAVERSE, os.path.join(OUT, "traverse_box_gen.inc"), os.path.join(OUT, "traverse_rest_gen.inc"))
    prepass.run_traverse_loop(HELPER, TRAVERSE, os.path.join(OUT, "traverse_loop_gen.inc"))
    prepass.run_intersect("/root/reference/TracerBoy/SharedShaderStructs.h", "/root/reference/TracerBoy/SharedHitGroup.h",
                          "/root/reference/TracerBoy/RayGenCommon.h", os.path.join(OUT, "intersect_gen.inc"))
    common = [GXX, "-O2", "-std=c++17", "-fPIC", "-mfma", "-fsingle-precision-constant", "-fno-fast-math", "-fvisibility=hidden", "-w",
              "-I" + os.path.join(ROOT, "include"), "-I" + HERE, "-c"]
    prepass.run_instance_desc("/root/reference/D3D12RaytracingFallback/src/RayTracingHlslCompat.h", os.path.join(OUT, "instance_desc_gen.inc"))
    objs = []
    for name, contract in (("ref_traverse_box", "fast"), ("ref_traverse_rest", "off"), ("ref_traverse_loop", "off"), ("ref_traverse_loop2", "off")):
        obj = os.path.join(OUT, name + ".o")
        r = subprocess.run(common + ["-fopenmp", "-ffp-contract=" + contract, os.path.join(HERE, "ref", name + ".cpp"), "-o", obj],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle/_ref traversal build failed:\n" + r.stdout)
        objs.append(obj)
    # two libraries from the same objects: the single-level loop (FAST_PATH 1, TracerBoy's configuration) and the two-level
    # loop (FAST_PATH 0) define the same functions, so each is linked with the three pure functions on its own
    for out, mine in ((target, objs[:3]), (os.path.join(OUT, "libref_traverse2.so"), objs[:2] + objs[3:])):
        r = subprocess.run([GXX, "-shared", "-fopenmp", "-o", out] + mine, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle/_ref traversal link failed:\n" + r.stdout)
    return target


MORTON = "/root/reference/D3D12RaytracingFallback/src/CalculateMortonCodesBindings.h"


def morton_lib_path():
    return os.path.join(OUT, "libref_morton.so")


def build_morton(force=False):
    """oracle/_ref/libref_morton.so: the reference's CalculateMortonCode as host C++."""
    target = morton_lib_path()
    if not os.path.exists(MORTON):
        return target if os.path.exists(target) else None
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE, "ref", f) for f in ("prepass.py", "hlsl_compat.h", "ref_morton.cpp")] + [MORTON]
    if not force and os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs):
        return target
    sys.path.insert(0, os.path.join(HERE, "ref"))
    import prepass
    prepass.run_morton(MORTON, os.path.join(OUT, "morton_gen.inc"))
    cmd = [GXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-mfma", "-ffp-contract=off", "-fsingle-precision-constant", "-fno-fast-math",
           "-fvisibility=hidden", "-w", "-I" + os.path.join(ROOT, "include"), "-I" + HERE, os.path.join(HERE, "ref", "ref_morton.cpp"), "-o", target]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref morton build failed:\n" + r.stdout)
    return target


KARRAS = "/root/reference/D3D12RaytracingFallback/src/BuildBVHSplits.hlsli"


def karras_lib_path():
    return os.path.join(OUT, "libref_karras.so")


def build_karras(force=False):
    """oracle/_ref/libref_karras.so: the reference's GenerateHierarchy (Karras 2012) as host C++."""
    target = karras_lib_path()
    if not os.path.exists(KARRAS):
        return target if os.path.exists(target) else None
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE, "ref", f) for f in ("prepass.py", "hlsl_compat.h", "ref_karras.cpp")] + [KARRAS]
    if not force and os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs):
        return target
    sys.path.insert(0, os.path.join(HERE, "ref"))
    import prepass
    prepass.run_karras(KARRAS, os.path.join(OUT, "karras_gen.inc"))
    cmd = [GXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-fno-fast-math", "-fvisibility=hidden", "-w",
           "-I" + os.path.join(ROOT, "include"), "-I" + HERE, os.path.join(HERE, "ref", "ref_karras.cpp"), "-o", target]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref karras build failed:\n" + r.stdout)
    return target


TREELET_H = "/root/reference/D3D12RaytracingFallback/src/TreeletReorderBindings.h"
TREELET = "/root/reference/D3D12RaytracingFallback/src/TreeletReorder.hlsl"


def treelet_lib_path():
    return os.path.join(OUT, "libref_treelet.so")


def build_treelet(force=False):
    """oracle/_ref/libref_treelet.so: the reference's treelet optimisation (group shader) as host C++, 32 threads + barrier."""
    target = treelet_lib_path()
    if not (os.path.exists(TREELET_H) and os.path.exists(TREELET)):
        return target if os.path.exists(target) else None
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE, "ref", f) for f in ("prepass.py", "hlsl_compat.h", "ref_treelet.cpp")] + [TREELET_H, TREELET]
    if not force and os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs):
        return target
    sys.path.insert(0, os.path.join(HERE, "ref"))
    import prepass
    prepass.run_treelet(TREELET_H, TREELET, os.path.join(OUT, "treelet_gen.inc"))
    cmd = [GXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-mfma", "-ffp-contract=off", "-fsingle-precision-constant",
           "-fno-fast-math", "-fvisibility=hidden", "-w", "-I" + os.path.join(ROOT, "include"), "-I" + HERE,
           os.path.join(HERE, "ref", "ref_treelet.cpp"), "-o", target]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref treelet build failed:\n" + r.stdout)
    return target


HELPER = "/root/reference/D3D12RaytracingFallback/src/RayTracingHelper.hlsli"


def boxes_lib_path():
    return os.path.join(OUT, "libref_boxes.so")


def build_boxes(force=False):
    """oracle/_ref/libref_boxes.so: the reference's leaf / parent box constructors as host C++."""
    target = boxes_lib_path()
    if not os.path.exists(HELPER):
        return target if os.path.exists(target) else None
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE, "ref", f) for f in ("prepass.py", "hlsl_compat.h", "ref_boxes.cpp")] + [HELPER]
    if not force and os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs):
        return target
    sys.path.insert(0, os.path.join(HERE, "ref"))
    import prepass
    prepass.run_boxes(HELPER, os.path.join(OUT, "boxes_gen.inc"))
    cmd = [GXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-mfma", "-ffp-contract=off", "-fsingle-precision-constant", "-fno-fast-math",
           "-fvisibility=hidden", "-w", "-I" + os.path.join(ROOT, "include"), "-I" + HERE, os.path.join(HERE, "ref", "ref_boxes.cpp"), "-o", target]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref boxes build failed:\n" + r.stdout)
    return target


STRUCTS = "/root/reference/TracerBoy/SharedShaderStructs.h"
RAYGEN = "/root/reference/TracerBoy/RayGenCommon.h"


def raygen_lib_path():
    return os.path.join(OUT, "libref_raygen.so")


def build_raygen(force=False):
    """oracle/_ref/libref_raygen.so: light sampling, environment lookup, Halton and hash13 of RayGenCommon.h as host C++."""
    target = raygen_lib_path()
    if not (os.path.exists(STRUCTS) and os.path.exists(RAYGEN)):
        return target if os.path.exists(target) else None
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE, "ref", f) for f in ("prepass.py", "hlsl_compat.h", "ref_raygen.cpp")] + \
           [STRUCTS, RAYGEN, os.path.join(HERE, "glue.h"), os.path.join(HERE, "liboracle.so")]
    if not force and os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs):
        return target
    sys.path.insert(0, os.path.join(HERE, "ref"))
    import prepass
    prepass.run_raygen(STRUCTS, RAYGEN, os.path.join(OUT, "raygen_light_gen.inc"), os.path.join(OUT, "raygen_gen.inc"))
    prepass.run_material(STRUCTS, TONEMAP, "/root/reference/TracerBoy/SharedRaytracing.h", RAYGEN, os.path.join(OUT, "raygen_material_gen.inc"))
    cmd = [GXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-mfma", "-ffp-contract=off", "-fsingle-precision-constant", "-fno-fast-math",
           "-fvisibility=hidden", "-w", "-I" + os.path.join(ROOT, "include"), "-I" + HERE, os.path.join(HERE, "ref", "ref_raygen.cpp"),
           "-o", target, "-L" + HERE, "-loracle", "-Wl,-rpath,$ORIGIN/.."]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref raygen build failed:\n" + r.stdout)
    return target


if __name__ == "__main__":
    print(build_temporal(force="--force" in sys.argv))
    print(build_frame(force="--force" in sys.argv))
    print(build_refit(force="--force" in sys.argv))
    print(build_load(force="--force" in sys.argv))
    print(build_treelet_pass(force="--force" in sys.argv))
    print(build_hist(force="--force" in sys.argv))
    print(build_raygen(force="--force" in sys.argv))
    print(build_boxes(force="--force" in sys.argv))
    print(build_treelet(force="--force" in sys.argv))
    print(build_karras(force="--force" in sys.argv))
    print(build_morton(force="--force" in sys.argv))
    print(build_traverse(force="--force" in sys.argv))
    print(build_post(force="--force" in sys.argv))
    print(build(force="--force" in sys.argv))
