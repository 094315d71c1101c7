"""Build the native libraries in-tree (no torch JIT cache: the .so files travel with the repo).

  tracerboy_b200/lib/libtracerboy_b200.so   CUDA kernels (sm_100a) + C ABI          [product]
  tracerboy_b200/lib/libtb_pbrtimport.so    optional PBRT importer, only when the reference's
                                            vendored pbrt-parser sources are mounted   [product, optional]
  oracle/liboracle.so                       CPU oracle                               [test infrastructure]
  oracle/_ref/libref_core.so                reference core compiled from the mount   [test infrastructure, optional]
  scenes/_cache/*.tbscene                   flattened bundled scenes (need the mount to regenerate)
"""
import ctypes
import hashlib
import os
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "tracerboy_b200")
LIB = os.path.join(PKG, "lib")
BUILD = os.path.join(ROOT, "build")
REF = "/root/reference"
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
              "-I" + os.path.join(ROOT, "include"), "-ccbin", GXX]

CUDA_SRCS = ["csrc/cuda/bvh_build.cu", "csrc/cuda/pathtrace.cu", "csrc/cuda/postprocess.cu"]
HOST_SRCS = ["csrc/host/api.cpp", "csrc/host/scene.cpp", "csrc/host/image_io.cpp"]
HEADERS = ["csrc/cuda/device_types.h", "csrc/cuda/pathtrace.h", "csrc/cuda/traverse.cuh", "csrc/cuda/launch.h", "csrc/cuda/postprocess.h",
           "csrc/common/tb_math.h", "csrc/common/tb_vec.h", "csrc/host/scene.h", "../include/tracerboy_b200.h"]


def _run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        raise RuntimeError("build step failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def _stamp(paths, extra=""):
    h = hashlib.sha1(extra.encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _up_to_date(target, stamp):
    sf = target + ".stamp"
    return os.path.exists(target) and os.path.exists(sf) and open(sf).read() == stamp


def _write_stamp(target, stamp):
    with open(target + ".stamp", "w") as f:
        f.write(stamp)


def build_product(force=False, verbose=False):
    os.makedirs(LIB, exist_ok=True)
    os.makedirs(BUILD, exist_ok=True)
    srcs = [os.path.join(PKG, s) for s in CUDA_SRCS + HOST_SRCS]
    deps = srcs + [os.path.join(PKG, h) for h in HEADERS]
    target = os.path.join(LIB, "libtracerboy_b200.so")
    stamp = _stamp(deps, " ".join(NVCC_FLAGS))
    if not force and _up_to_date(target, stamp):
        return target
    objs = []

    def compile_one(src):
        obj = os.path.join(BUILD, os.path.basename(src) + ".o")
        extra = ["-Xptxas", "-v"] if verbose else []
        out = _run([NVCC] + NVCC_FLAGS + extra + ["-x", "cu", "-c", src, "-o", obj])
        if verbose:
            print(out)
        return obj

    with ThreadPoolExecutor(4) as ex:
        objs = list(ex.map(compile_one, srcs))
    _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", GXX, "-o", target] + objs + ["-ldl"])
    _write_stamp(target, stamp)
    return target


def build_pbrt_import(force=False):
    """Optional: needs the reference mount (third-party pbrt-parser compiled in place)."""
    parser = os.path.join(REF, "PBRTParser")
    target = os.path.join(LIB, "libtb_pbrtimport.so")
    if not os.path.isdir(parser):
        return target if os.path.exists(target) else None
    mine = [os.path.join(PKG, "csrc/host/pbrt_import.cpp"), os.path.join(PKG, "csrc/host/scene.cpp"),
            os.path.join(PKG, "csrc/host/scene.h")]
    stamp = _stamp(mine)
    if not force and _up_to_date(target, stamp):
        return target
    odir = os.path.join(BUILD, "pbrt")
    os.makedirs(odir, exist_ok=True)
    third = []
    for sub in ("impl/semantic", "impl/syntactic"):
        d = os.path.join(parser, sub)
        third += [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(".cpp")]

    def cc(src):
        obj = os.path.join(odir, os.path.basename(os.path.dirname(src)) + "_" + os.path.basename(src) + ".o")
        _run([GXX, "-O2", "-std=c++14", "-fPIC", "-w", "-include", "cstdint", "-I" + os.path.join(parser, "include"),
              "-I" + os.path.join(parser, "impl"), "-c", src, "-o", obj])
        return obj

    with ThreadPoolExecutor(8) as ex:
        objs = list(ex.map(cc, third))
    rply = os.path.join(odir, "rply.o")
    _run([GXX.replace("g++", "gcc"), "-O2", "-fPIC", "-w", "-c", os.path.join(parser, "impl/3rdParty/rply.c"), "-o", rply])
    _run([GXX, "-O2", "-std=c++14", "-fPIC", "-shared", "-fvisibility=hidden", "-w", "-include", "cstdint", "-mfma",
          "-ffp-contract=off", "-I" + os.path.join(parser, "include"), "-I" + os.path.join(ROOT, "include"),
          "-I" + os.path.join(PKG, "csrc/host"), mine[0], mine[1]] + objs + [rply, "-o", target])
    _write_stamp(target, stamp)
    return target


def build_oracle(force=False):
    odir = os.path.join(ROOT, "oracle")
    if force:
        _run(["make", "-C", odir, "clean"])
    _run(["make", "-C", odir])
    return os.path.join(odir, "liboracle.so")


BUNDLED_SCENES = {
    "cornell-box": "Scenes/cornell-box/scene.pbrt",
    "teapot": "Scenes/Teapot/scene.pbrt",
    "vw-van": "variant:vw-van",
}


def _vw_van_variant(tmp):
    """BASELINE.json configs[3]: the mount lacks geometry/mesh_00125.ply and the scene's only
    light, textures/pisa_latlong.hdr. The variant is the reference's own vw-van.pbrt, read from the mount at
    build time, minus the one Shape line that names the missing mesh, with Teapot's envmap.hdr standing in
    for the missing lat-long map. Both back-ends render the same variant. Nothing is written into the repo
    except the flattened scenes/_cache/vw-van.tbscene (git-ignored)."""
    src = os.path.join(REF, "Scenes/vw-van")
    env = os.path.join(REF, "Scenes/Teapot/textures/envmap.hdr")
    if not os.path.exists(os.path.join(src, "vw-van.pbrt")) or not os.path.exists(env):
        return None
    os.makedirs(os.path.join(tmp, "textures"))
    os.symlink(os.path.join(src, "geometry"), os.path.join(tmp, "geometry"))
    os.symlink(env, os.path.join(tmp, "textures", "pisa_latlong.hdr"))
    with open(os.path.join(src, "vw-van.pbrt")) as f:
        lines = [ln for ln in f if "geometry/mesh_00125.ply" not in ln]
    out = os.path.join(tmp, "vw-van.pbrt")
    with open(out, "w") as f:
        f.writelines(lines)
    return out


def build_scene_cache(force=False):
    """Flatten the reference's bundled scenes into scenes/_cache/*.tbscene (needs the mount)."""
    cache = os.path.join(ROOT, "scenes", "_cache")
    os.makedirs(cache, exist_ok=True)
    imp = os.path.join(LIB, "libtb_pbrtimport.so")
    if not os.path.isdir(REF) or not os.path.exists(imp):
        return cache
    lib = ctypes.CDLL(imp)
    err = ctypes.create_string_buffer(1024)
    for name, rel in BUNDLED_SCENES.items():
        out = os.path.join(cache, name + ".tbscene")
        if os.path.exists(out) and not force:
            continue
        tmp = None
        if rel.startswith("variant:"):
            tmp = tempfile.mkdtemp(prefix="tb_variant_")
            src = _vw_van_variant(tmp)
            if not src:
                shutil.rmtree(tmp, ignore_errors=True)
                continue
        else:
            src = os.path.join(REF, rel)
        rc = lib.tb_pbrt_convert(src.encode(), out.encode(), err, 1024)
        if tmp:
            shutil.rmtree(tmp, ignore_errors=True)
        if rc != 0:
            raise RuntimeError("scene conversion failed for %s: %s" % (src, err.value.decode()))
    return cache


def build_all(force=False, verbose=False):
    build_product(force, verbose)
    build_pbrt_import(force)
    build_oracle(force)
    build_scene_cache(force)
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import build_ref  # checker only: the reference's kernel.glsl compiled from the mount, when present
    build_ref.build(force)
    build_ref.build_post(force)
    build_ref.build_traverse(force)
    build_ref.build_morton(force)
    build_ref.build_karras(force)
    build_ref.build_treelet(force)
    build_ref.build_boxes(force)
    build_ref.build_raygen(force)
    build_ref.build_temporal(force)
    build_ref.build_hist(force)
    build_ref.build_frame(force)
    build_ref.build_refit(force)
    build_ref.build_load(force)
    build_ref.build_treelet_pass(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", os.listdir(LIB))
