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

CUDA_SRCS = ["csrc/cuda/bvh_build.cu", "csrc/cuda/pathtrace.cu", "csrc/cuda/postprocess.cu", "csrc/cuda/reduce.cu"]
HOST_SRCS = ["csrc/host/api.cpp", "csrc/host/comm.cpp", "csrc/host/scene.cpp", "csrc/host/image_io.cpp", "csrc/host/image_decode.cpp", "csrc/host/tlas.cpp"]
HEADERS = ["csrc/cuda/device_types.h", "csrc/cuda/pathtrace.h", "csrc/cuda/traverse.cuh", "csrc/cuda/launch.h", "csrc/cuda/postprocess.h", "csrc/cuda/reduce.h", "csrc/host/handle.h",
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
            os.path.join(PKG, "csrc/host/image_decode.cpp"), os.path.join(PKG, "csrc/host/scene.h")]
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
          "-I" + os.path.join(PKG, "csrc/host"), mine[0], mine[1], mine[2]] + objs + [rply, "-o", target])
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
    "dragon": "variant:dragon",
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


def _write_ply(path, pos, nrm, uv, faces):
    """Binary little-endian PLY with the vertex layout of the scene's own meshes (x y z nx ny nz u v; uint8-counted int faces)."""
    import numpy as np
    v = np.concatenate([pos, nrm, uv], axis=1).astype("<f4")
    f = np.zeros(faces.shape[0], dtype=[("n", "u1"), ("i", "<i4", 3)])
    f["n"] = 3
    f["i"] = faces
    with open(path, "wb") as out:
        out.write(("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                   "property float nx\nproperty float ny\nproperty float nz\nproperty float u\nproperty float v\n"
                   "element face %d\nproperty list uint8 int vertex_indices\nend_header\n" % (v.shape[0], f.shape[0])).encode())
        out.write(v.tobytes())
        out.write(f.tobytes())


def _grid_surface(point_fn, nu, nv, closed_u=True):
    """Triangulated (nu x nv) parametric surface with smooth normals from the grid's own tangents."""
    import numpy as np
    u = np.arange(nu + (0 if closed_u else 1), dtype=np.float64) / nu
    v = np.arange(nv + 1, dtype=np.float64) / nv
    U, V = np.meshgrid(u, v, indexing="ij")
    P = point_fn(U, V)
    cols = P.shape[0]
    du = (np.roll(P, -1, 0) - np.roll(P, 1, 0)) if closed_u else np.gradient(P, axis=0)
    dv = np.gradient(P, axis=1)
    N = np.cross(dv, du)
    ln = np.linalg.norm(N, axis=2, keepdims=True)
    radial = P - P.reshape(-1, 3).mean(0)
    radial /= np.maximum(np.linalg.norm(radial, axis=2, keepdims=True), 1e-20)
    N = np.where(ln > 1e-12, N / np.maximum(ln, 1e-20), radial)
    N = np.where((N * radial).sum(2, keepdims=True) < 0, -N, N)
    idx = np.arange(cols * (nv + 1)).reshape(cols, nv + 1)
    a = idx[:-1 if not closed_u else None, :-1]
    b = (np.roll(idx, -1, 0) if closed_u else idx[1:])[: a.shape[0], :-1]
    c = (np.roll(idx, -1, 0) if closed_u else idx[1:])[: a.shape[0], 1:]
    d = idx[: a.shape[0], 1:]
    faces = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, d], -1).reshape(-1, 3)])
    return P.reshape(-1, 3), N.reshape(-1, 3), np.stack([U, V], -1).reshape(-1, 2), faces


def _dragon_variant(tmp):
    """BASELINE.json configs[2]: the mount holds Scenes/dragon/scene.pbrt, its environment map and 12 of
    its 16 meshes (51 140 triangles); the four missing ones are the two ground pieces (Mesh012 GroundInner, Mesh007
    GroundOuter) and the dragon's body (Mesh008, Mesh013). The variant is the reference's own scene.pbrt, read from
    the mount at build time and left untouched -- same camera, same nine matte materials, lit only by textures/envmap.hdr
    -- with the four missing .ply files generated here: a deterministic high-poly body of multi-scale bumps where the
    dragon stands (819 200 + 32 768 triangles, so the whole scene is the ~0.9 M-triangle, all-diffuse, environment-lit
    workload the config names) and two small ground pieces. Nothing is written into the repo except the flattened
    scenes/_cache/dragon.tbscene (git-ignored)."""
    import numpy as np
    src = os.path.join(REF, "Scenes/dragon")
    if not os.path.exists(os.path.join(src, "scene.pbrt")):
        return None
    os.makedirs(os.path.join(tmp, "models"))
    os.symlink(os.path.join(src, "textures"), os.path.join(tmp, "textures"))
    for f in os.listdir(os.path.join(src, "models")):
        os.symlink(os.path.join(src, "models", f), os.path.join(tmp, "models", f))
    shutil.copy(os.path.join(src, "scene.pbrt"), os.path.join(tmp, "scene.pbrt"))
    tau = 2.0 * np.pi

    def body(U, V):  # closed lat-long surface around (0, 5.6, -1.2): lobes, ridges and fine "scales"
        th, ph = tau * U, np.pi * V
        r = 1.0 + 0.18 * np.sin(3 * th + 0.7) * np.sin(2 * ph) + 0.10 * np.sin(7 * th) * np.sin(5 * ph + 0.3) \
            + 0.035 * np.sin(29 * th + 1.1) * np.sin(23 * ph) + 0.008 * np.sin(131 * th) * np.sin(113 * ph + 0.5)
        x = 4.2 * r * np.sin(ph) * np.cos(th)
        y = 5.6 + 5.0 * r * np.cos(ph)
        z = -1.2 + 5.2 * r * np.sin(ph) * np.sin(th)
        return np.stack([x, y, z], -1)

    def tail(U, V):  # a tapering tube coiled around the pedestal
        t = V
        ang = tau * (1.35 * t + 0.1)
        rad = 7.6 - 1.8 * t
        cx, cy, cz = rad * np.cos(ang), 0.9 + 2.4 * t * t, -1.2 + rad * np.sin(ang)
        w = 0.75 * (1.0 - 0.85 * t) * (1.0 + 0.06 * np.sin(40 * tau * t) )
        a = tau * U
        # frame: radial (outwards from the pedestal axis) and up
        ox, oz = np.cos(ang), np.sin(ang)
        x = cx + w * np.cos(a) * ox
        y = cy + w * np.sin(a)
        z = cz + w * np.cos(a) * oz
        return np.stack([x, y, z], -1)

    def disc(r0, r1, height):
        def f(U, V):
            r = r0 + (r1 - r0) * V
            return np.stack([r * np.cos(tau * U), np.full_like(U, height) + 0.02 * np.sin(9 * tau * U) * (r / r1), r * np.sin(tau * U)], -1)
        return f

    _write_ply(os.path.join(tmp, "models", "Mesh008.ply"), *_grid_surface(body, 640, 640))
    _write_ply(os.path.join(tmp, "models", "Mesh013.ply"), *_grid_surface(tail, 32, 512))
    _write_ply(os.path.join(tmp, "models", "Mesh012.ply"), *_grid_surface(disc(0.0, 9.0, 0.24), 64, 16))
    _write_ply(os.path.join(tmp, "models", "Mesh007.ply"), *_grid_surface(disc(9.0, 13.0, 0.1), 64, 8))
    return os.path.join(tmp, "scene.pbrt")


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
            src = _dragon_variant(tmp) if rel == "variant:dragon" else _vw_van_variant(tmp)
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
    build_ref.build_tlas(force)
    build_ref.build_flatten(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", os.listdir(LIB))
