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


TRAVERSE = "/root/reference/D3D12RaytracingFallback/src/TraverseFunction.hlsli"


def traverse_lib_path():
    return os.path.join(OUT, "libref_traverse.so")


def build_traverse(force=False):
    """oracle/_ref/libref_traverse.so: the reference's RayBoxTest (contraction on), GetRayData and RayTriangleIntersect
    (contraction off) as host C++."""
    target = traverse_lib_path()
    if not os.path.exists(TRAVERSE):
        return target if os.path.exists(target) else None
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE, "ref", f) for f in ("prepass.py", "hlsl_compat.h", "ref_traverse_box.cpp", "ref_traverse_rest.cpp")] + [TRAVERSE]
    if not force and os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs):
        return target
    sys.path.insert(0, os.path.join(HERE, "ref"))
    import prepass
    prepass.run_traverse(TRAVERSE, os.path.join(OUT, "traverse_box_gen.inc"), os.path.join(OUT, "traverse_rest_gen.inc"))
    common = [GXX, "-O2", "-std=c++17", "-fPIC", "-mfma", "-fsingle-precision-constant", "-fno-fast-math", "-fvisibility=hidden", "-w",
              "-I" + os.path.join(ROOT, "include"), "-I" + HERE, "-c"]
    objs = []
    for name, contract in (("ref_traverse_box", "fast"), ("ref_traverse_rest", "off")):
        obj = os.path.join(OUT, name + ".o")
        r = subprocess.run(common + ["-ffp-contract=" + contract, os.path.join(HERE, "ref", name + ".cpp"), "-o", obj],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle/_ref traversal build failed:\n" + r.stdout)
        objs.append(obj)
    r = subprocess.run([GXX, "-shared", "-o", target] + objs, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref traversal link failed:\n" + r.stdout)
    return target


if __name__ == "__main__":
    print(build_traverse(force="--force" in sys.argv))
    print(build_post(force="--force" in sys.argv))
    print(build(force="--force" in sys.argv))
