"""Build the product library with extra -D flags into tracerboy_b200/lib/libtb_var_<name>.so (experiments only;
select one at run time with TB_LIB=<path>):

    python tools/build_variant.py pf_far -DTB_EXP_PF_FAR
"""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tracerboy_b200 import build as B

name, flags = sys.argv[1], sys.argv[2:]
out_dir = B.LIB  # next to the product library: the blue-noise table is found relative to the .so
obj_dir = os.path.join(B.BUILD, "variant_" + name)
os.makedirs(out_dir, exist_ok=True)
os.makedirs(obj_dir, exist_ok=True)
srcs = [os.path.join(B.PKG, s) for s in B.CUDA_SRCS + B.HOST_SRCS]


def compile_one(src):
    obj = os.path.join(obj_dir, os.path.basename(src) + ".o")
    B._run([B.NVCC] + B.NVCC_FLAGS + flags + ["-x", "cu", "-c", src, "-o", obj])
    return obj


with ThreadPoolExecutor(5) as ex:
    objs = list(ex.map(compile_one, srcs))
target = os.path.join(out_dir, "libtb_var_%s.so" % name)
B._run([B.NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", B.GXX, "-o", target] + objs + ["-ldl"])
print(target)
