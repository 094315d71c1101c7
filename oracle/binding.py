"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE). Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference leg may import this module."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
BLUE_NOISE = os.path.join(_ROOT, "tracerboy_b200", "data", "bluenoise_rgba8_256.bin")
_lib = None


def lib_path():
    return os.path.join(_HERE, "liboracle.so")


_ref = None


def ref_lib_path():
    return os.path.join(_HERE, "_ref", "libref_core.so")


def reference_core_available():
    return os.path.exists(ref_lib_path())


def use_reference_core(on):
    """Route the oracle's per-pixel PathTrace through oracle/_ref (the reference's kernel.glsl compiled
    as host C++) instead of the hand-restated core. Returns the number of pixels it has traced so far."""
    global _ref
    load()
    if _ref is None:
        _ref = C.CDLL(ref_lib_path())
        _ref.ref_core_calls.restype = C.c_ulonglong
    _ref.ref_core_enable(1 if on else 0)
    return _ref.ref_core_calls()


def ref_traverse_lib_path():
    return os.path.join(_HERE, "_ref", "libref_traverse.so")


def reference_traverse_available():
    return os.path.exists(ref_traverse_lib_path())


def ray_query_functions(which):
    """The three pure functions of the ray query — ray_data(org, dir), ray_box(...), ray_tri(...) — of the oracle
    (`which` = "oracle") or of the reference's own text compiled from the mount ("reference"). All take / return numpy
    float32 arrays; see tests/test_cpu_oracle.py::test_ray_query_functions_equal_reference_text."""
    lib = load() if which == "oracle" else C.CDLL(ref_traverse_lib_path())
    pre = "oracle_" if which == "oracle" else "ref_"
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    data, box, tri = getattr(lib, pre + "ray_data"), getattr(lib, pre + "ray_box"), getattr(lib, pre + "ray_tri")
    data.argtypes = [fp, fp, fp, fp, fp, ip]; data.restype = None
    box.argtypes = [C.c_float, fp, fp, fp, fp, fp]; box.restype = C.c_int
    tri.argtypes = [fp, fp, ip, fp, fp, fp]; tri.restype = C.c_int

    def f(a):
        return a.ctypes.data_as(fp)

    def ray_data(org, dir_):
        inv, oinv, shear, swz = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(3, np.int32)
        data(f(org), f(dir_), f(inv), f(oinv), f(shear), swz.ctypes.data_as(ip))
        return inv, oinv, shear, swz

    def ray_box(closest, oinv, inv, c, h):
        t = np.zeros(1, np.float32)
        hit = box(C.c_float(closest), f(oinv), f(inv), f(c), f(h), f(t))
        return hit, t[0]

    def ray_tri(hit_t, org, swz, shear, v9):
        t = np.array([hit_t], np.float32)
        bary = np.zeros(2, np.float32)
        hit = tri(f(t), f(org), swz.ctypes.data_as(ip), f(shear), f(v9), f(bary))
        return hit, t[0], bary
    return ray_data, ray_box, ray_tri


def ref_post_lib_path():
    return os.path.join(_HERE, "_ref", "libref_post.so")


def reference_post_available():
    return os.path.exists(ref_post_lib_path())


_ref_post = None


def postprocess_image(img, output_type, settings, aux=None):
    """oracle_postprocess_image (oracle/postprocess.cpp): returns (float4 image, rgba8 image, histogram[256],
    averaged luminance). `settings` is a tracerboy_b200.api.PostProcessSettings."""
    lib = load()
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape[:2]
    auxp = None
    if aux is not None:
        aux = np.ascontiguousarray(aux, np.float32)
        auxp = aux.ctypes.data_as(C.c_void_p)
    out = np.empty((h, w, 4), np.float32)
    rgba8 = np.empty((h, w, 4), np.uint8)
    hist = np.zeros(256, np.uint32)
    avg = C.c_float(0.0)
    rc = lib.oracle_postprocess_image(img.ctypes.data_as(C.c_void_p), auxp, w, h, int(output_type), C.byref(settings),
                                      out.ctypes.data_as(C.c_void_p), rgba8.ctypes.data_as(C.c_void_p),
                                      hist.ctypes.data_as(C.c_void_p), C.byref(avg))
    assert rc == 0
    return out, rgba8, hist, avg.value


def reference_postprocess_image(img, output_type, settings, averaged_luminance, aux=None):
    """The reference's own Tonemap.h + PostProcessCS.hlsl Process* functions compiled from the mount as host C++
    (oracle/_ref/libref_post.so). The averaged luminance is an input here."""
    global _ref_post
    if _ref_post is None:
        _ref_post = C.CDLL(ref_post_lib_path())
        _ref_post.ref_postprocess_image.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                                    C.c_float, C.c_void_p]
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape[:2]
    auxp = None
    if aux is not None:
        aux = np.ascontiguousarray(aux, np.float32)
        auxp = aux.ctypes.data_as(C.c_void_p)
    out = np.empty((h, w, 4), np.float32)
    rc = _ref_post.ref_postprocess_image(img.ctypes.data_as(C.c_void_p), auxp, w, h, int(output_type), C.byref(settings),
                                         C.c_float(averaged_luminance), out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out


def temporal_accumulate_image(params, history, current, world_pos, prev_world_pos, normals, moment_history=None):
    """oracle_temporal_accumulate_image (oracle/temporal.cpp): returns (color float4, moment float4 or None)."""
    lib = load()
    imgs = [np.ascontiguousarray(a, np.float32) for a in (history, current, world_pos, prev_world_pos, normals)]
    h, w = imgs[0].shape[:2]
    mh = np.ascontiguousarray(moment_history, np.float32) if moment_history is not None else None
    out = np.empty((h, w, 4), np.float32)
    mom = np.empty((h, w, 4), np.float32) if params.OutputMomentInformation else None
    rc = lib.oracle_temporal_accumulate_image(C.byref(params), w, h, *[a.ctypes.data_as(C.c_void_p) for a in imgs],
                                              mh.ctypes.data_as(C.c_void_p) if mh is not None else None,
                                              out.ctypes.data_as(C.c_void_p),
                                              mom.ctypes.data_as(C.c_void_p) if mom is not None else None)
    assert rc == 0
    return out, mom


def set_literal_mode(mask):
    """Test hook: bit 0 = deviation D6 off (literal rcp(0) = inf), bit 1 = deviation D7 off (NaN rays walk the tree)."""
    load().oracle_set_literal_mode(int(mask))


def set_literal_rcp(on):
    """Test hook: literal rcp(0) = inf in GetRayData (deviation D6 off)."""
    load().oracle_set_literal_rcp(1 if on else 0)


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(lib_path(), mode=C.RTLD_GLOBAL)
        lib.oracle_create.restype = C.c_void_p
        lib.oracle_last_error.restype = C.c_char_p
        lib.oracle_bvh_size.restype = C.c_uint64
        lib.oracle_max_treelet_climb.restype = C.c_uint32
        lib.oracle_num_triangles.restype = C.c_uint32
        lib.oracle_samples.restype = C.c_uint32
        for n in ("oracle_destroy", "oracle_last_error", "oracle_bvh_size", "oracle_max_treelet_climb",
                  "oracle_num_triangles", "oracle_samples", "oracle_invalidate"):
            getattr(lib, n).argtypes = [C.c_void_p]
        lib.oracle_load_scene.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        lib.oracle_build_bvh.argtypes = [C.c_void_p, C.c_int]
        lib.oracle_get_bvh.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        lib.oracle_trace_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        lib.oracle_get_camera.argtypes = [C.c_void_p, C.c_void_p]
        lib.oracle_set_camera.argtypes = [C.c_void_p, C.c_void_p]
        lib.oracle_resize.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        lib.oracle_select_pixel.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.oracle_set_samples.argtypes = [C.c_void_p, C.c_uint32]
        lib.oracle_set_shard.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        lib.oracle_set_row_shard.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        lib.oracle_math_eval.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        lib.oracle_math_eval.restype = None
        lib.oracle_morton.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_morton.restype = C.c_uint32
        lib.oracle_render.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_float, C.c_int, C.POINTER(C.c_double)]
        lib.oracle_get_counts.argtypes = [C.c_void_p, C.c_void_p]
        lib.oracle_get_stats.argtypes = [C.c_void_p, C.c_void_p]
        lib.oracle_readback.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64]
        lib.oracle_temporal_accumulate_image.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32] + [C.c_void_p] * 8
        lib.oracle_postprocess_image.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


MATH_FN = {"sin": 0, "cos": 1, "acos": 2, "atan2": 3, "exp": 4, "log": 5, "pow": 6, "hash13": 7, "halton": 8}


def math_eval(fn, x, y=None):
    """Evaluate one of the pinned intrinsics (tb_math.h) / noise functions over float32 arrays."""
    lib = load()
    x = np.ascontiguousarray(x, np.float32)
    n = x.size // 3 if fn == "hash13" else x.size
    y = np.zeros(n, np.float32) if y is None else np.ascontiguousarray(np.broadcast_to(np.asarray(y, np.float32), (n,)), np.float32)
    out = np.empty(n, np.float32)
    lib.oracle_math_eval(MATH_FN[fn], x.ctypes.data, y.ctypes.data, out.ctypes.data, n)
    return out


def morton(centroid, smin, smax):
    a = [np.ascontiguousarray(v, np.float32) for v in (centroid, smin, smax)]
    return load().oracle_morton(a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data)


class Oracle:
    """Same verbs as tracerboy_b200.TracerBoy so the parity tests read symmetrically."""
    _shape = {0: (np.float32, 4), 1: (np.float32, 4), 2: (np.float32, 3), 3: (np.float32, 4), 4: (np.float32, 4),
              5: (np.float32, 1), 6: (np.float32, 4), 7: (np.float32, 4), 8: (np.uint32, 2), 9: (np.uint32, 2)}

    def __init__(self):
        self.lib = load()
        self.h = C.c_void_p(self.lib.oracle_create())
        self.width = self.height = 0

    def close(self):
        if self.h:
            self.lib.oracle_destroy(self.h)
            self.h = None

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError("oracle error %d: %s" % (rc, self.lib.oracle_last_error(self.h).decode()))

    def LoadScene(self, tbscene, treelet_passes=3, blue_noise=BLUE_NOISE):
        self._ck(self.lib.oracle_load_scene(self.h, tbscene.encode(), blue_noise.encode()))
        self._ck(self.lib.oracle_build_bvh(self.h, treelet_passes))

    def GetBVH(self):
        n = self.lib.oracle_bvh_size(self.h)
        buf = np.empty(n, np.uint8)
        self._ck(self.lib.oracle_get_bvh(self.h, buf.ctypes.data, n))
        return buf

    def ScenePositions(self):
        """The scene's pooled vertex positions [V, 3] (a copy)."""
        self.lib.oracle_scene_positions.restype = C.c_void_p
        self.lib.oracle_scene_num_positions.restype = C.c_uint64
        n = self.lib.oracle_scene_num_positions(self.h)
        p = self.lib.oracle_scene_positions(self.h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n, 3)).copy()

    def UpdateBVH(self, positions):
        """PERFORM_UPDATE: new vertex positions (same count), hierarchy kept, boxes refitted."""
        positions = np.ascontiguousarray(positions, np.float32)
        self._ck(self.lib.oracle_update_bvh(self.h, positions.ctypes.data_as(C.c_void_p), C.c_uint64(positions.shape[0])))

    def MaxTreeletClimb(self):
        return self.lib.oracle_max_treelet_climb(self.h)

    def NumTriangles(self):
        return self.lib.oracle_num_triangles(self.h)

    def TraceRays(self, rays):
        from tracerboy_b200.api import HIT_DTYPE, RAY_DTYPE
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        hits = np.empty(rays.shape[0], HIT_DTYPE)
        self._ck(self.lib.oracle_trace_rays(self.h, rays.ctypes.data, rays.shape[0], hits.ctypes.data))
        return hits

    def GetCamera(self):
        from tracerboy_b200.api import Camera
        c = Camera()
        self.lib.oracle_get_camera(self.h, C.byref(c))
        return c

    def SetCamera(self, cam):
        self.lib.oracle_set_camera(self.h, C.byref(cam))

    def Resize(self, w, h):
        self._ck(self.lib.oracle_resize(self.h, w, h))
        self.width, self.height = w, h

    def SelectPixel(self, x, y):
        self.lib.oracle_select_pixel(self.h, x, y)

    def SetFrameShard(self, offset, stride):
        self._ck(self.lib.oracle_set_shard(self.h, offset, stride))

    def SetRowShard(self, offset, stride):
        self._ck(self.lib.oracle_set_row_shard(self.h, offset, stride))

    def SetSamples(self, n):
        self.lib.oracle_set_samples(self.h, n)

    def Render(self, settings, samples=1, time=0.0, threads=0):
        sec = C.c_double()
        self._ck(self.lib.oracle_render(self.h, C.byref(settings), samples, C.c_float(time), threads, C.byref(sec)))
        return sec.value

    def Counts(self):
        c = (C.c_uint64 * 3)()
        self.lib.oracle_get_counts(self.h, c)
        return {"rays": c[0], "boxes": c[1], "tris": c[2]}

    def GetReadbackStats(self):
        from tracerboy_b200.api import ReadbackStats
        s = ReadbackStats()
        self.lib.oracle_get_stats(self.h, C.byref(s))
        return s

    def Readback(self, kind):
        dt, ch = self._shape[kind]
        shape = (self.height, self.width, ch) if ch > 1 else (self.height, self.width)
        out = np.empty(shape, dt)
        self._ck(self.lib.oracle_readback(self.h, kind, out.ctypes.data, out.nbytes))
        return out

    @staticmethod
    def max_threads():
        return load().oracle_max_threads()
