"""CPU tests (no GPU) of the host side: the C-ABI library loads and exports every symbol the
header declares, error behaviour without a device, the .tbscene cache, struct layouts, and the
sample-sharded multi-process reduction (gloo, world size 2)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, scene_path


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "tracerboy_b200.h")).read()
    return sorted(set(re.findall(r"TB_API\s+[\w\s\*]+?\b(tb_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    import tracerboy_b200 as tb
    from tracerboy_b200.api import EXPORTED_SYMBOLS
    lib = C.CDLL(tb.lib_path())
    declared = _header_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), "header declares %s but the library does not export it" % name
    assert sorted(EXPORTED_SYMBOLS) == declared, "python binding list out of sync with the header"


def test_struct_layouts_match_the_reference_contract(built):
    """SharedShaderStructs.h sizes: Material 84, Light 104, TextureData 80, Vertex 32."""
    from tracerboy_b200 import api
    assert C.sizeof(api.Material) == 84
    assert C.sizeof(api.Camera) == 56
    assert C.sizeof(api.Ray) == 32 and C.sizeof(api.Hit) == 32
    assert api.RAY_DTYPE.itemsize == 32 and api.HIT_DTYPE.itemsize == 32
    assert C.sizeof(api.OutputSettings) == 18 * 4


def test_default_settings_match_reference(built):
    """TracerBoy::GetDefaultOutputSettings (TracerBoy.h:290-360) and SURVEY appendix C."""
    import tracerboy_b200 as tb
    s = tb.get_default_output_settings()
    assert (s.OutputType, s.EnableNormalMaps, s.RenderMode) == (0, 0, 0)
    assert (s.EnableNextEventEstimation, s.EnableSamplingImportanceResampling, s.EnableBlueNoise) == (1, 0, 1)
    assert s.MaxBounces == 6 and s.FilterType == 0 and s.FilterWidth == 1.0
    assert s.DOFFocalDistance == 0.0 and s.FireflyClampValue == 0.0 and s.MaxZ == 10000.0
    assert abs(s.ApertureWidth - 0.075) < 1e-7 and s.DebugValue == 1.0 and s.SampleLimit == 0


@pytest.mark.skipif("__import__('conftest').has_cuda()")
def test_no_cpu_fallback(built):
    """Without a CUDA device the product fails loudly (TB_ERR_CUDA); it never routes to the oracle."""
    import tracerboy_b200 as tb
    with pytest.raises(tb.TracerBoyError) as e:
        tb.TracerBoy(0)
    assert e.value.code == -5 and "no CPU fallback" in str(e.value)


def test_product_does_not_reference_the_oracle():
    for d, _, files in os.walk(os.path.join(ROOT, "tracerboy_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(d, f), errors="ignore").read()
                assert "liboracle" not in text and "oracle.binding" not in text and "from oracle" not in text or f == "build.py", f


def test_prebuild_info(built):
    """GetRaytracingAccelerationStructurePrebuildInfo: the reference layout is 116 N - 16 bytes (GpuBVH2Builder.cpp:459);
    the result buffer a caller allocates also holds the traversal layout (64 (N-1) + 48 N) behind it, and the scratch
    size is what the builder really carves its temporaries from."""
    import tracerboy_b200 as tb
    from tracerboy_b200.api import GeometryDesc, PrebuildInfo
    lib = tb.load_library()
    pos = np.zeros((12, 3), np.float32)
    idx = np.arange(9, dtype=np.uint32)
    d = (GeometryDesc * 2)()
    d[0].Positions = pos.ctypes.data; d[0].PositionStrideBytes = 12; d[0].VertexCount = 12
    d[0].Indices = idx.ctypes.data; d[0].IndexFormat = 4; d[0].IndexCount = 9
    d[1].Positions = pos.ctypes.data; d[1].PositionStrideBytes = 12; d[1].VertexCount = 12  # non-indexed: 4 tris
    info = PrebuildInfo()
    assert lib.tb_bvh_prebuild_info(d, 2, C.byref(info)) == 0
    assert info.ReferenceLayoutSizeInBytes == 116 * 7 - 16 and info.ScratchDataSizeInBytes > 7 * (40 + 12 + 16)
    assert info.ResultDataMaxSizeInBytes >= 116 * 7 - 16 + 64 * 6 + 48 * 7 and info.ResultDataMaxSizeInBytes % 16 == 0
    # the layout's byte offsets are 32 bit: 116 N - 16 must stay below 4 GiB, larger inputs are E_INVALIDARG
    limit = lib.tb_max_triangles()
    assert 116 * limit - 16 < 2 ** 32 <= 116 * (limit + 1) - 16
    big = (GeometryDesc * 1)()
    big[0].Positions = pos.ctypes.data; big[0].PositionStrideBytes = 12; big[0].VertexCount = 4000000000  # non-indexed
    assert lib.tb_bvh_prebuild_info(big, 1, C.byref(info)) == -1
    big[0].VertexCount = 3 * limit
    assert lib.tb_bvh_prebuild_info(big, 1, C.byref(info)) == 0 and info.ReferenceLayoutSizeInBytes == 116 * limit - 16
    d[1].IndexFormat = 4  # "If the index buffer is null, the index format must be UNKNOWN" (LoadPrimitivesPass.cpp:73-76)
    assert lib.tb_bvh_prebuild_info(d, 2, C.byref(info)) == -1


def test_tbscene_roundtrip_and_validation(tmp_path, built):
    import tracerboy_b200 as tb
    a, b = str(tmp_path / "a.tbscene"), str(tmp_path / "b.tbscene")
    tb.convert_scene("synthetic:blobs?copies=3&tris=50&seed=4", a)
    tb.convert_scene(a, b)
    assert open(a, "rb").read() == open(b, "rb").read()
    # corrupt an index -> rejected with TB_ERR_IO, not a crash
    raw = bytearray(open(a, "rb").read())
    with pytest.raises(tb.TracerBoyError):
        tb.convert_scene(str(tmp_path / "missing.tbscene"), b)
    open(a, "wb").write(raw[:200])
    with pytest.raises(tb.TracerBoyError):
        tb.convert_scene(a, b)
    with pytest.raises(tb.TracerBoyError):
        tb.convert_scene("synthetic:nope", b)
    with pytest.raises(tb.TracerBoyError) as e:
        tb.convert_scene("scene.fbx", b)  # AssimpImporter slot: not available
    assert e.value.code == -2


def test_tbscene_loader_rejects_every_dangling_index(tmp_path, built):
    """The kernels index materials -> textures -> images (and mix-material ids, the environment image) without bounds
    checks; the reference leans on D3D12 robust buffer access. A stale or malformed cache must be refused on the host:
    each case below flips one field of a valid file and expects TB_ERR_IO, not a device fault later."""
    import struct
    import tracerboy_b200 as tb
    src = str(tmp_path / "ok.tbscene")
    tb.convert_scene("synthetic:showcase?tris=60&seed=3", src)   # has image / checker / scale textures, mix materials, env image
    raw = bytearray(open(src, "rb").read())
    hdr = struct.unpack_from("<8sII7Ii", raw, 0)
    ng, nv, ni, nm, nl, nt, nimg, env = hdr[3:]
    assert nm > 3 and nt > 2 and nimg > 0
    off_geoms = 8 + 4 + 4 + 7 * 4 + 4 + 56 + 48 + 12 + 32
    off_mats = off_geoms + ng * 32 + nv * 12 + nv * 32 + ni * 4
    off_lights = off_mats + nm * 84
    off_tex = off_lights + nl * 104

    def mutated(offset, fmt, value):
        b = bytearray(raw)
        struct.pack_into(fmt, b, offset, value)
        p = str(tmp_path / "bad.tbscene")
        open(p, "wb").write(b)
        return p

    def rejected(p):
        with pytest.raises(tb.TracerBoyError) as e:
            tb.convert_scene(p, str(tmp_path / "out.tbscene"))
        return e.value.code == -3
    tb.convert_scene(mutated(off_mats + 0, "<f", struct.unpack_from("<f", raw, off_mats)[0]), str(tmp_path / "out.tbscene"))  # identity: still loads
    assert rejected(mutated(off_mats + 12, "<I", nt + 5))            # albedoIndex
    assert rejected(mutated(off_mats + 20, "<I", 0x7fffffff))        # normalMapIndex
    assert rejected(mutated(off_mats + 28, "<I", nt))                # specularMapIndex == count
    mats = np.frombuffer(bytes(raw[off_mats:off_mats + nm * 84]), np.uint32).reshape(nm, 21)
    mix = int(np.flatnonzero(mats[:, 19] & 0x8)[0])
    assert rejected(mutated(off_mats + mix * 84 + 0, "<f", float(nm)))   # mix material id in albedo.x
    assert rejected(mutated(off_mats + mix * 84 + 4, "<f", -1.0))        # ... albedo.y
    tex = np.frombuffer(bytes(raw[off_tex:off_tex + nt * 80]), np.uint32).reshape(nt, 20)
    img_tex = int(np.flatnonzero(tex[:, 0] == 0)[0])
    assert rejected(mutated(off_tex + img_tex * 80 + 4, "<I", nimg))     # DescriptorHeapIndex
    scale_tex = int(np.flatnonzero(tex[:, 0] == 2)[0])
    assert rejected(mutated(off_tex + scale_tex * 80 + 48, "<I", nt + 1))  # TextureIndex1
    assert rejected(mutated(8 + 4 + 4 + 7 * 4, "<i", nimg))              # envImage
    assert rejected(mutated(8 + 4 + 4, "<I", 0x40000000))                # numGeoms far beyond the file size: no giant resize
    off_img0 = off_tex + nt * 80 + nm * 64
    assert rejected(mutated(off_img0, "<I", struct.unpack_from("<I", raw, off_img0)[0] + 1))  # image width vs its pixel bytes


def test_cornell_flatten_matches_reference_scene(cornell):
    """LoadScene flatten of Scenes/cornell-box: 8 shapes, 36 triangles, 2 per-triangle area lights."""
    from oracle.binding import Oracle
    import struct
    raw = open(cornell, "rb").read()
    magic, version, flip, ng, nv, ni, nm, nl, nt, nimg, env = struct.unpack_from("<8sII7Ii", raw, 0)
    assert magic == b"TBSCENE1" and (ng, ni // 3, nl, nt, nimg, env) == (8, 36, 2, 0, 0, -1)
    assert flip == 1 and nm == 8
    cam = np.frombuffer(raw, np.float32, 14, 48)
    lens_height, focal = cam[12], cam[13]
    assert abs(lens_height - 2.0) < 1e-6
    assert abs(focal - 1.0 / np.tan(np.radians(19.5) / 2)) < 1e-4   # (LensHeight/2)/tan(fov/2), TracerBoy.cpp:1259-1260
    o = Oracle(); o.LoadScene(cornell, 3)
    assert o.NumTriangles() == 36


SHARD_SCRIPT = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import tracerboy_b200 as tb
from oracle.binding import Oracle
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
s = tb.get_default_output_settings(); s.MaxBounces = 3
o = Oracle(); o.LoadScene(%(scene)r, 3); o.Resize(24, 24)
o.SetFrameShard(rank, 2)            # frame f on rank f mod N, as bench.py / tb_set_frame_shard do
o.Render(s, 3, 0.0)                 # this rank's 3 of the 6 frames
acc = torch.from_numpy(o.Readback(0).copy())
gathered = [torch.empty_like(acc) for _ in range(2)]
dist.all_gather(gathered, acc)
total = gathered[0] + gathered[1]   # fixed rank order => deterministic
if rank == 0:
    ref = Oracle(); ref.LoadScene(%(scene)r, 3); ref.Resize(24, 24); ref.Render(s, 6, 0.0)
    want = ref.Readback(0)
    assert np.array_equal(total.numpy()[..., 3], want[..., 3])
    assert np.allclose(total.numpy(), want, rtol=1e-5, atol=1e-6), np.abs(total.numpy() - want).max()
    print("SHARD_OK")
dist.destroy_process_group()
'''


def test_sample_sharding_world_size_2_gloo(cornell, tmp_path):
    """The N>1 path: interleaved sample indices + fixed-order sum == single-process accumulation
    (up to float summation order). Two processes over gloo on the CPU."""
    script = tmp_path / "shard.py"
    script.write_text(SHARD_SCRIPT % {"root": ROOT, "port": 29000 + os.getpid() % 2000, "scene": cornell})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "SHARD_OK" in outs[0]


ROW_SHARD_SCRIPT = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import tracerboy_b200 as tb
from oracle.binding import Oracle
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
s = tb.get_default_output_settings(); s.MaxBounces = 3
o = Oracle(); o.LoadScene(%(scene)r, 3); o.Resize(40, 44)      # 5.5 bands of 8 rows
o.SetRowShard(rank, 2)              # band b on rank b mod N, as bench.py --shard rows / tb_set_row_shard do
o.Render(s, 4, 0.0)                 # all 4 frames, this rank's bands only
acc = torch.from_numpy(o.Readback(0).copy())
own = ((np.arange(44) // 8) %% 2) == rank
assert not acc.numpy()[~own].any() and acc.numpy()[own][..., 3].all()
dist.all_reduce(acc)                # the path's one exchange step: every pixel is non-zero on exactly one rank
if rank == 0:
    ref = Oracle(); ref.LoadScene(%(scene)r, 3); ref.Resize(40, 44); ref.Render(s, 4, 0.0)
    want = ref.Readback(0)
    assert np.array_equal(acc.numpy().view(np.uint32), want.view(np.uint32))   # x + 0 = x: bit-identical to one process
    print("ROW_SHARD_OK")
dist.destroy_process_group()
'''


def test_row_band_sharding_world_size_2_gloo(cornell, tmp_path):
    """The other partitioning of SURVEY 8e: bands of 8 rows interleaved over the ranks, one all-reduce of the accumulation
    buffer; every pixel is computed entirely by one rank, so the sum is bit-identical to the single-process image.
    Two processes over gloo on the CPU (the GPU version of this property is test_row_band_sharding_is_bit_identical)."""
    script = tmp_path / "rows.py"
    script.write_text(ROW_SHARD_SCRIPT % {"root": ROOT, "port": 31000 + os.getpid() % 2000, "scene": cornell})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "ROW_SHARD_OK" in outs[0]


CURVES_PBRT = """LookAt 0 0 8  1.5 0.5 0  0 1 0
Camera "perspective" "float fov" [40]
WorldBegin
LightSource "distant" "point from" [0 5 5] "point to" [0 0 0] "rgb L" [3 3 3]
Material "matte" "rgb Kd" [0.5 0.4 0.3]
Shape "curve" "string type" "cylinder" "point P" [0 0 0  1 1 0  2 -1 0  3 0 0] "float width0" 0.1 "float width1" 0.05
Shape "curve" "string type" "cylinder" "point P" [0 1 0  1 2 0  2 0 0  3 1 0  4 2 1] "float width0" 0.2 "float width1" 0.2
Material "matte" "rgb Kd" [0.1 0.9 0.3]
Shape "curve" "string type" "cylinder" "point P" [0 -1 0  1 -1 1  2 -1 -1  3 -1 0] "float width0" 0.3 "float width1" 0.1
Shape "trianglemesh" "point P" [-5 -2 -5  5 -2 -5  5 -2 5  -5 -2 5] "integer indices" [0 1 2 0 2 3]
WorldEnd
"""


def _read_tbscene_geometry(path):
    import struct
    raw = open(path, "rb").read()
    magic, version, flip, ng, nv, ni, nm, nl, nt, nimg, env = struct.unpack_from("<8sII7Ii", raw, 0)
    off = 8 + 4 + 4 + 7 * 4 + 4 + 14 * 4 + 3 * 16 + 12 + 8 * 4
    geoms = np.frombuffer(raw, np.uint32, ng * 8, off).reshape(ng, 8); off += ng * 32
    pos = np.frombuffer(raw, np.float32, nv * 3, off).reshape(nv, 3); off += nv * 12
    vtx = np.frombuffer(raw, np.float32, nv * 8, off).reshape(nv, 8); off += nv * 32
    idx = np.frombuffer(raw, np.uint32, ni, off)
    return geoms, pos, vtx, idx


def test_curve_tessellation_matches_the_reference_rules(tmp_path, built):
    """LoadScene's hair path (TracerBoy.cpp:1426-1524, Curves.cpp): three 3-vertex rings per cubic segment, ring faces
    that always index the first segment's rings, the (1.0 - 1) tangent term, and the ten-pass merge loop that
    re-tessellates a curve without a mergeable successor. Positions against an independent float64 restatement."""
    import tracerboy_b200 as tb
    from tracerboy_b200 import build
    if not os.path.exists(os.path.join(build.LIB, "libtb_pbrtimport.so")):
        pytest.skip("PBRT importer not built (needs the reference mount at build time)")
    src, dst = str(tmp_path / "c.pbrt"), str(tmp_path / "c.tbscene")
    open(src, "w").write(CURVES_PBRT)
    tb.convert_scene(src, dst)
    geoms, pos, vtx, idx = _read_tbscene_geometry(dst)
    # geometry 0: curve A once + curve B on each of the nine remaining passes; geometry 1: the lone curve C ten times
    assert geoms.shape[0] == 3
    assert (geoms[0][4] // 3, geoms[1][4] // 3, geoms[2][4] // 3) == (12 + 9 * 24, 10 * 12, 2)
    assert (geoms[0][2], geoms[1][2]) == (9 + 9 * 18, 10 * 9)
    assert geoms[0][0] != geoms[1][0]  # two materials

    def rings(P, w0, w1):
        P = np.asarray(P, np.float64).reshape(-1, 3)
        nseg = len(P) - 3
        out = []
        for seg in range(nseg):
            p0, p1, p2, p3 = P[seg:seg + 4]
            for ring in range(3):
                t = seg / nseg + ring / nseg / 3
                radius = t * w1 + (1 - t) * w0
                q = lambda a, b, c: (a * (1 - t) + b * t) * (1 - t) + (b * (1 - t) + c * t) * t
                centre = q(p0, p1, p2) * (1 - t) + q(p1, p2, p3) * t
                tangent = (p2 - p1) * 6 * t * (1 - t) + (p3 - p2) * 3 * t * t   # first term: * (1.0 - 1)
                with np.errstate(invalid="ignore"):
                    fwd = tangent / np.linalg.norm(tangent)
                up = np.array([0, 1, 0.0]) if fwd[1] < 0.99 else np.array([0, 0, 1.0])
                n0 = np.cross(fwd, up); n0 /= np.linalg.norm(n0)
                n1 = np.cross(n0, fwd); n1 /= np.linalg.norm(n1)
                for v in range(3):
                    th = v * (3.14 * 2.0 / 3)
                    out.append(centre + (np.cos(th) * n0 + np.sin(th) * n1) * radius)
        return np.array(out)
    first = int(geoms[1][1])
    want = rings([0, -1, 0, 1, -1, 1, 2, -1, -1, 3, -1, 0], 0.3, 0.1)
    got = pos[first:first + 9]
    # ring 0 sits at t = 0 where the reference's tangent is exactly zero: normalize(0) = NaN, reproduced as is
    assert np.isnan(got[:3]).all() and np.isnan(want[:3]).all()
    assert np.allclose(got[3:], want[3:], atol=1e-5)
    assert np.array_equal(pos[first + 9:first + 18][3:], got[3:])  # second pass: the same curve again
    # faces of the first tessellation: ring r joins ring r-1, vertex offsets relative to this pass
    f = idx[int(geoms[1][3]):int(geoms[1][3]) + 36].reshape(12, 3)
    assert f[0].tolist() == [3, 4, 0] and f[1].tolist() == [4, 0, 1] and f[5].tolist() == [3, 2, 0] and f[6].tolist() == [6, 7, 3]
    # segment 1 of curve B reuses segment 0's ring indices (loopStartIndex ignores curveIndex)
    gb = idx[int(geoms[0][3]) + 36:int(geoms[0][3]) + 36 + 72].reshape(24, 3)
    assert np.array_equal(gb[:12], gb[12:])


def test_image_writers_roundtrip(tmp_path, built):
    """tb_write_image: .png decodes (PIL) to the same bytes; .exr (scanline, uncompressed, float) and .pfm parse back
    to the same floats with an independent reader."""
    import struct
    import tracerboy_b200 as tb
    rng = np.random.default_rng(0)
    h, w = 37, 53
    u8 = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    tb.write_image(tmp_path / "a.png", u8)
    try:
        from PIL import Image
        assert np.array_equal(np.asarray(Image.open(tmp_path / "a.png").convert("RGBA")), u8)
    except ImportError:
        pass
    raw = open(tmp_path / "a.png", "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n" and raw[12:16] == b"IHDR" and struct.unpack(">II", raw[16:24]) == (w, h)
    import zlib
    pos, idat = 8, b""
    while pos < len(raw):
        n, typ = struct.unpack(">I4s", raw[pos:pos + 8])
        body = raw[pos + 8:pos + 8 + n]
        assert zlib.crc32(typ + body) == struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])[0]
        if typ == b"IDAT":
            idat += body
        pos += 12 + n
    rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 4 * w)
    assert (rows[:, 0] == 0).all() and np.array_equal(rows[:, 1:].reshape(h, w, 4), u8)

    for ch in (3, 4):
        f = rng.normal(0, 10, (h, w, ch)).astype(np.float32)
        f[0, 0, 0] = np.inf
        tb.write_image(tmp_path / "b.exr", f)
        raw = open(tmp_path / "b.exr", "rb").read()
        assert struct.unpack("<II", raw[:8]) == (20000630, 2)
        pos, attrs = 8, {}
        while raw[pos] != 0:
            e = raw.index(b"\0", pos); name = raw[pos:e].decode(); pos = e + 1
            e = raw.index(b"\0", pos); typ = raw[pos:e].decode(); pos = e + 1
            size = struct.unpack("<I", raw[pos:pos + 4])[0]; pos += 4
            attrs[name] = (typ, raw[pos:pos + size]); pos += size
        pos += 1
        assert attrs["compression"][1] == b"\0" and struct.unpack("<4i", attrs["dataWindow"][1]) == (0, 0, w - 1, h - 1)
        names = [c.split(b"\0")[0].decode() for c in [attrs["channels"][1][i * 18:(i + 1) * 18] for i in range(ch)]]
        assert names == (["A", "B", "G", "R"] if ch == 4 else ["B", "G", "R"])
        offsets = struct.unpack("<%dQ" % h, raw[pos:pos + 8 * h])
        got = np.empty((h, w, ch), np.float32)
        for y in range(h):
            yy, nbytes = struct.unpack("<iI", raw[offsets[y]:offsets[y] + 8])
            assert yy == y and nbytes == 4 * w * ch
            line = np.frombuffer(raw, np.float32, w * ch, offsets[y] + 8).reshape(ch, w)
            for i, nm in enumerate(names):
                got[y, :, "RGBA".index(nm)] = line[i]
        assert np.array_equal(got.view(np.uint32), f.view(np.uint32))
        tb.write_image(tmp_path / "c.pfm", f)
        raw = open(tmp_path / "c.pfm", "rb").read()
        hdr = b"PF\n%d %d\n-1.0\n" % (w, h)
        assert raw.startswith(hdr)
        back = np.frombuffer(raw, np.float32, w * h * 3, len(hdr)).reshape(h, w, 3)[::-1]
        assert np.array_equal(back.view(np.uint32), f[..., :3].view(np.uint32))
    with pytest.raises(tb.TracerBoyError):
        tb.write_image(tmp_path / "x.jpg", u8)
    with pytest.raises(tb.TracerBoyError):
        tb.write_image(tmp_path / "x.png", u8[..., :3])


def _camera(pos=(0, 1, 5), look=(0, 1, 4), right=(1, 0, 0), up=(0, 1, 0)):
    import tracerboy_b200 as tb
    c = tb.Camera()
    for name, v in (("Position", pos), ("LookAt", look), ("Right", right), ("Up", up)):
        f = getattr(c, name)
        f.x, f.y, f.z = v
    c.LensHeight, c.FocalDistance = 2.0, 7.0
    return c


def _v(f):
    return np.array([f.x, f.y, f.z], np.float64)


def test_camera_update_follows_tracerboy_update(built):
    """tb_camera_update against TracerBoy::Update (TracerBoy.cpp:3386-3500), evaluated here in float64:
    mouse look = yaw about +Y by 0.5 * 2 * 6.28 * dx / width then pitch about the XZ-aligned right axis by
    0.5 * 3.14 * dy / height; WASD/QE move Position and LookAt by dt * speed along view / right / up; sticks beyond the
    0.2 dead zone scale the motion; the frame is re-orthonormalised; nothing is stored unless the camera moved.
    Tolerance 1e-5 (float32 trigonometry; DirectXMath's own polynomial sin/cos differs at that level too)."""
    import tracerboy_b200 as tb
    W, H = 1920, 1080

    def rot(v, axis, a):
        n = axis / np.linalg.norm(axis)
        return v * np.cos(a) + np.cross(n, v) * np.sin(a) + n * np.dot(n, v) * (1 - np.cos(a))

    # 1) no input at the stored mouse position: not moved, camera untouched (even a non-orthonormal one)
    cam = _camera(right=(2, 0, 0))
    last = (C.c_uint32 * 2)(10, 20)
    assert tb.camera_update(cam, last, W, H, 10, 20) is False
    assert _v(cam.Right).tolist() == [2, 0, 0] and _v(cam.Position).tolist() == [0, 1, 5]

    # 2) mouse look
    cam = _camera()
    last = (C.c_uint32 * 2)(100, 100)
    assert tb.camera_update(cam, last, W, H, 340, 40) is True
    assert (last[0], last[1]) == (340, 40)
    yaw, pitch = 0.5 * 2.0 * 6.28 * 240 / W, 0.5 * 3.14 * -60 / H
    view = rot(rot(np.array([0.0, 0, -1]), np.array([0.0, 1, 0]), yaw), np.array([1.0, 0, 0]), pitch)
    view /= np.linalg.norm(view)
    right = np.cross([0, 1, 0], view); right /= np.linalg.norm(right)
    up = np.cross(view, right); up /= np.linalg.norm(up)
    assert np.allclose(_v(cam.LookAt) - _v(cam.Position), view, atol=1e-5)
    assert np.allclose(_v(cam.Right), right, atol=1e-5) and np.allclose(_v(cam.Up), up, atol=1e-5)
    assert np.allclose(_v(cam.Position), [0, 1, 5])
    assert cam.LensHeight == 2.0 and cam.FocalDistance == 7.0

    # 3) m_ignoreMouse: the mouse position is remembered but neither rotates nor counts as movement
    cam = _camera()
    last = (C.c_uint32 * 2)(0, 0)
    ignore = tb.CameraSettings(3.0, 1)
    assert tb.camera_update(cam, last, W, H, 500, 500, cameraSettings=ignore) is False
    assert (last[0], last[1]) == (500, 500) and _v(cam.LookAt).tolist() == [0, 1, 4]

    # 4) keys: W forward, D right, Q up, and their opposites cancel
    for keys, want in (("w", (0, 0, -1)), ("S", (0, 0, 1)), ("d", (-1, 0, 0)), ("A", (1, 0, 0)), ("q", (0, 1, 0)), ("E", (0, -1, 0)),
                       ("ws", (0, 0, 0)), ("wd", (-1, 0, -1))):
        cam = _camera()
        last = (C.c_uint32 * 2)(7, 7)
        assert tb.camera_update(cam, last, W, H, 7, 7, keys, 0.25, cameraSettings=tb.CameraSettings(2.0, 0)) is True
        step = 0.25 * 2.0 * np.array(want, np.float64)
        assert np.allclose(_v(cam.Position), np.array([0, 1, 5]) + step, atol=1e-6), keys
        assert np.allclose(_v(cam.LookAt), np.array([0, 1, 4]) + step, atol=1e-6), keys
        # right = normalize(cross(+Y, view)): for view = -Z that is -X (the reference's convention)
        assert np.allclose(_v(cam.Right), [-1, 0, 0], atol=1e-6) and np.allclose(_v(cam.Up), [0, 1, 0], atol=1e-6)

    # 5) controller: dead zone, stick-scaled motion, right stick rotation with its 0.001 * dt scale
    cam = _camera()
    last = (C.c_uint32 * 2)(0, 0)
    assert tb.camera_update(cam, last, 0, 0, 0, 0, None, 1.0, tb.ControllerState(0.1, -0.2, 0.2, 0.15, 0.2, 0.0)) is False
    cam = _camera()
    cs = tb.ControllerState(0.0, 0.0, 0.5, 0.0, -0.5, 0.0)   # right trigger: up, left stick Y: backwards at half speed
    assert tb.camera_update(cam, last, 0, 0, 0, 0, None, 2.0, cs) is True     # default settings: speed 1
    assert np.allclose(_v(cam.Position), [0, 1 + 2.0 * 0.5, 5 + 2.0 * 0.5], atol=1e-6)
    cam = _camera()
    cs = tb.ControllerState(0.8, 0.0, 0, 0, 0, 0)
    assert tb.camera_update(cam, last, 0, 0, 0, 0, None, 100.0, cs) is True
    view = rot(np.array([0.0, 0, -1]), np.array([0.0, 1, 0]), 0.8 * 0.001 * 100.0)
    assert np.allclose(_v(cam.LookAt) - _v(cam.Position), view, atol=1e-5)

    # 6) argument checking
    lib = tb.load_library()
    assert lib.tb_camera_update(None, last, 0, 0, 0, 0, None, 0.0, None, None, None) != 0
    assert lib.tb_update(None, 0, 0, None, 0.0, None, None) != 0


MATERIALS_PBRT = """LookAt 0 0 8  0 0 0  0 1 0
Camera "perspective" "float fov" [40]
WorldBegin
Material "disney" "rgb color" [0.9 0.5 0.5] "float roughness" 0.3 "float eta" 1.4 "float metallic" 0.8
Shape "trianglemesh" "point P" [0 0 0  1 0 0  0 1 0] "integer indices" [0 1 2]
Material "disney" "rgb color" [0.5 0.6 0.7] "float spectrans" 0.5 "float roughness" 0.4 "float eta" 1.2
Shape "trianglemesh" "point P" [0 0 1  1 0 1  0 1 1] "integer indices" [0 1 2]
Material "uber" "rgb Kd" [0.1 0.2 0.3] "float roughness" 0.25 "rgb opacity" [0.5 0.5 0.5] "float index" 1.33 "rgb Kt" [0.2 0.3 0.4]
Shape "trianglemesh" "point P" [0 0 2  1 0 2  0 1 2] "integer indices" [0 1 2]
Material "uber" "rgb Kd" [0.3 0.2 0.1] "float roughness" 0.25 "float uroughness" 0.1 "float vroughness" 0.1
Shape "trianglemesh" "point P" [0 0 3  1 0 3  0 1 3] "integer indices" [0 1 2]
Material "mirror" "rgb Kr" [0.9 0.8 0.7]
Shape "trianglemesh" "point P" [0 0 4  1 0 4  0 1 4] "integer indices" [0 1 2]
Material "metal" "rgb eta" [0.2 0.9 1.1] "float uroughness" 0.05 "float vroughness" 0.05
Shape "trianglemesh" "point P" [0 0 5  1 0 5  0 1 5] "integer indices" [0 1 2]
Material "substrate" "rgb Kd" [0.4 0.5 0.6] "rgb Ks" [0.04 0.04 0.04] "float uroughness" 0.1 "float vroughness" 0.1
Shape "trianglemesh" "point P" [0 0 6  1 0 6  0 1 6] "integer indices" [0 1 2]
Material "glass" "float index" 1.7
Shape "trianglemesh" "point P" [0 0 7  1 0 7  0 1 7] "integer indices" [0 1 2]
Material "matte" "rgb Kd" [0.7 0.6 0.5] "float sigma" 20
Shape "trianglemesh" "point P" [0 0 8  1 0 8  0 1 8] "integer indices" [0 1 2]
Material "plastic" "rgb Kd" [0.2 0.4 0.6] "rgb Ks" [0.25 0.25 0.25] "float roughness" 0.15
Shape "trianglemesh" "point P" [0 0 9  1 0 9  0 1 9] "integer indices" [0 1 2]
Material "translucent" "rgb Kd" [0.2 0.2 0.2]
Shape "trianglemesh" "point P" [0 0 10  1 0 10  0 1 10] "integer indices" [0 1 2]
Material "hair"
Shape "trianglemesh" "point P" [0 0 11  1 0 11  0 1 11] "integer indices" [0 1 2]
MakeNamedMaterial "a" "string type" "matte" "rgb Kd" [0.1 0.1 0.1]
MakeNamedMaterial "b" "string type" "mirror" "rgb Kr" [1 1 1]
MakeNamedMaterial "m" "string type" "mix" "string namedmaterial1" "a" "string namedmaterial2" "b" "rgb amount" [0.3 0.3 0.3]
NamedMaterial "m"
Shape "trianglemesh" "point P" [0 0 12  1 0 12  0 1 12] "integer indices" [0 1 2]
AttributeBegin
AreaLightSource "diffuse" "rgb L" [5 4 3]
Material "matte" "rgb Kd" [0.5 0.5 0.5]
Shape "trianglemesh" "point P" [0 0 13  1 0 13  0 1 13] "integer indices" [0 1 2]
AttributeEnd
WorldEnd
"""


def test_create_material_rules(tmp_path, built):
    """CreateMaterial (TracerBoy.cpp:273-505) through the PBRT importer, one shape per rule: disney (albedo.x >
    0.7 forced to 0.2, metallic > 0.5, specTrans => subsurface with roughness 0), uber (uroughness wins, opacity < 1 =>
    subsurface + single-sided with IOR = index and absorption = Kt), mirror, metal (white albedo, IOR = average eta),
    substrate / plastic (IOR = (sqrt(ks) + 1) / (1 - sqrt(ks)), SpecularCoef = ks), glass, matte (NO_SPECULAR, sigma as
    roughness), translucent without a map, an unsupported type (default brown), mix (sub-material indices and amount in
    albedo, +2 materials), the LIGHT flag from an area light's emission, and NO_ALPHA everywhere (no alpha textures)."""
    import struct
    import tracerboy_b200 as tb
    from tracerboy_b200 import build
    if not os.path.exists(os.path.join(build.LIB, "libtb_pbrtimport.so")):
        pytest.skip("PBRT importer not built (needs the reference mount at build time)")
    src, dst = str(tmp_path / "m.pbrt"), str(tmp_path / "m.tbscene")
    open(src, "w").write(MATERIALS_PBRT)
    tb.convert_scene(src, dst)
    raw = open(dst, "rb").read()
    magic, version, flip, ng, nv, ni, nm, nl, nt, nimg, env = struct.unpack_from("<8sII7Ii", raw, 0)
    off = 8 + 4 + 4 + 7 * 4 + 4 + 14 * 4 + 3 * 16 + 12 + 8 * 4
    geoms = np.frombuffer(raw, np.uint32, ng * 8, off).reshape(ng, 8)
    off += ng * 32 + nv * 12 + nv * 32 + ni * 4
    M = np.dtype([("albedo", "3f4"), ("albedoIndex", "u4"), ("alphaIndex", "u4"), ("normalMapIndex", "u4"), ("emissiveIndex", "u4"),
                  ("specularMapIndex", "u4"), ("IOR", "f4"), ("absorption", "3f4"), ("roughness", "f4"), ("scattering", "3f4"),
                  ("emissive", "3f4"), ("Flags", "i4"), ("SpecularCoef", "f4")])
    assert M.itemsize == 84
    mats = np.frombuffer(raw, M, nm, off)
    assert ng == 14 and nm == 16 and nt == 0           # 13 materials + 2 sub-materials of the mix + the emissive matte
    METAL, SSS, NOSPEC, MIX, LIGHT, NOALPHA, SINGLE = 0x1, 0x2, 0x4, 0x8, 0x10, 0x20, 0x80
    f = np.float32

    def ior(ks):
        return f((np.sqrt(ks) + 1.0) / (1.0 - np.sqrt(ks)))
    # per shape, in file order: albedo, IOR, roughness, flags, SpecularCoef, absorption
    want = [((0.2, 0.2, 0.2), 1.4, 0.3, METAL | NOALPHA, 0.0, (0, 0, 0)),                      # disney, bright and metallic
            ((0.5, 0.6, 0.7), 1.2, 0.0, SSS | NOALPHA, 0.0, (0, 0, 0)),                        # disney, specTrans
            ((0.1, 0.2, 0.3), 1.33, 0.25, SSS | SINGLE | NOALPHA, 0.0, (0.2, 0.3, 0.4)),       # uber, opacity 0.5
            ((0.3, 0.2, 0.1), 1.5, 0.1, NOALPHA, 0.0, (0, 0, 0)),                              # uber, uroughness
            ((0.9, 0.8, 0.7), 1.5, 0.0, METAL | NOALPHA, 1.0, (0, 0, 0)),                      # mirror
            ((1.0, 1.0, 1.0), f((0.2 + 0.9 + 1.1) / 3.0), 0.05, METAL | NOALPHA, 0.0, (0, 0, 0)),   # metal
            ((0.4, 0.5, 0.6), ior(0.04), 0.1, NOALPHA, 0.04, (0, 0, 0)),                       # substrate
            ((0.0, 0.0, 0.0), 1.7, 0.0, SSS | NOALPHA, 0.0, (0, 0, 0)),                        # glass
            ((0.7, 0.6, 0.5), 1.5, 20.0, NOSPEC | NOALPHA, 0.0, (0, 0, 0)),                    # matte
            ((0.2, 0.4, 0.6), ior(0.25), 0.15, NOALPHA, 0.25, (0, 0, 0)),                      # plastic
            ((0.0, 0.0, 0.0), 1.5, 0.0, SSS | NOALPHA, 0.0, (0.001, 0.001, 0.001)),            # translucent, no map
            ((153.0 / 255.0, 102.0 / 255.0, 58.0 / 255.0), 1.5, 0.2, NOALPHA, 0.0, (0, 0, 0))]  # unsupported type
    for shape, (alb, io, rough, flags, spec, absorb) in enumerate(want):
        m = mats[geoms[shape][0]]
        assert np.allclose(m["albedo"], np.array(alb, f), rtol=0, atol=1e-7), (shape, m["albedo"])
        assert np.isclose(m["IOR"], f(io), rtol=1e-6) and np.isclose(m["roughness"], f(rough), rtol=1e-6), (shape, m["IOR"], m["roughness"])
        assert m["Flags"] == flags, (shape, hex(m["Flags"]))
        assert np.isclose(m["SpecularCoef"], f(spec), rtol=1e-6) and np.allclose(m["absorption"], np.array(absorb, f)), shape
        assert not m["emissive"].any() and not m["scattering"].any()
        for k in ("albedoIndex", "alphaIndex", "normalMapIndex", "emissiveIndex", "specularMapIndex"):
            assert m[k] == 0xffffffff
    mix = mats[geoms[12][0]]
    assert mix["Flags"] == MIX | NOALPHA and np.isclose(mix["albedo"][2], f(0.3))
    a, b = mats[int(mix["albedo"][0])], mats[int(mix["albedo"][1])]
    assert np.allclose(a["albedo"], f(0.1)) and a["Flags"] == NOSPEC | NOALPHA
    assert np.allclose(b["albedo"], 1.0) and b["Flags"] == METAL | NOALPHA and b["SpecularCoef"] == 1.0
    light = mats[geoms[13][0]]
    assert light["Flags"] == LIGHT | NOSPEC | NOALPHA and light["emissive"].tolist() == [5.0, 4.0, 3.0]
    assert nl == 1                                      # its one triangle is the scene's one light (TracerBoy.cpp:1780-1835)


def _flatten_case(which, tmp_path):
    """(scene path, insert-instances flag) of a flatten test case."""
    if which == "materials":
        src = str(tmp_path / "m.pbrt"); open(src, "w").write(MATERIALS_PBRT)
    elif which == "instanced":
        src = str(tmp_path / "i.pbrt"); open(src, "w").write(INSTANCED_PBRT)
    elif which == "textured":
        from test_cpu_images import write_textured_scene
        src = write_textured_scene(str(tmp_path))
    else:
        src = "/root/reference/Scenes/%s/scene.pbrt" % {"cornell": "cornell-box", "teapot": "Teapot"}[which]
        if not os.path.exists(src):
            pytest.skip("reference mount not present")
    return src, 1 if which == "instanced" else 0


def _tbscene_arrays(path):
    """materials [nm, 21] u32, per-geometry material index, lights [nl, 26] u32, texture count of a .tbscene."""
    import struct
    raw = open(path, "rb").read()
    magic, version, flip, ng, nv, ni, nm, nl, nt, nimg, env = struct.unpack_from("<8sII7Ii", raw, 0)
    geoms = np.frombuffer(raw, np.uint32, ng * 8, 196).reshape(ng, 8)
    off = 196 + ng * 32 + nv * 12 + nv * 32 + ni * 4
    mats = np.frombuffer(raw, np.uint32, nm * 21, off).reshape(nm, 21)
    off += nm * 84
    lights = np.frombuffer(raw, np.uint32, nl * 26, off).reshape(nl, 26)
    return mats, geoms[:, 0].copy(), lights, nt


INSTANCED_PBRT = """
LookAt 0 2 -12  0 0 0  0 1 0
Camera "perspective" "float fov" [40]
Film "image" "integer xresolution" [64] "integer yresolution" [48]
WorldBegin
MakeNamedMaterial "red" "string type" ["matte"] "rgb Kd" [0.8 0.1 0.1]
MakeNamedMaterial "steel" "string type" ["metal"] "float roughness" [0.2]
# object a: a smooth-shaded quad followed by a second shape that LoadScene ignores (it instantiates shapes[0] only)
ObjectBegin "a"
  NamedMaterial "red"
  Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-1 0 -1  1 0 -1  1 0.5 1  -1 0.5 1]
        "normal N" [0 1 0  0 1 0  0.3 0.9 0.1  -0.3 0.9 0.1] "float uv" [0 0 1 0 1 1 0 1]
  Shape "trianglemesh" "integer indices" [0 1 2] "point P" [5 5 5  6 5 5  5 6 5]
ObjectEnd
# object b: no normals (flat-normal rule under a rotation and a non-uniform scale), one degenerate face
ObjectBegin "b"
  NamedMaterial "steel"
  Shape "trianglemesh" "integer indices" [0 1 2 0 2 3 0 0 1] "point P" [-1 -1 0  1 -1 0  1 1 0.25  -1 1 0]
ObjectEnd
# object c: an emissive triangle (its light is built from the UNtransformed vertices, TracerBoy.cpp:1538-1540)
ObjectBegin "c"
  AttributeBegin
    AreaLightSource "diffuse" "rgb L" [4 3 2]
    NamedMaterial "red"
    Shape "trianglemesh" "integer indices" [0 1 2] "point P" [0 0 0  1 0 0  0 0 1]
  AttributeEnd
ObjectEnd
NamedMaterial "red"
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-20 -2 -20  20 -2 -20  20 -2 20  -20 -2 20]
AttributeBegin
  Translate 3 0.5 0
  ObjectInstance "a"
AttributeEnd
AttributeBegin
  Translate -3 1 2
  Rotate 37 0.3 1 0.2
  Scale 1.5 0.5 2
  ObjectInstance "a"
AttributeEnd
AttributeBegin
  Rotate -70 1 0 0
  Scale 2 1 3
  Translate 0.5 0 -1
  ObjectInstance "b"
AttributeEnd
AttributeBegin
  Translate 0 6 0
  Rotate 180 1 0 0
  ObjectInstance "c"
AttributeEnd
WorldEnd
"""


@pytest.mark.parametrize("which", ["materials", "textured", "cornell", "teapot", "instanced"])
def test_flatten_equals_the_reference_rules_compiled_from_the_mount(which, tmp_path, built):
    """An independent check of the scene flatten: CreateMaterial, MaterialTracker and the per-triangle
    area-light rule of LoadScene compiled from the reference's own TracerBoy.cpp / TracerBoy.h / SharedShaderStructs.h
    (oracle/ref/ref_flatten.cpp -> oracle/_ref/libref_flatten.so, linked against the vendored pbrt-parser) against what
    the product's importer wrote into the .tbscene: every 84-byte Material in tracker order (mix sub-materials
    included), the material index of every shape, every 104-byte Light, the number of textures created."""
    import ctypes as C
    import tracerboy_b200 as tb
    from tracerboy_b200 import build
    from oracle import binding
    lib_path = os.path.join(os.path.dirname(binding.ref_traverse_lib_path()), "libref_flatten.so")
    if not os.path.exists(lib_path) or not os.path.exists(os.path.join(build.LIB, "libtb_pbrtimport.so")):
        pytest.skip("oracle/_ref/libref_flatten.so or the PBRT importer not built (need the reference mount at build time)")
    src, insert = _flatten_case(which, tmp_path)
    dst = str(tmp_path / "s.tbscene")
    tb.convert_scene(src, dst, tb.INSTANCES_INSERT_INTO_BLAS if insert else tb.INSTANCES_SKIP)
    mats, shape_mat, lights, ntex = _tbscene_arrays(dst)
    ref = C.CDLL(lib_path)
    ref.ref_flatten.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    rm, rs, rl, counts = np.zeros((4096, 21), np.uint32), np.zeros(65536, np.int32), np.zeros((1 << 16, 26), np.uint32), np.zeros(4, np.int32)
    assert ref.ref_flatten(src.encode(), rm.ctypes.data, 4096, rs.ctypes.data, 65536, rl.ctypes.data, 1 << 16, counts.ctypes.data, insert) == 0
    if which == "instanced":
        assert counts[1] == 5 and shape_mat.shape[0] == 5   # the floor + four instances, one geometry each
    nm, ns, nl, nt = (int(x) for x in counts)
    assert nm == mats.shape[0], (nm, mats.shape[0])
    want = rm[:nm].copy()
    # the stand-in texture allocator cannot look into image files: it reports "no alpha" for every map, so NO_ALPHA
    # (0x20) of materials with an albedo map is left to test_pbrt_scene_with_png_and_tga_albedo_maps_flattens
    textured = mats[:, 3] != 0xffffffff
    want[textured, 19] = (want[textured, 19] & ~np.uint32(0x20)) | (mats[textured, 19] & np.uint32(0x20))
    bad = np.argwhere(want != mats)
    assert bad.shape[0] == 0, "material %d word %d: reference %#x, importer %#x" % (bad[0][0], bad[0][1], want[bad[0][0], bad[0][1]], mats[bad[0][0], bad[0][1]])
    meshes = rs[:ns][rs[:ns] >= 0]
    assert np.array_equal(meshes, shape_mat.astype(np.int32)), "per-shape material index"
    area = lights[lights[:, 0] == 0]    # LIGHT_TYPE_AREA; directional lights come from LoadScene's light-source loop (:1896-1917)
    assert nl == area.shape[0] and np.array_equal(rl[:nl], area), "area lights"
    assert nt == ntex


def test_seam_size_queries_and_argument_checks_without_a_device(built):
    """Host-only parts of the SW-RT seam and of the communicator: size queries are pure arithmetic, bad arguments are
    E_INVALIDARG, and nothing here needs (or silently replaces) a GPU."""
    import tracerboy_b200 as tb
    from tracerboy_b200.api import GeometryDesc, PrebuildInfo
    lib = tb.load_library()
    # top level: 16-byte header + 32-byte nodes (2N - 1) + 116-byte BVHMetadata per instance; scratch grows linearly
    for n in (1, 2, 7, 20000):
        info = tb.tlas_prebuild_info(n)
        assert info.ReferenceLayoutSizeInBytes == 16 + 32 * (2 * n - 1) + 116 * n
        assert info.ResultDataMaxSizeInBytes >= info.ReferenceLayoutSizeInBytes + 256 + 80 * n
        assert 320 * n <= info.ScratchDataSizeInBytes <= 324 * n + (1 << 13) and info.ScratchDataSizeInBytes % 256 == 0
    assert tb.tlas_prebuild_info(0).ResultDataMaxSizeInBytes == 0
    info = PrebuildInfo()
    assert lib.tb_tlas_prebuild_info((1 << 24) + 1, C.byref(info)) == -1      # InstanceID / hit-group contribution are 24 bit
    # bottom level: scratch and update scratch grow linearly and are 256-byte granular
    pos = np.zeros((3, 3), np.float32)
    d = (GeometryDesc * 1)()
    d[0].Positions = pos.ctypes.data; d[0].PositionStrideBytes = 12
    sizes = []
    for tris in (1, 1000, 1000000):
        d[0].VertexCount = 3 * tris
        assert lib.tb_bvh_prebuild_info(d, 1, C.byref(info)) == 0
        assert info.UpdateScratchDataSizeInBytes >= 12 * tris - 4 and info.ScratchDataSizeInBytes >= 150 * tris
        sizes.append(info.ScratchDataSizeInBytes)
    assert sizes[0] < sizes[1] < sizes[2] < 200 * 1000000 + (1 << 20)
    d[0].IndexFormat = 3                                                        # only none / uint16 / uint32
    d[0].Indices = pos.ctypes.data
    assert lib.tb_bvh_prebuild_info(d, 1, C.byref(info)) == -1
    # communicator: the id is 128 bytes (ncclUniqueId); calls on a null handle are errors, not crashes
    buf = C.create_string_buffer(64)
    assert lib.tb_comm_get_unique_id(buf, 64) == -1
    assert lib.tb_comm_init(None, buf, 0, 1, 1) == -1 and lib.tb_comm_reduce(None) == -1 and lib.tb_comm_destroy(None) == -1
    assert lib.tb_bvh_build_device(None, d, 1, 0, None, 0, None, 0, None) == -1
    assert lib.tb_tlas_build_device(None, None, 1, 0, None, 0, None, 0, None) == -1
    assert lib.tb_set_material_sort(None, 1) == -1


@pytest.mark.parametrize("which", ["cornell", "teapot", "textured", "materials", "instanced"])
def test_geometry_flatten_equals_the_reference_loops_compiled_from_the_mount(which, tmp_path, built):
    """The geometry half of the flatten: LoadScene's per-vertex loop (positions, normalised normals, uvs,
    tangents with the (0,0,1) default; TracerBoy.cpp:1638-1661) and its index loop with the flat-normal rule for meshes
    without normals (last face to touch a vertex wins, degenerate faces get (0,1,0); :1704-1730), both compiled from the
    mount, against the product importer's pooled arrays in the .tbscene: every position, every 32-byte Vertex, every
    index of every shape, bit for bit."""
    import ctypes as C
    import struct
    import tracerboy_b200 as tb
    from tracerboy_b200 import build
    from oracle import binding
    lib_path = os.path.join(os.path.dirname(binding.ref_traverse_lib_path()), "libref_flatten.so")
    if not os.path.exists(lib_path) or not os.path.exists(os.path.join(build.LIB, "libtb_pbrtimport.so")):
        pytest.skip("oracle/_ref/libref_flatten.so or the PBRT importer not built (need the reference mount at build time)")
    ref = C.CDLL(lib_path)
    if not hasattr(ref, "ref_flatten_geometry"):
        pytest.skip("oracle/_ref/libref_flatten.so predates the geometry entry point")
    src, insert = _flatten_case(which, tmp_path)
    dst = str(tmp_path / "s.tbscene")
    tb.convert_scene(src, dst, tb.INSTANCES_INSERT_INTO_BLAS if insert else tb.INSTANCES_SKIP)
    raw = open(dst, "rb").read()
    magic, version, flip, ng, nv, ni, nm, nl, nt, nimg, env = struct.unpack_from("<8sII7Ii", raw, 0)
    geoms = np.frombuffer(raw, np.uint32, ng * 8, 196).reshape(ng, 8)
    off = 196 + ng * 32
    positions = np.frombuffer(raw, np.uint32, nv * 3, off).reshape(nv, 3); off += nv * 12
    vertices = np.frombuffer(raw, np.uint32, nv * 8, off).reshape(nv, 8); off += nv * 32
    indices = np.frombuffer(raw, np.uint32, ni, off)
    ref.ref_flatten_geometry.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    checked = 0
    for g in range(ng):
        mat, vfirst, vcount, ifirst, icount = (int(x) for x in geoms[g][:5])
        P, V, I, counts = np.zeros((vcount, 3), np.uint32), np.zeros((vcount, 8), np.uint32), np.zeros(icount, np.uint32), np.zeros(2, np.int32)
        assert ref.ref_flatten_geometry(src.encode(), g, P.ctypes.data, V.ctypes.data, I.ctypes.data, vcount, icount, counts.ctypes.data, insert) == 0, g
        assert counts.tolist() == [vcount, icount]
        assert np.array_equal(P, positions[vfirst:vfirst + vcount]), "positions of shape %d" % g
        assert np.array_equal(I, indices[ifirst:ifirst + icount]), "indices of shape %d" % g
        bad = np.argwhere(V != vertices[vfirst:vfirst + vcount])
        assert bad.shape[0] == 0, "shape %d vertex %d word %d (normal 0-2, uv 3-4, tangent 5-7)" % (g, bad[0][0], bad[0][1])
        checked += vcount
    assert checked == nv
    if which == "instanced":   # instances dropped (the reference build's software path) leaves the floor alone
        tb.convert_scene(src, dst)
        assert struct.unpack_from("<8sII7Ii", open(dst, "rb").read(), 0)[3] == 1
