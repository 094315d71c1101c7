"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / initcheck / racecheck):

    compute-sanitizer --tool memcheck python tools/sanitize_run.py
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tracerboy_b200 as tb

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
g = tb.TracerBoy(0)
s = tb.get_default_output_settings()
# cornell (area lights, inline shadow rays), the showcase scene (glass walk stages, textures, mix, directional light, sky)
# and a larger blob scene with the shadow queue stage forced on
for spec, w, h, shadow in (("tests/golden/cornell-box.tbscene", 96, 64, 2), ("synthetic:showcase?tris=300&seed=1", 96, 56, 1),
                           ("synthetic:blobs?copies=8&tris=500&seed=2", 80, 48, 1)):
    path = os.path.join(root, spec) if spec.endswith(".tbscene") else spec
    g.LoadScene(path)
    g.Resize(w, h)
    g.SetShadowMode(shadow)
    g.SetRaySort(3 if shadow == 1 else 4)  # the queue sort kernels on two of the scenes
    g.SetMaterialSort(1)                   # material-class hit queues (k_class_scatter)
    s.MaxBounces = 5
    s.EnableNormalMaps = 1
    g.Render(s, 3, 0.0)          # frame graphs captured here
    g.Render(s, 2, 0.0)          # and replayed here
    g.SetRowShard(1, 2); g.Render(s, 1, 0.0); g.SetRowShard(0, 1)
    g.SetProfiling(True); g.InvalidateHistory(); g.Render(s, 1, 0.0); g.SetProfiling(False)
    for k in range(10):
        g.Readback(k)
    g.PostProcess(tb.OutputType.LIT, tb.get_default_postprocess_settings())
    g.Readback(tb.BufferKind.BACKBUFFER_RGBA8)
    g.SaveImage(tb.BufferKind.BACKBUFFER_RGBA8, os.path.join(tmp, "a.png"))
    rays = np.zeros(1000, tb.api.RAY_DTYPE)
    rays["Direction"] = np.random.default_rng(0).normal(0, 1, (1000, 3)); rays["TMin"] = 0.001; rays["TMax"] = 1e6
    g.TraceRays(rays)
img = np.random.default_rng(1).random((40, 60, 4)).astype(np.float32)
img[..., 3] = 1
for ot in range(10):
    g.PostProcessImage(img, ot, tb.get_default_postprocess_settings(), img)
p = tb.TemporalAccumulationParams()
p.Camera = g.GetCamera(); p.PrevCamera = g.GetCamera()
p.HistoryWeight, p.IgnoreHistory, p.OutputMomentInformation = 0.95, 0, 1
g.TemporalAccumulateImage(p, img, img, img, img, img, img)
# the SW-RT seam on caller-owned device memory: build, update in place, top level, single- and two-level queries
import ctypes as C
import torch
from tracerboy_b200.api import GeometryDesc, InstanceDesc
rng = np.random.default_rng(2)
pos = torch.from_numpy(rng.uniform(-1, 1, (300, 3)).astype(np.float32)).cuda()
idx = torch.from_numpy(rng.integers(0, 300, 900).astype(np.uint16).view(np.uint8)).cuda()
d = (GeometryDesc * 1)()
d[0].Positions = pos.data_ptr(); d[0].PositionStrideBytes = 12; d[0].VertexCount = 300
d[0].Indices = idx.data_ptr(); d[0].IndexFormat = 2; d[0].IndexCount = 900; d[0].GeometryFlags = 1
info = tb.prebuild_info(d, 1)
dst = torch.zeros(info.ResultDataMaxSizeInBytes, dtype=torch.uint8, device="cuda")
scratch = torch.empty(info.ScratchDataSizeInBytes, dtype=torch.uint8, device="cuda")
g.BuildRaytracingAccelerationStructureDevice(d, 1, dst.data_ptr(), dst.numel(), scratch.data_ptr(), scratch.numel(), None)
pos += 0.01
g.UpdateRaytracingAccelerationStructureDevice(d, 1, dst.data_ptr(), dst.numel(), scratch.data_ptr(), scratch.numel(), None)
inst = (InstanceDesc * 5)()
for i in range(5):
    for j, v in enumerate((1, 0, 0, 3.0 * i, 0, 1, 0, 0, 0, 0, 1, 0)):
        inst[i].Transform[j] = float(v)
    inst[i].InstanceIDAndMask = i | (0xff << 24); inst[i].AccelerationStructure = dst.data_ptr()
tinfo = tb.tlas_prebuild_info(5)
tlas = torch.zeros(tinfo.ResultDataMaxSizeInBytes, dtype=torch.uint8, device="cuda")
tscratch = torch.empty(tinfo.ScratchDataSizeInBytes, dtype=torch.uint8, device="cuda")
g.BuildTopLevelAccelerationStructureDevice(inst, 5, tlas.data_ptr(), tlas.numel(), tscratch.data_ptr(), tscratch.numel(), None)
rays = np.zeros(2000, tb.api.RAY_DTYPE)
rays["Origin"] = rng.uniform(-2, 14, (2000, 3)) * [1, 0.2, 0.2]; rays["Direction"] = rng.normal(0, 1, (2000, 3)); rays["TMin"] = 0.001; rays["TMax"] = 1e6
d_rays = torch.from_numpy(rays.view(np.uint8)).cuda()
d_hits = torch.zeros(2000 * 32, dtype=torch.uint8, device="cuda")
g.TraceRaysDevice(dst.data_ptr(), dst.numel(), d_rays.data_ptr(), 2000, d_hits.data_ptr(), None)
g.TraceRaysTopLevelDevice(tlas.data_ptr(), tlas.numel(), d_rays.data_ptr(), 2000, d_hits.data_ptr(), None)
g.Synchronize()
print("sanitize_run ok")
