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
print("sanitize_run ok")
