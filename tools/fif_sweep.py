"""Throughput against frames in flight: python tools/fif_sweep.py workload spp fif [fif ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import tracerboy_b200 as tb

wl, spp = sys.argv[1], int(sys.argv[2])
spec, w, h, _, bounces = bench.WORKLOADS[wl]
g = tb.TracerBoy(0)
g.LoadScene(bench.scene_arg(spec))
g.Resize(w, h)
s = tb.get_default_output_settings()
s.MaxBounces = bounces
for fif in [int(a) for a in sys.argv[3:]]:
    g.SetFramesInFlight(fif)
    g.Render(s, max(fif, 2), 0.0)
    g.ResetRenderStats(); g.InvalidateHistory()
    g.Render(s, spp, 0.0)
    st = g.GetRenderStats()
    print(wl, "fif", fif, "ms/frame %.3f" % (st.DeviceMilliseconds / spp), "Mrays/s %.1f" % (st.RaysTraced / st.DeviceMilliseconds / 1e3))
