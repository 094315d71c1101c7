"""How many rays does the glass / subsurface walk trace? python tools/walk_stats.py [workload] [spp]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import tracerboy_b200 as tb

wl = sys.argv[1] if len(sys.argv) > 1 else "vwvan"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 4
spec, w, h, _, bounces = bench.WORKLOADS[wl]
g = tb.TracerBoy(0)
g.LoadScene(bench.scene_arg(spec))
g.Resize(w, h)
s = tb.get_default_output_settings()
s.MaxBounces = bounces
g.Render(s, 2, 0.0)
for fif in (1, 8):
    g.SetFramesInFlight(fif)
    g.Render(s, 2, 0.0)
    g.ResetRenderStats(); g.InvalidateHistory()
    g.Render(s, spp, 0.0)
    st = g.GetRenderStats()
    inside = st.RaysTraced - st.ExtendRays - st.ResumeRays
    print("fif", fif, "paths", st.PathsStarted, "rays", st.RaysTraced, "extend", st.ExtendRays, "resume", st.ResumeRays, "walk/inline", inside,
          "boxes", st.BoxesTested, "extend boxes", st.ExtendBoxesTested, "ms", st.DeviceMilliseconds,
          "Mrays/s", st.RaysTraced / st.DeviceMilliseconds / 1e3)
