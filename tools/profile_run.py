"""Short render used under ncu (never a bench number): python tools/profile_run.py [workload] [spp]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import tracerboy_b200 as tb

wl = sys.argv[1] if len(sys.argv) > 1 else "teapot"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 2
spec, w, h, _, bounces = bench.WORKLOADS[wl]
g = tb.TracerBoy(0)
g.LoadScene(bench.scene_arg(spec))
g.Resize(w, h)
if os.environ.get("TB_FIF"):
    g.SetFramesInFlight(int(os.environ["TB_FIF"]))
s = tb.get_default_output_settings()
s.MaxBounces = bounces
g.Render(s, 2, 0.0)
g.ResetRenderStats()
g.InvalidateHistory()
g.Render(s, spp, 0.0)
st = g.GetRenderStats()
print("rays", st.RaysTraced, "ms", st.DeviceMilliseconds, "Mrays/s", st.RaysTraced / st.DeviceMilliseconds / 1e3)
