"""BVH build only (for profiling the builder): python tools/build_only.py [workload]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import tracerboy_b200 as tb

wl = sys.argv[1] if len(sys.argv) > 1 else "teapot"
g = tb.TracerBoy(0)
for _ in range(2):
    g.LoadScene(bench.scene_arg(bench.WORKLOADS[wl][0]))
    print(wl, g.GetSceneInfo().NumTriangles, "triangles, BVH build", g.GetBVHBuildMilliseconds(), "ms")
