#!/bin/bash
# BVH build: byte-identity tests against the oracle, then build milliseconds of the bench scenes (four builds each).
python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_seam.py -m gpu -q -k "bvh or dragon or blobs or seam or update or build" 2>&1 | grep -v "pbrt-parser\|created object" | tail -3
for w in teapot dragon vwvan blobs20m; do python - <<PY 2>&1 | grep -v "pbrt-parser\|created object"
import sys; sys.path.insert(0, ".")
import bench, tracerboy_b200 as tb
spec = bench.WORKLOADS["$w"][0]
g = tb.TracerBoy(0); g.LoadScene(bench.scene_arg(spec))
ms = []
for i in range(4):
    g.LoadScene(bench.scene_arg(spec)); ms.append(g.GetBVHBuildMilliseconds())
print("$w", "tris", g.GetSceneInfo().NumTriangles, "build ms", [round(m, 3) for m in ms])
PY
done
