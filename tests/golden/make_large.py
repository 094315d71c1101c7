"""Mint tests/golden/blobs20m.npz: what the CPU oracle produces for BASELINE.json configs[4] (20.8 M triangles), so that
the default GPU suite can require bit-identity with the oracle at that size without paying the oracle's
single-threaded 160 s build on every run (tests/test_gpu_scale.py also has the live comparison, marked slow).

  bvh_sha256 / bvh_bytes / depth   the BVH in the reference's byte layout (2.4 GB), hashed
  rays, hits                       32 768 ray queries, every field of the hit records
  accum, primary_hit, counters     one 192x108 sample at 6 bounces (area light, glass walk, sky), counts = (rays, boxes, tris)

    python tests/golden/make_large.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SPEC = "synthetic:blobs?copies=20000&tris=1000&seed=1"
W, H, NRAYS = 192, 108, 32768


def sha256_chunks(a):
    h = hashlib.sha256()
    step = 1 << 26
    flat = a.reshape(-1).view(np.uint8)
    for lo in range(0, flat.size, step):
        h.update(flat[lo:lo + step].tobytes())
    return np.frombuffer(h.digest(), np.uint8)


def main():
    import tracerboy_b200 as tb
    from oracle.binding import Oracle
    from test_gpu_parity import _random_rays
    from test_gpu_scale import _tree_depth
    import tempfile
    path = os.path.join(tempfile.gettempdir(), "tb_blobs20m.tbscene")  # 770 MB: kept out of the repo tree
    if not os.path.exists(path):
        tb.convert_scene(SPEC, path)
    o = Oracle()
    o.LoadScene(path, 3)
    bvh = o.GetBVH()
    out = {"bvh_sha256": sha256_chunks(bvh), "bvh_bytes": np.array([bvh.size], np.uint64), "depth": np.array([_tree_depth(bvh)])}
    del bvh
    rays = _random_rays(NRAYS, o.GetCamera(), 23)
    out["rays"] = rays.view(np.uint8)
    out["hits"] = o.TraceRays(rays).view(np.uint8)
    o.Resize(W, H)
    s = tb.get_default_output_settings()
    o.Render(s, 1, 0.0, threads=os.cpu_count())
    out["accum"] = o.Readback(tb.BufferKind.ACCUM_RGBW)
    out["primary_hit"] = o.Readback(tb.BufferKind.PRIMARY_HIT_IDS)
    out["counters"] = o.Readback(tb.BufferKind.RAY_COUNTERS)
    c = o.Counts()
    out["counts"] = np.array([c["rays"], c["boxes"], c["tris"]], np.uint64)
    np.savez_compressed(os.path.join(HERE, "blobs20m.npz"), **out)
    print("depth", out["depth"], "counts", out["counts"], "hits with t>0:", (o.TraceRays(rays)["t"] > 0).sum())


if __name__ == "__main__":
    main()
