"""Regenerates tests/golden/*.npz from the CPU oracle (the reference ships no golden vectors
for this path, SURVEY §4, so the pins are minted here and committed with this script).

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import tracerboy_b200 as tb  # noqa: E402  (settings struct + host-only scene conversion)
from oracle.binding import Oracle  # noqa: E402

CASES = {
    # name: (scene, w, h, spp, bounces, settings overrides)
    "cornell_64_default": ("cornell-box.tbscene", 64, 64, 4, 4, {}),
    "cornell_48_nobluenoise": ("cornell-box.tbscene", 48, 48, 3, 5, {"EnableBlueNoise": 0}),
    "blobs_64x36": ("synthetic:blobs?copies=8&tris=200&seed=2", 64, 36, 3, 8, {}),
    "showcase_96x54": ("synthetic:showcase?tris=300&seed=1", 96, 54, 4, 8, {"EnableNormalMaps": 1}),
    # bundled scenes of BASELINE.json configs[1] and [3]: flattened at build time from the reference mount into
    # scenes/_cache (tracerboy_b200/build.py); the cases are skipped where the cache is absent
    "teapot_96x54": ("cache:teapot", 96, 54, 3, 6, {}),
    "vwvan_96x54": ("cache:vw-van", 96, 54, 3, 6, {}),
}


def scene_file(spec):
    if spec.startswith("synthetic:"):
        out = os.path.join(ROOT, "scenes", "_cache", "golden_%s.tbscene" % hashlib.sha1(spec.encode()).hexdigest()[:10])
        os.makedirs(os.path.dirname(out), exist_ok=True)
        tb.convert_scene(spec, out)
        return out
    if spec.startswith("cache:"):
        p = os.path.join(ROOT, "scenes", "_cache", spec[6:] + ".tbscene")
        return p if os.path.exists(p) else None
    return os.path.join(HERE, spec)


def render_case(name):
    spec, w, h, spp, bounces, over = CASES[name]
    if scene_file(spec) is None:
        return None
    o = Oracle()
    o.LoadScene(scene_file(spec), 3)
    o.Resize(w, h)
    s = tb.get_default_output_settings()
    s.MaxBounces = bounces
    for k, v in over.items():
        setattr(s, k, v)
    o.Render(s, spp, 0.0)
    bvh = o.GetBVH()
    out = {
        "accum": o.Readback(0), "jittered": o.Readback(1), "normal": o.Readback(3), "depth": o.Readback(5),
        "albedo": o.Readback(6), "primary_hit": o.Readback(8), "counters": o.Readback(9),
        "bvh_sha256": np.frombuffer(hashlib.sha256(bvh.tobytes()).digest(), np.uint8),
        "bvh_bytes": np.array([bvh.size], np.uint64),
        "counts": np.array([o.Counts()[k] for k in ("rays", "boxes", "tris")], np.uint64),
    }
    o.close()
    return out


if __name__ == "__main__":
    for name in CASES:
        data = render_case(name)
        if data is None:
            print(name, "skipped: scene cache missing")
            continue
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
        print(name, {k: (v.shape, str(v.dtype)) for k, v in data.items()})
