"""Mint tests/golden/cornell_64_4096spp_refcore.npz: the converged reference image of the north star's
"converged-image RMSE against a 4096-spp reference", rendered by the REFERENCE'S OWN CORE -- TracerBoy/kernel.glsl
compiled from the mount as host C++ (oracle/_ref/libref_core.so, built by oracle/build_ref.py) inside the oracle's
per-pixel loop -- not by the CUDA library. Needs /root/reference at build time; the fixture travels.

    python tests/golden/make_converged.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import tracerboy_b200 as tb
    from oracle import binding
    from oracle.binding import Oracle
    assert binding.reference_core_available(), "oracle/_ref/libref_core.so is missing (needs the reference mount)"
    before = binding.use_reference_core(True)
    o = Oracle()
    o.LoadScene(os.path.join(ROOT, "tests", "golden", "cornell-box.tbscene"), 3)
    o.Resize(64, 64)
    s = tb.get_default_output_settings()
    s.MaxBounces = 4
    o.Render(s, 4096, 0.0, threads=os.cpu_count())
    traced = binding.use_reference_core(False) - before
    assert traced == 64 * 64 * 4096, traced  # every pixel sample went through the reference text
    acc = o.Readback(tb.BufferKind.ACCUM_RGBW)
    rgb = (acc[..., :3] / acc[..., 3:4]).astype(np.float32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cornell_64_4096spp_refcore.npz"), resolved_rgb=rgb,
                        spp=np.array([4096]), reference_core=np.array([1]), bounces=np.array([4]))
    print("wrote fixture, mean radiance", rgb.mean())


if __name__ == "__main__":
    main()
