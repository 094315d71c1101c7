"""Post-process oracle (oracle/postprocess.cpp) on the CPU: pinned against the reference's own Tonemap.h /
PostProcessCS.hlsl text compiled from the mount (oracle/_ref/libref_post.so), against independent float64 numpy
restatements of the published tone curves, and against hand-computed histogram known answers."""
import itertools

import numpy as np
import pytest


def _inputs(seed=1, h=48, w=80):
    """HDR accumulation-like image: rgb * w with w = frame count, plus the edge cases the shader meets: black pixels,
    zero weight (0/0), negative, huge and NaN radiance."""
    rng = np.random.default_rng(seed)
    img = np.exp(rng.normal(0, 3, (h, w, 4))).astype(np.float32)
    img[..., 3] = rng.integers(1, 64, (h, w))
    img[..., :3] *= img[..., 3:]
    img[0, :8, :3] = 0
    img[1, :4, 3] = 0
    img[2, :4, 0] = -1.0
    img[3, :4, :3] = 1e30
    img[4, 0, 0] = np.nan
    aux = (rng.random((h, w, 4)) * 0.2).astype(np.float32)
    return img, aux


def _same(a, b):
    return ((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all()


def test_restated_postprocess_equals_reference_text(built):
    """Every OutputType x every TonemapType (+ an unknown one = the default branch) x auto exposure x gamma."""
    import tracerboy_b200 as tb
    from oracle import binding
    if not binding.reference_post_available():
        pytest.skip("oracle/_ref/libref_post.so not built (needs the reference mount at build time)")
    img, aux = _inputs()
    for ot, tm, auto, gam in itertools.product(range(10), range(9), (0, 1), (0, 1)):
        s = tb.PostProcessSettings(1.7, tm, gam, auto, 2.5)
        o, _, _, avg = binding.postprocess_image(img, ot, s, aux)
        r = binding.reference_postprocess_image(img, ot, s, avg, aux)
        assert _same(o, r), "output type %d tonemap %d auto %d gamma %d" % (ot, tm, auto, gam)


def test_tone_curves_against_float64_numpy(built):
    """Independent restatement of the published curves in float64: Reinhard x/(1+x), the Narkowicz/Hill ACES fit,
    Uncharted2 (Hable), clamp; each followed by the 1/2.2 gamma. Tolerance: float32 evaluation error."""
    import tracerboy_b200 as tb
    from oracle import binding
    rng = np.random.default_rng(3)
    x = np.exp(rng.normal(-1, 2, (32, 32, 3)))
    img = np.concatenate([x, np.ones((32, 32, 1))], -1).astype(np.float32)
    x = img[..., :3].astype(np.float64)

    def run(tm):
        s = tb.PostProcessSettings(1.0, tm, 1, 0, 1.0)
        return binding.postprocess_image(img, 0, s)[0][..., :3].astype(np.float64)
    g = 1 / 2.2
    assert np.allclose(run(tb.TonemapType.REINHARD), (x / (1 + x)) ** g, rtol=2e-5)
    assert np.allclose(run(tb.TonemapType.CLAMP), np.clip(x, 0, 1) ** g, rtol=2e-5)
    A, B, C, D, E, F = 0.15, 0.50, 0.10, 0.20, 0.02, 0.30
    def part(v): return ((v * (A * v + C * B) + D * E) / (v * (A * v + B) + D * F)) - E / F
    assert np.allclose(run(tb.TonemapType.UNCHARTED), (part(2 * x) / part(11.2)) ** g, rtol=1e-4)
    mi = np.array([[0.59719, 0.35458, 0.04823], [0.07600, 0.90834, 0.01566], [0.02840, 0.13383, 0.83777]])
    mo = np.array([[1.60475, -0.53108, -0.07367], [-0.10208, 1.10813, -0.00605], [-0.00327, -0.07276, 1.07602]])
    v = x @ mi.T
    v = (v * (v + 0.0245786) - 0.000090537) / (v * (0.983729 * v + 0.4329510) + 0.238081)
    assert np.allclose(run(tb.TonemapType.ACES), np.clip(v @ mo.T, 0, 1) ** g, rtol=1e-4, atol=1e-5)


def test_histogram_and_average_known_answers(built):
    """GenerateHistogramCS / CalculateAveragedLuminanceCS: bin 0 for luminance < 1e-5, bin = uint(sat((log2 L + 10)/16)
    * 254 + 1), average = exp2(((sum(bin*count) // (N - count[0])) - 1)/254 * 16 - 10) with the *integer* division.
    And the shader's edge behaviour: thread Gid of a 16x16 group adds bin Gid to the global histogram after the
    threads outside the image have returned, so a partial group never adds the bins of its missing threads."""
    import tracerboy_b200 as tb
    from oracle import binding
    lum = np.array([0.0, 1e-6, 2.0 ** -10, 2.0 ** -2, 1.0, 2.0 ** 6, 1e9, 0.5], np.float32)
    want_bins = [0, 0, 1, int((8 / 16) * 254 + 1), int((10 / 16) * 254 + 1), 255, 255, int((9 / 16) * 254 + 1)]
    s = tb.PostProcessSettings(1.0, tb.TonemapType.CLAMP, 1, 1, 1.0)
    # a full 16x16 group: every bin has its thread. 8 probe pixels, the other 248 black (bin 0)
    img = np.zeros((16, 16, 4), np.float32)
    img[..., 3] = 3.0
    img[0, :8, :3] = lum[:, None] * 3.0  # rgb / w with w = 3; luma weights sum to 1
    _, _, hist, avg = binding.postprocess_image(img, 0, s)
    got = np.repeat(np.arange(256), hist)
    assert sorted(got.tolist()) == sorted(want_bins + [0] * 248)
    q = sum(want_bins) // (256 - 250)
    assert np.isclose(avg, 2.0 ** ((q - 1) / 254 * 16 - 10), rtol=1e-5)
    # the same row alone, 1x8: the only group has threads 0..7, so only bins 0..7 are ever added
    _, _, hist, _ = binding.postprocess_image(img[:1, :8], 0, s)
    assert hist.sum() == 3 and hist[0] == 2 and hist[1] == 1
    # 1080p-like: 24 rows = one full row of groups + 8 rows; bright pixels (bin 159 >= 128) count only in the full groups
    tall = np.ones((24, 16, 4), np.float32)
    _, _, hist, _ = binding.postprocess_image(tall, 0, s)
    assert hist[int((10 / 16) * 254 + 1)] == 256 and hist.sum() == 256
    # all-black image: N - count[0] == 0, D3D unsigned division by zero = 0xffffffff -> exp2(huge) = inf, exposure 0
    _, _, hist, avg = binding.postprocess_image(np.zeros((2, 2, 4), np.float32) + np.array([0, 0, 0, 1], np.float32), 0, s)
    assert hist[0] == 4 and np.isinf(avg)


def test_auto_exposure_centres_the_image(built):
    """A constant image of any brightness comes out at (about) linear mid-gray before the tonemap."""
    import tracerboy_b200 as tb
    from oracle import binding
    s = tb.PostProcessSettings(1.0, tb.TonemapType.CLAMP, 1, 1, 1.0)
    for level in (0.01, 1.0, 37.0):
        img = np.full((32, 32, 4), level, np.float32)  # full 16x16 groups: partial groups drop bins (see the known answers)
        img[..., 3] = 1.0
        out = binding.postprocess_image(img, 0, s)[0]
        # gamma(0.5^2.2 * L / avg): the averaged luminance is quantised to 254 bins over 16 stops (about 4.5 % per bin)
        assert abs(out[0, 0, 0] - 0.5) < 0.03


def test_unorm8_store(built):
    import tracerboy_b200 as tb
    from oracle import binding
    img = np.zeros((1, 6, 4), np.float32)
    img[0, :, 0] = [0.0, 0.5 / 255.0, 1.0, 7.0, -3.0, np.nan]
    s = tb.PostProcessSettings(1.0, 0, 0, 0, 1.0)
    out, out8, _, _ = binding.postprocess_image(img, tb.OutputType.LUMINANCE_VARIANCE, s)
    assert out8[0, :, 0].tolist() == [0, 1, 255, 255, 0, 0] and (out8[..., 3] == 255).all()


def test_restated_histogram_equals_reference_shader_text(built):
    """The auto-exposure pair against the reference's own GenerateHistogramCS.hlsl and CalculateAveragedLuminanceCS.hlsl
    compiled from the mount (oracle/_ref/libref_hist.so: a 16x16 group is 256 host threads with a real group barrier,
    InterlockedAdd is an atomic): all 256 bins and the averaged luminance bit for bit, on HDR images with black,
    zero-weight, negative, huge and NaN pixels, sizes that are not multiples of the 16x16 group (threads outside the
    image return before the second barrier), an all-black image (unsigned division by zero = 0xffffffff in D3D) and a
    one-pixel image."""
    import ctypes as C
    import os
    import tracerboy_b200 as tb
    from oracle import binding
    path = os.path.join(os.path.dirname(binding.ref_traverse_lib_path()), "libref_hist.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libref_hist.so not built (needs the reference mount at build time)")
    ref = C.CDLL(path)
    ref.ref_luminance_histogram.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    ref.ref_averaged_luminance.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    s = tb.PostProcessSettings(1.0, 0, 1, 1, 1.0)   # auto exposure on: the histogram pass runs (TracerBoy.cpp:2948)
    cases = []
    for seed, (h, w) in enumerate([(48, 80), (33, 47), (16, 16), (1, 1), (17, 5)]):
        img, _ = _inputs(seed + 3, h, w) if h > 4 else (np.full((h, w, 4), 2.0, np.float32), None)
        cases.append(img)
    cases.append(np.zeros((20, 20, 4), np.float32) + np.array([0, 0, 0, 1], np.float32))   # all black
    dim = np.full((24, 40, 4), 1e-7, np.float32); dim[..., 3] = 1                           # everything below epsilon
    cases.append(dim)
    for img in cases:
        h, w = img.shape[:2]
        _, _, hist_o, avg_o = binding.postprocess_image(img, 0, s)
        hist_r = np.zeros(256, np.uint32)
        img_c = np.ascontiguousarray(img, np.float32)
        assert ref.ref_luminance_histogram(img_c.ctypes.data_as(C.c_void_p), w, h, hist_r.ctypes.data_as(C.c_void_p)) == 0
        assert int(hist_r.sum()) <= h * w   # partial edge groups drop the bins of their missing threads
        assert np.array_equal(hist_o, hist_r), (h, w, np.flatnonzero(hist_o != hist_r)[:8])
        avg_r = np.zeros(1, np.float32)
        assert ref.ref_averaged_luminance(hist_r.ctypes.data_as(C.c_void_p), h * w, avg_r.ctypes.data_as(C.c_void_p)) == 0
        a = np.array([avg_o], np.float32)
        assert a.view(np.uint32)[0] == avg_r.view(np.uint32)[0] or (np.isnan(a[0]) and np.isnan(avg_r[0])), (h, w, avg_o, avg_r[0])
