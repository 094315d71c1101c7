"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs. Integer / index work must be bit-exact; so is radiance here, because both
sides evaluate the pinned intrinsics (tb_math.h) in the same order. The looser tolerance the
north star states (per-pixel relMSE <= 1e-3) is asserted as well so a future relaxation of
bit-exactness still has a stated bound."""
import os

import numpy as np
import pytest

from conftest import scene_path, ROOT

pytestmark = pytest.mark.gpu

REL_MSE_TOL = 1e-3  # BASELINE.json north_star: per-pixel relative MSE at matched seed and bounce count


def _pair(path, w, h, passes=3):
    import tracerboy_b200 as tb
    from oracle.binding import Oracle
    g = tb.TracerBoy(0)
    g.LoadScene(path, {3: tb.api.BVH_BUILD_PREFER_FAST_TRACE, 0: tb.api.BVH_BUILD_PREFER_FAST_BUILD, 1: 0}[passes])
    g.Resize(w, h)
    o = Oracle()
    o.LoadScene(path, passes)
    o.Resize(w, h)
    return g, o


def _tbscene(spec, tmp_path):
    import tracerboy_b200 as tb
    if spec.endswith(".tbscene"):
        return spec
    out = str(tmp_path / "scene.tbscene")
    tb.convert_scene(spec, out)
    return out


def _rel_mse(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    return ((a - b) ** 2 / (b ** 2 + 1e-4)).mean(axis=-1)


def _compare_render(g, o, settings, spp, kinds=(0, 1, 3, 4, 5, 6, 7, 8, 9)):
    import tracerboy_b200 as tb
    g.Render(settings, spp, 0.0)
    o.Render(settings, spp, 0.0)
    for k in kinds:
        a, b = g.Readback(k), o.Readback(k)
        if a.dtype == np.float32:
            same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
        else:
            same = a == b
        bad = np.argwhere(~same)
        assert bad.shape[0] == 0, "buffer kind %d differs at %d elements, first %s gpu=%s oracle=%s" % (
            k, bad.shape[0], bad[0], a[tuple(bad[0])], b[tuple(bad[0])])
    acc_g, acc_o = g.Readback(tb.BufferKind.ACCUM_RGBW), o.Readback(tb.BufferKind.ACCUM_RGBW)
    assert _rel_mse(acc_g[..., :3], acc_o[..., :3]).max() <= REL_MSE_TOL
    rs = g.GetRenderStats(); oc = o.Counts()
    assert (rs.RaysTraced, rs.BoxesTested, rs.TrianglesTested) == (oc["rays"], oc["boxes"], oc["tris"])


@pytest.mark.parametrize("spec,passes", [
    ("cornell", 3), ("cornell", 0), ("teapot", 3),
    ("synthetic:blobs?copies=27&tris=300&seed=3", 3), ("synthetic:blobs?copies=1&tris=20000&seed=5", 3),
    ("synthetic:blobs?copies=8&tris=1000&seed=7", 1),
])
def test_bvh_bytes_identical(spec, passes, tmp_path, built):
    """BVH build determinism: the GPU builder's output equals the oracle's byte for byte."""
    path = scene_path("cornell-box") if spec == "cornell" else scene_path("teapot") if spec == "teapot" else _tbscene(spec, tmp_path)
    if path is None:
        pytest.skip("scene cache missing")
    g, o = _pair(path, 8, 8, passes)
    a, b = g.GetBVH(), o.GetBVH()
    assert a.shape == b.shape
    diff = np.flatnonzero(a != b)
    assert diff.size == 0, "BVH differs at %d bytes, first at %d" % (diff.size, diff[0])
    g2, _ = _pair(path, 8, 8, passes)  # run-to-run determinism
    assert np.array_equal(g2.GetBVH(), a)


def _random_rays(n, cam, seed):
    from tracerboy_b200.api import RAY_DTYPE
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, RAY_DTYPE)
    eye = np.array(cam.Position.tuple(), np.float32)
    rays["Origin"] = eye + rng.normal(0, 0.05, (n, 3)).astype(np.float32)
    d = np.array(cam.LookAt.tuple(), np.float32) - eye + rng.normal(0, 0.25, (n, 3)).astype(np.float32)
    rays["Direction"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays["TMin"] = 0.001
    rays["TMax"] = 999999.0
    # a few incoherent rays from inside the scene
    k = n // 4
    rays["Direction"][:k] = rng.normal(0, 1, (k, 3)).astype(np.float32)
    return rays


@pytest.mark.parametrize("spec", ["cornell", "teapot", "synthetic:blobs?copies=27&tris=300&seed=3"])
def test_trace_rays_bit_exact(spec, tmp_path, built):
    """tb_trace_rays == SoftwareRayQuery restatement: t, barycentrics, ids and both counters."""
    path = scene_path("cornell-box") if spec == "cornell" else scene_path("teapot") if spec == "teapot" else _tbscene(spec, tmp_path)
    if path is None:
        pytest.skip("scene cache missing")
    g, o = _pair(path, 8, 8)
    rays = _random_rays(200000, g.GetCamera(), 11)
    hg, ho = g.TraceRays(rays), o.TraceRays(rays)
    for f in hg.dtype.names:
        a, b = hg[f], ho[f]
        same = (a.view(np.uint32) == b.view(np.uint32))
        assert same.all(), "field %s differs for %d rays" % (f, (~same).sum())
    assert (hg["t"] > 0).sum() > 1000


def test_cornell_render_bit_exact(cornell):
    import tracerboy_b200 as tb
    g, o = _pair(cornell, 128, 128)
    s = tb.get_default_output_settings()
    s.MaxBounces = 4
    _compare_render(g, o, s, 4)


def test_teapot_render_bit_exact(teapot):
    import tracerboy_b200 as tb
    g, o = _pair(teapot, 256, 144)
    s = tb.get_default_output_settings()
    _compare_render(g, o, s, 2)


@pytest.mark.parametrize("variant", ["no_blue_noise", "no_nee", "sir", "dof", "triangle", "gaussian", "firefly", "heatmap"])
def test_settings_variants(variant, cornell):
    import tracerboy_b200 as tb
    g, o = _pair(cornell, 96, 96)
    s = tb.get_default_output_settings()
    s.MaxBounces = 5
    if variant == "no_blue_noise": s.EnableBlueNoise = 0
    if variant == "no_nee": s.EnableNextEventEstimation = 0
    if variant == "sir": s.EnableSamplingImportanceResampling = 1
    if variant == "dof": s.DOFFocalDistance = 5.0
    if variant == "triangle": s.FilterType = 1; s.FilterWidth = 2.0
    if variant == "gaussian": s.FilterType = 2; s.FilterWidth = 2.0
    if variant == "firefly": s.FireflyClampValue = 2.0
    if variant == "heatmap": s.OutputType = 9
    _compare_render(g, o, s, 3)


def test_materials_scene_bit_exact(tmp_path, built):
    """Glass (SSS walk), metal, substrate, matte, area light + constant sky."""
    import tracerboy_b200 as tb
    path = _tbscene("synthetic:blobs?copies=27&tris=300&seed=3", tmp_path)
    g, o = _pair(path, 160, 90)
    s = tb.get_default_output_settings()
    s.MaxBounces = 8
    _compare_render(g, o, s, 3)
