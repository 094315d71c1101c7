"""GPU parity of realtime temporal accumulation (tb_temporal_accumulate_image, k_temporal_accumulate) against the CPU
oracle (oracle/temporal.cpp): bit-exact colour, variance alpha and moments."""
import numpy as np
import pytest

from test_cpu_temporal import make_inputs

pytestmark = pytest.mark.gpu


def _same(a, b):
    return ((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all()


@pytest.mark.parametrize("shift,moments_on,ignore", [((0.0, 0.0, 0.0), 1, 0), ((0.4, 0.1, 0.0), 1, 0), ((0.0, -0.3, 0.5), 0, 0),
                                                     ((0.2, 0.0, 0.0), 1, 1), ((5.0, 0.0, 0.0), 1, 0)])
def test_temporal_accumulate_image_bit_exact(shift, moments_on, ignore, built):
    import tracerboy_b200 as tb
    from oracle import binding
    g = tb.TracerBoy(0)
    p, history, current, wp, pwp, nn, moments = make_inputs(5, w=200, h=120, shift=shift)
    p.OutputMomentInformation, p.IgnoreHistory = moments_on, ignore
    nn[:7, :30] = 0                      # misses
    wp[50, 60, :3] = np.nan              # a NaN world position must not poison its neighbours' history test differently
    mh = moments if moments_on else None
    want, wmom = binding.temporal_accumulate_image(p, history, current, wp, pwp, nn, mh)
    got, gmom = g.TemporalAccumulateImage(p, history, current, wp, pwp, nn, mh)
    assert _same(want, got)
    if moments_on:
        assert _same(wmom, gmom)


def test_temporal_accumulation_of_rendered_frames(cornell):
    """Realtime mode end to end: two one-sample frames of the cornell box with a camera step between them, world
    positions and normals from the tracer's own AOVs; the GPU pass equals the oracle's and converges towards the
    history where the surface is unchanged."""
    import tracerboy_b200 as tb
    from oracle import binding
    g = tb.TracerBoy(0)
    g.LoadScene(cornell)
    g.Resize(192, 192)
    s = tb.get_default_output_settings()
    s.RenderMode = 1  # RealTime: every frame overwrites the accumulation buffer
    K = tb.BufferKind
    cam0 = g.GetCamera()
    g.Render(s, 1, 0.0)
    hist, pwp = g.Readback(K.ACCUM_RGBW).copy(), g.Readback(K.AOV_WORLDPOS).copy()
    cam1 = g.GetCamera()
    cam1.Position.x += 0.02; cam1.LookAt.x += 0.02
    g.SetCamera(cam1)
    g.Render(s, 1, 0.0)
    cur, wp, nn = g.Readback(K.ACCUM_RGBW).copy(), g.Readback(K.AOV_WORLDPOS).copy(), g.Readback(K.AOV_NORMAL).copy()
    p = tb.TemporalAccumulationParams()
    p.Camera, p.PrevCamera = cam1, cam0
    p.HistoryWeight, p.IgnoreHistory, p.OutputMomentInformation = 0.95, 0, 1
    mom0 = np.zeros_like(cur)
    want, wmom = binding.temporal_accumulate_image(p, hist, cur, wp, pwp, nn, mom0)
    got, gmom = g.TemporalAccumulateImage(p, hist, cur, wp, pwp, nn, mom0)
    assert _same(want, got) and _same(wmom, gmom)
    used = (got[..., :3] != cur[..., :3]).any(-1)
    assert used.mean() > 0.5  # most of the box is seen by both frames
