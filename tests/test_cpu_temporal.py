"""Realtime temporal accumulation oracle (oracle/temporal.cpp, TemporalAccumulationCS.hlsl:100-235) on the CPU:
analytic known answers, and bit equality with the reference's own shader text compiled from the mount
(oracle/_ref/libref_temporal.so, resources shimmed)."""
import numpy as np
import pytest


def make_inputs(seed=0, w=96, h=64, shift=(0.0, 0.0, 0.0)):
    """A camera looking down -z at a tilted plane; world positions / normals of the visible surface computed
    analytically for the current camera and for a previous camera displaced by `shift`."""
    import tracerboy_b200 as tb
    rng = np.random.default_rng(seed)

    def camera(pos):
        c = tb.Camera()
        c.Position.x, c.Position.y, c.Position.z = pos
        c.LookAt.x, c.LookAt.y, c.LookAt.z = pos[0], pos[1], pos[2] - 1.0
        c.Right.x, c.Right.y, c.Right.z = 1.0, 0.0, 0.0
        c.Up.x, c.Up.y, c.Up.z = 0.0, 1.0, 0.0
        c.LensHeight = 2.0
        c.FocalDistance = 3.0
        return c

    def world_positions(pos):
        # pixel (x, y) -> lens point -> ray from the focal point (behind the lens) -> plane z = -10 - 0.3 x
        aspect = w / h
        u = (np.arange(w) + 0.5) / w
        v = 1.0 - (np.arange(h) + 0.5) / h
        lx = pos[0] + (u * 2 - 1) * (2.0 * aspect) / 2
        ly = pos[1] + (v * 2 - 1) * 2.0 / 2
        L = np.stack(np.broadcast_arrays(lx[None, :], ly[:, None], np.full((h, w), pos[2])), -1)
        F = np.array([pos[0], pos[1], pos[2] + 3.0])
        d = L - F
        d /= np.linalg.norm(d, axis=-1, keepdims=True)
        # plane: z + 0.3 x + 10 = 0
        nrm = np.array([0.3, 0.0, 1.0]) / np.linalg.norm([0.3, 0.0, 1.0])
        t = -(F @ np.array([0.3, 0.0, 1.0]) + 10.0) / (d @ np.array([0.3, 0.0, 1.0]))
        P = F + d * t[..., None]
        wp = np.concatenate([P, np.ones((h, w, 1))], -1).astype(np.float32)
        nn = np.concatenate([np.broadcast_to(nrm, (h, w, 3)), np.ones((h, w, 1))], -1).astype(np.float32)
        return wp, nn

    cur_cam, prev_cam = camera((0.0, 0.0, 0.0)), camera(shift)
    wp, nn = world_positions((0.0, 0.0, 0.0))
    pwp, _ = world_positions(shift)
    p = tb.TemporalAccumulationParams()
    p.Camera, p.PrevCamera = cur_cam, prev_cam
    p.HistoryWeight, p.IgnoreHistory, p.OutputMomentInformation = 0.95, 0, 1
    history = rng.random((h, w, 4)).astype(np.float32)
    current = rng.random((h, w, 4)).astype(np.float32)
    moments = rng.random((h, w, 4)).astype(np.float32)
    moments[..., 2] = rng.integers(0, 50, (h, w))
    return p, history, current, wp, pwp, nn, moments


def test_static_camera_blends_history_per_pixel(built):
    """Previous camera == current camera: every pixel reprojects onto itself (bilinear weights (1,0)), so
    out = lerp(current, history, 0.95) and the moments follow their recurrence."""
    from oracle import binding
    p, history, current, wp, pwp, nn, moments = make_inputs()
    out, mom = binding.temporal_accumulate_image(p, history, current, wp, pwp, nn, moments)
    inner = (slice(2, -2), slice(2, -2))
    want = current[..., :3] + 0.95 * (history[..., :3] - current[..., :3])
    assert np.allclose(out[inner][..., :3], want[inner], atol=2e-3)  # reprojection is exact up to rounding of the uv
    lum = current[..., :3] @ np.array([0.212671, 0.715160, 0.072169], np.float32)
    cnt = moments[..., 2] + 1
    assert np.allclose(mom[inner][..., 2], cnt[inner], atol=2e-3)  # the bilinear fetch carries ~1e-6 of the neighbours
    f = 1 / np.minimum(cnt, 32)
    m1 = moments[..., 0] + f * (lum - moments[..., 0])
    assert np.allclose(mom[inner][..., 0], m1[inner], atol=2e-3)
    assert np.allclose(out[inner][..., 3], np.maximum(mom[..., 1] - mom[..., 0] ** 2, 0)[inner], atol=1e-6)


def test_ignore_history_and_invalid_hits_pass_the_current_frame_through(built):
    from oracle import binding
    p, history, current, wp, pwp, nn, moments = make_inputs(1)
    p.IgnoreHistory = 1
    p.OutputMomentInformation = 0
    out, mom = binding.temporal_accumulate_image(p, history, current, wp, pwp, nn)
    assert mom is None and np.array_equal(out[..., :3], current[..., :3]) and (out[..., 3] == 1).all()
    p.IgnoreHistory = 0
    nn[:10] = 0  # no hit (zero normal): history is not used there
    out, _ = binding.temporal_accumulate_image(p, history, current, wp, pwp, nn)
    assert np.array_equal(out[:10, :, :3], current[:10, :, :3])
    assert not np.array_equal(out[12:, :, :3], current[12:, :, :3])


def test_moved_camera_rejects_disoccluded_history(built):
    """The previous frame saw a different surface (world positions far away): the world-position test rejects all four
    taps and the pixel falls back to the current frame."""
    from oracle import binding
    p, history, current, wp, pwp, nn, moments = make_inputs(2, shift=(0.4, 0.1, 0.0))
    p.OutputMomentInformation = 0
    out, _ = binding.temporal_accumulate_image(p, history, current, wp, pwp, nn)
    blended = (out[..., :3] != current[..., :3]).any(-1)
    assert 0.5 < blended.mean() < 1.0  # most pixels find valid history, the band that left the previous frame does not
    far = pwp.copy(); far[..., :3] += 1000.0
    out, _ = binding.temporal_accumulate_image(p, history, current, wp, far, nn)
    assert np.array_equal(out[..., :3], current[..., :3])


def test_restatement_equals_reference_shader_text(built):
    """oracle/temporal.cpp against the reference's own TemporalAccumulationCS.hlsl main() compiled from the mount
    (oracle/_ref/libref_temporal.so; only the resources are shims there): colour, alpha (variance) and moments
    bit for bit, for a static camera, shifted previous cameras (reprojection, history rejection, off-screen
    history), rotated previous frames, IgnoreHistory, moments on and off, invalid hits (zero normals) and random
    (non-planar) world positions that exercise the neighbourhood rejection test."""
    import ctypes as C
    import os
    from oracle import binding
    path = os.path.join(os.path.dirname(binding.ref_traverse_lib_path()), "libref_temporal.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libref_temporal.so not built (needs the reference mount at build time)")
    ref = C.CDLL(path)
    ref.ref_temporal_accumulate_image.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32] + [C.c_void_p] * 8
    ref.ref_temporal_accumulate_image.restype = C.c_int
    rng = np.random.default_rng(12)

    def run_ref(p, history, current, wp, pwp, nn, moments):
        h, w = current.shape[:2]
        color = np.zeros((h, w, 4), np.float32)
        mom = np.zeros((h, w, 4), np.float32)
        imgs = [np.ascontiguousarray(a, np.float32) for a in (history, current, wp, pwp, nn, moments)]
        rc = ref.ref_temporal_accumulate_image(C.byref(p), w, h, *[a.ctypes.data_as(C.c_void_p) for a in imgs],
                                               color.ctypes.data_as(C.c_void_p), mom.ctypes.data_as(C.c_void_p))
        assert rc == 0
        return color, mom

    cases = 0
    for seed, shift in ((0, (0.0, 0.0, 0.0)), (1, (0.31, 0.0, 0.0)), (2, (-0.2, 0.13, 0.05)), (3, (4.0, 0.0, 0.0)), (4, (0.0, -0.4, 0.6))):
        for variant in ("plain", "no_moments", "ignore_history", "rotated", "random_positions", "holes"):
            p, history, current, wp, pwp, nn, moments = make_inputs(seed, 80, 48, shift)
            if variant == "no_moments": p.OutputMomentInformation = 0
            if variant == "ignore_history": p.IgnoreHistory = 1
            if variant == "rotated":
                a = 0.07
                p.PrevCamera.LookAt.x = p.PrevCamera.Position.x + np.sin(a)
                p.PrevCamera.LookAt.z = p.PrevCamera.Position.z - np.cos(a)
                p.PrevCamera.Right.x, p.PrevCamera.Right.z = np.cos(a), np.sin(a)
            if variant == "random_positions":
                wp[..., :3] += rng.normal(0, 0.4, wp[..., :3].shape).astype(np.float32)
                pwp[..., :3] += rng.normal(0, 0.4, pwp[..., :3].shape).astype(np.float32)
            if variant == "holes":
                nn[rng.random(nn.shape[:2]) < 0.2] = 0.0
                history[rng.random(nn.shape[:2]) < 0.05] = np.nan
            c0, m0 = binding.temporal_accumulate_image(p, history, current, wp, pwp, nn, moments)
            c1, m1 = run_ref(p, history, current, wp, pwp, nn, moments)
            same = (c0.view(np.uint32) == c1.view(np.uint32)) | (np.isnan(c0) & np.isnan(c1))
            assert same.all(), (seed, variant, int((~same).sum()))
            if p.OutputMomentInformation:
                same = (m0.view(np.uint32) == m1.view(np.uint32)) | (np.isnan(m0) & np.isnan(m1))
                assert same.all(), (seed, variant, "moments", int((~same).sum()))
            cases += 1
    assert cases == 30
