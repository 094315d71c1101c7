"""Multi-GPU inside the product: N processes, one GPU and one TbHandle each, joined by tb_comm_init; the
only exchange step is the deterministic reduction of the accumulation buffers, over both transports: one kernel per rank
over the other GPUs' memory (CUDA IPC mappings, NVLink), and the NCCL all-gather + the library's own combine kernels. Needs >= 2 GPUs (skipped on a one-GPU box; run with `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _gpus():
    import torch
    return torch.cuda.device_count()


def _run_ranks(nranks, mode, scene, w, h, spp, bounces, workdir, transport="peer"):
    env = dict(os.environ, TB_COMM_TRANSPORT=transport)
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "nccl_worker.py"), str(r), str(nranks), mode, scene, str(w), str(h),
                               str(spp), str(bounces), str(workdir)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
             for r in range(nranks)]
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for r, p in enumerate(procs):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, outs[r])
    return [np.load(os.path.join(str(workdir), "rank%d.npz" % r)) for r in range(nranks)]


def _single(scene, w, h, spp, bounces):
    import tracerboy_b200 as tb
    g = tb.TracerBoy(0); g.LoadScene(scene); g.Resize(w, h)
    s = tb.get_default_output_settings(); s.MaxBounces = bounces
    g.Render(s, spp, 0.0)
    out = {k: g.Readback(k).copy() for k in (0, 1, 2)}
    g.Render(s, spp, 0.0)
    out["accum2"] = g.Readback(0).copy()
    out["rays"] = g.GetRenderStats().RaysTraced
    return out


def _check_after_resize(ranks):
    """The job-wide image rendered after a mid-job tb_resize (buffers unmapped, freed and mapped again): the fixed-order
    sum of the ranks' local buffers (row bands: disjoint, so the sum is their union), the same bits on every rank."""
    want = ranks[0]["small_local"].copy()
    for d in ranks[1:]:
        want = want + d["small_local"]
    assert want[..., 3].min() > 0
    for r, d in enumerate(ranks):
        assert np.array_equal(d["small"].view(np.uint32), want.view(np.uint32)), "rank %d after the resize" % r


@pytest.mark.parametrize("transport", ["peer", "nccl"])
@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_row_bands_over_nccl_are_bit_identical_to_one_gpu(nranks, transport, cornell, tmp_path):
    if _gpus() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    w, h, spp, bounces = 200, 150, 6, 4   # 19 bands of 8 rows (the last one partial): uneven over the ranks
    ranks = _run_ranks(nranks, "rows", cornell, w, h, spp, bounces, tmp_path, transport)
    _check_after_resize(ranks)
    one = _single(cornell, w, h, spp, bounces)
    for r, d in enumerate(ranks):
        rows = (np.arange(h) // 8) % nranks == r
        assert not d["local"][~rows].any(), "rank %d wrote outside its bands" % r
        assert np.array_equal(d["accum"].view(np.uint32), one[0].view(np.uint32)), "rank %d: reduced accumulation differs from one GPU" % r
        assert np.array_equal(d["jittered"].view(np.uint32), one[1].view(np.uint32))
        assert np.array_equal(d["rgb"].view(np.uint32), one[2].view(np.uint32))
        assert np.array_equal(d["accum2"].view(np.uint32), one["accum2"].view(np.uint32))
    assert sum(int(d["rays"][0]) for d in ranks) == one["rays"]


@pytest.mark.parametrize("transport", ["peer", "nccl"])
@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_sample_shards_over_nccl_sum_in_rank_order(nranks, transport, cornell, tmp_path):
    if _gpus() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    w, h, spp, bounces = 128, 96, 8 * nranks, 4
    ranks = _run_ranks(nranks, "samples", cornell, w, h, spp, bounces, tmp_path, transport)
    _check_after_resize(ranks)
    want = ranks[0]["local"].copy()
    for d in ranks[1:]:
        want = want + d["local"]          # float32, rank order: ((r0 + r1) + r2) + ...
    for r, d in enumerate(ranks):
        assert np.array_equal(d["accum"].view(np.uint32), want.view(np.uint32)), "rank %d: not the fixed-order sum" % r
        assert np.array_equal(d["accum"].view(np.uint32), ranks[0]["accum"].view(np.uint32))  # every rank holds the same bits
        assert np.array_equal(d["rgb"].view(np.uint32), (want[..., :3] / want[..., 3:4]).view(np.uint32))
    one = _single(cornell, w, h, spp, bounces)
    # the same samples as one GPU's first `spp` frames, in another summation order
    assert np.array_equal(want[..., 3], one[0][..., 3])
    assert np.allclose(want, one[0], rtol=2e-5, atol=1e-6)
    assert sum(int(d["rays"][0]) for d in ranks) == one["rays"]  # both sides rendered frames 0 .. 2 spp - 1
