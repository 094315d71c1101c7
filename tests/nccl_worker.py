"""One rank of the multi-GPU test (tests/test_gpu_multi.py): its own process, its own GPU, the product's C ABI only.
The NCCL id travels through a file written by rank 0 (the host's job; any transport works)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, nranks, mode, scene, w, h, spp, bounces, workdir = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], \
        int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7]), int(sys.argv[8]), sys.argv[9]
    import tracerboy_b200 as tb
    idfile = os.path.join(workdir, "nccl_id.bin")
    if rank == 0:
        uid = tb.comm_get_unique_id()
        with open(idfile + ".tmp", "wb") as f:
            f.write(uid)
        os.rename(idfile + ".tmp", idfile)
    else:
        t0 = time.time()
        while not os.path.exists(idfile):
            if time.time() - t0 > 120:
                raise SystemExit("rank %d: no NCCL id after 120 s" % rank)
            time.sleep(0.05)
        uid = open(idfile, "rb").read()
    g = tb.TracerBoy(rank)  # device ordinal = rank
    g.LoadScene(scene)
    g.Resize(w, h)
    g.CommInit(uid, rank, nranks, tb.SHARD_ROWS if mode == "rows" else tb.SHARD_SAMPLES)
    s = tb.get_default_output_settings()
    s.MaxBounces = bounces
    per_rank = spp if mode == "rows" else spp // nranks
    g.Render(s, per_rank, 0.0)
    local = g.Readback(tb.BufferKind.LOCAL_ACCUM_RGBW).copy()
    accum = g.Readback(tb.BufferKind.ACCUM_RGBW).copy()       # collective: runs the reduction
    jit = g.Readback(tb.BufferKind.JITTERED_RGBW).copy()      # served from the same reduction
    rgb = g.Readback(tb.BufferKind.RESOLVED_RGB).copy()
    info = g.CommInfo()
    assert info.Reductions == 1 and info.NumRanks == nranks and info.Rank == rank
    want_peer = os.environ.get("TB_COMM_TRANSPORT", "peer") != "nccl" and nranks > 1
    assert info.Transport == (tb.api.COMM_TRANSPORT_PEER if want_peer else tb.api.COMM_TRANSPORT_NCCL), info.Transport
    # progressive: more frames, then the image again
    g.Render(s, per_rank, 0.0)
    accum2 = g.Readback(tb.BufferKind.ACCUM_RGBW).copy()
    info = g.CommInfo()
    assert info.Reductions == 2
    rays = g.GetRenderStats().RaysTraced
    # a resize in the middle of the job (collective while peers are mapped): buffers are unmapped, freed, mapped again
    g.Resize(w // 2, h // 2)
    g.Render(s, per_rank, 0.0)
    small = g.Readback(tb.BufferKind.ACCUM_RGBW).copy()
    small_local = g.Readback(tb.BufferKind.LOCAL_ACCUM_RGBW).copy()
    np.savez(os.path.join(workdir, "rank%d.npz" % rank), local=local, accum=accum, jittered=jit, rgb=rgb, accum2=accum2, small=small, small_local=small_local,
             rays=np.array([rays], np.uint64),
             comm=np.array([info.NcclVersion, info.BytesReceivedPerReduction, int(info.LastReductionMilliseconds * 1e3)], np.uint64))
    g.CommDestroy()
    g.close()


if __name__ == "__main__":
    main()
