"""SIMD-utilisation model of the persistent traversal kernel (CPU, uses the oracle's visit sequences).

For a sample of primary and diffuse-bounce rays of a scene, replays the warp scheduling of k_extend — per-lane refill
below a threshold, one code path (internal / leaf) per iteration chosen by majority vote — and variants of it, and
reports useful lane-steps per issued warp-step. Costs: internal step 45 instructions, leaf step 75, service phase
(retire + refill) 150 per warp pass.

    python tools/simd_model.py [scene.tbscene] [rays]
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import binding  # noqa: E402
from tracerboy_b200.api import RAY_DTYPE, HIT_DTYPE  # noqa: E402

C_INT, C_LEAF, C_SERVICE = 45, 75, 150


def visit_sequences(o, rays):
    lib = binding.load()
    n = len(rays)
    hits = np.zeros(n, HIT_DTYPE)
    cap = 4096 * n
    seq = np.zeros(cap, np.uint8)
    off = np.zeros(n + 1, np.uint64)
    lib.oracle_trace_rays_visits.argtypes = [C.c_void_p] * 2 + [C.c_uint64] + [C.c_void_p] * 2 + [C.c_uint64, C.c_void_p]
    rc = lib.oracle_trace_rays_visits(o.h, rays.ctypes.data, n, hits.ctypes.data, seq.ctypes.data, cap, off.ctypes.data)
    assert rc == 0
    return [seq[int(off[i]):int(off[i + 1])] for i in range(n)], hits


def simulate(seqs, rays_per_lane=1, refill_below=14):
    """returns (useful instruction-lanes, issued instruction-lanes)"""
    nxt = 0
    slots = [[None, 0] for _ in range(32 * rays_per_lane)]  # [sequence, position]
    useful = issued = 0

    def refill():
        nonlocal nxt
        got = False
        for s in slots:
            if (s[0] is None or s[1] >= len(s[0])) and nxt < len(seqs):
                s[0], s[1] = seqs[nxt], 0
                nxt += 1
                got = True
        return got

    refill()
    issued += C_SERVICE * 32
    while True:
        busy_lanes = 0
        want = [0, 0]
        lane_choice = []
        for lane in range(32):
            types = set()
            for r in range(rays_per_lane):
                s = slots[lane * rays_per_lane + r]
                if s[0] is not None and s[1] < len(s[0]):
                    types.add(int(s[0][s[1]]))
            lane_choice.append(types)
            if types:
                busy_lanes += 1
                for t in types:
                    want[t] += 1
        if busy_lanes == 0:
            if not refill():
                break
            issued += C_SERVICE * 32
            continue
        if busy_lanes < refill_below and nxt < len(seqs):
            refill()
            issued += C_SERVICE * 32
            continue
        t = 1 if want[1] > want[0] else 0
        cost = C_LEAF if t else C_INT
        issued += cost * 32
        for lane in range(32):
            if t in lane_choice[lane]:
                for r in range(rays_per_lane):
                    s = slots[lane * rays_per_lane + r]
                    if s[0] is not None and s[1] < len(s[0]) and int(s[0][s[1]]) == t:
                        s[1] += 1
                        useful += cost
                        break
    return useful, issued


if __name__ == "__main__":
    scene = sys.argv[1] if len(sys.argv) > 1 else "scenes/_cache/teapot.tbscene"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    o = binding.Oracle(); o.LoadScene(scene, 3)
    cam = o.GetCamera()
    rng = np.random.default_rng(1)
    eye = np.array(cam.Position.tuple(), np.float32)
    look = np.array(cam.LookAt.tuple(), np.float32) - eye
    right, up = np.array(cam.Right.tuple(), np.float32), np.array(cam.Up.tuple(), np.float32)
    # primary rays in 8x4 tiles over the lens
    w, h = 256, 128
    tiles = [(tx, ty) for ty in range(h // 4) for tx in range(w // 8)]
    rng.shuffle(tiles)
    prim = np.zeros(n, RAY_DTYPE)
    k = 0
    for tx, ty in tiles:
        for l in range(32):
            if k >= n:
                break
            x, y = tx * 8 + (l & 7), ty * 4 + (l >> 3)
            u, v = (x + 0.5) / w * 2 - 1, 1 - (y + 0.5) / h * 2
            lens = eye + right * (u * cam.LensHeight * w / h / 2) + up * (v * cam.LensHeight / 2)
            focal = eye - look / np.linalg.norm(look) * cam.FocalDistance
            d = lens - focal
            prim["Origin"][k] = focal; prim["Direction"][k] = d / np.linalg.norm(d)
            k += 1
    prim["TMin"] = 0.001; prim["TMax"] = 999999.0
    pseq, phits = visit_sequences(o, prim)
    # diffuse bounce rays from the primary hit points
    hitm = phits["t"] > 0
    org = prim["Origin"][hitm] + prim["Direction"][hitm] * phits["t"][hitm][:, None]
    d = rng.normal(0, 1, org.shape).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    sec = np.zeros(len(org), RAY_DTYPE)
    sec["Origin"] = org - prim["Direction"][hitm] * 1e-3; sec["Direction"] = d; sec["TMin"] = 0.001; sec["TMax"] = 999999.0
    sseq, _ = visit_sequences(o, sec)
    for name, seqs in (("primary (8x4 tiles)", pseq), ("diffuse bounce", sseq)):
        lens = np.array([len(s) for s in seqs])
        leaf = sum(int(s.sum()) for s in seqs)
        print("%s: %d rays, visits mean %.1f median %d p99 %d max %d, leaf share %.2f" % (
            name, len(seqs), lens.mean(), np.median(lens), np.percentile(lens, 99), lens.max(), leaf / max(1, lens.sum())))
        for rpl, thr in ((1, 20), (1, 14), (1, 8), (2, 14), (2, 24), (4, 24)):
            u, i = simulate(seqs, rpl, thr)
            print("   rays/lane %d refill<%2d: utilisation %.3f (issued %.1f M instruction-lanes)" % (rpl, thr, u / i, i / 1e6))
        print("   visit-count histogram: 0: %.3f, 1-2: %.3f, 3-8: %.3f, 9-32: %.3f, >32: %.3f" % (
            (lens == 0).mean(), ((lens >= 1) & (lens <= 2)).mean(), ((lens >= 3) & (lens <= 8)).mean(),
            ((lens >= 9) & (lens <= 32)).mean(), (lens > 32).mean()))
