"""Two-level ray query at the SW-RT seam: build N bottom levels + one top level over M instances in caller-owned memory,
then time tb_trace_rays_tlas_device and, for comparison, tb_trace_rays_device on ONE flattened bottom level of the same
triangles (CUDA events on the caller's stream). python tools/tlas_query_bench.py [instances] [tris per mesh] [rays]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tracerboy_b200 as tb
from tracerboy_b200.api import GeometryDesc, InstanceDesc, RAY_DTYPE

M = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
R = int(sys.argv[3]) if len(sys.argv) > 3 else 4000000
rng = np.random.default_rng(1)
g = tb.TracerBoy(0)


def sphere(tris, seed):
    """a displaced sphere of about `tris` triangles"""
    r = np.random.default_rng(seed)
    rings = max(3, int(np.sqrt(tris / 2.5)))
    segs = max(3, tris // (2 * rings))
    th = np.linspace(0, np.pi, rings + 1)[:, None]
    ph = np.linspace(0, 2 * np.pi, segs, endpoint=False)[None, :]
    rad = 1.0 + 0.15 * r.standard_normal((rings + 1, segs))
    p = np.stack([rad * np.sin(th) * np.cos(ph), rad * np.cos(th) * np.ones_like(ph), rad * np.sin(th) * np.sin(ph)], -1).reshape(-1, 3)
    idx = []
    for a in range(rings):
        for b in range(segs):
            i0, i1 = a * segs + b, a * segs + (b + 1) % segs
            idx += [[i0, i1, i0 + segs], [i1, i1 + segs, i0 + segs]]
    return p.astype(np.float32), np.array(idx, np.uint32)


def build(pos, idx):
    dpos = torch.from_numpy(pos).cuda()
    didx = torch.from_numpy(idx.view(np.uint8).reshape(-1)).cuda()
    d = (GeometryDesc * 1)()
    d[0].Positions = dpos.data_ptr(); d[0].PositionStrideBytes = 12; d[0].VertexCount = pos.shape[0]
    d[0].Indices = didx.data_ptr(); d[0].IndexFormat = 4; d[0].IndexCount = idx.size; d[0].GeometryFlags = 1
    info = tb.prebuild_info(d, 1)
    dst = torch.zeros(info.ResultDataMaxSizeInBytes, dtype=torch.uint8, device="cuda")
    scratch = torch.empty(info.ScratchDataSizeInBytes, dtype=torch.uint8, device="cuda")
    g.BuildRaytracingAccelerationStructureDevice(d, 1, dst.data_ptr(), dst.numel(), scratch.data_ptr(), scratch.numel(), None)
    return dst, (dpos, didx)


meshes = [sphere(T, s) for s in range(4)]
blas = [build(*m) for m in meshes]
side = int(np.ceil(M ** (1 / 3)))
cell = 200.0 / side
inst = (InstanceDesc * M)()
flat_pos, flat_idx, voff = [], [], 0
for i in range(M):
    x, y, z = i % side, (i // side) % side, i // (side * side)
    c = -100.0 + cell * (np.array([x, y, z]) + 0.5 + 0.2 * (rng.random(3) - 0.5))
    s = 0.36 * cell
    a = rng.random() * 2 * np.pi
    rot = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]) * s
    m = np.concatenate([rot, c[:, None]], 1).astype(np.float32)
    for j in range(12):
        inst[i].Transform[j] = float(m.reshape(-1)[j])
    inst[i].InstanceIDAndMask = i | (0xff << 24)
    inst[i].InstanceContributionToHitGroupIndexAndFlags = i
    inst[i].AccelerationStructure = blas[i % 4][0].data_ptr()
    p, ix = meshes[i % 4]
    flat_pos.append(p @ m[:, :3].T + m[:, 3]); flat_idx.append(ix + voff); voff += p.shape[0]
tinfo = tb.tlas_prebuild_info(M)
tlas = torch.zeros(tinfo.ResultDataMaxSizeInBytes, dtype=torch.uint8, device="cuda")
tscratch = torch.empty(tinfo.ScratchDataSizeInBytes, dtype=torch.uint8, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda.synchronize()
stream = torch.cuda.Stream()             # a real stream: handle 0 would mean "the library's own stream"
sp = C.c_void_p(stream.cuda_stream)
for _ in range(2):
    ev[0].record(stream)
    g.BuildTopLevelAccelerationStructureDevice(inst, M, tlas.data_ptr(), tlas.numel(), tscratch.data_ptr(), tscratch.numel(), sp)
    ev[1].record(stream); torch.cuda.synchronize()
print("top level over %d instances: %.3f ms (host resolve + upload + GPU build, wall by CUDA events)" % (M, ev[0].elapsed_time(ev[1])))
flat, keep = build(np.concatenate(flat_pos).astype(np.float32), np.concatenate(flat_idx).astype(np.uint32))
tris_total = sum(ix.shape[0] for ix in flat_idx)
# rays: a 1080p-like camera bundle from outside + incoherent rays from inside the cube
rays = np.zeros(R, RAY_DTYPE)
eye = np.array([150.0, 90.0, 330.0], np.float32)
tgt = rng.uniform(-100, 100, (R, 3)).astype(np.float32)
rays["Origin"] = eye; rays["Direction"] = tgt - eye
k = R // 2
rays["Origin"][k:] = rng.uniform(-100, 100, (R - k, 3)); rays["Direction"][k:] = rng.standard_normal((R - k, 3))
rays["TMin"] = 0.001; rays["TMax"] = 1e6
d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
d_hits = torch.zeros(R * 32, dtype=torch.uint8, device="cuda")
for name, fn in (("two-level (top level + %d instances of %d-triangle bottom levels)" % (M, T), lambda: g.TraceRaysTopLevelDevice(tlas.data_ptr(), tlas.numel(), d_rays.data_ptr(), R, d_hits.data_ptr(), sp)),
                 ("single level (%d triangles flattened)" % tris_total, lambda: g.TraceRaysDevice(flat.data_ptr(), flat.numel(), d_rays.data_ptr(), R, d_hits.data_ptr(), sp))):
    fn(); torch.cuda.synchronize()
    ev[0].record(stream)
    for _ in range(3):
        fn()
    ev[1].record(stream); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 3
    hits = d_hits.cpu().numpy().view(tb.api.HIT_DTYPE)
    print("%s: %.2f ms for %d rays = %.0f Mrays/s, %.1f %% hit" % (name, ms, R, R / ms / 1e3, 100 * (hits["t"] > 0).mean()))
