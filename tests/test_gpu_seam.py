"""The SW-RT seam with the reference's ownership model (D3D12RaytracingFallback.h:83-84, 134-136):
GetRaytracingAccelerationStructurePrebuildInfo sizes caller-allocated device memory,
BuildRaytracingAccelerationStructure builds into it from caller-owned device vertex / index buffers, and ray queries
run against that caller-owned structure on the caller's stream. Device memory and streams come from PyTorch (plumbing);
every call goes through the C ABI."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mesh(rng, nv, nt):
    pos = rng.uniform(-1, 1, (nv, 3)).astype(np.float32)
    idx = rng.integers(0, nv, (nt, 3)).astype(np.uint32)
    return pos, idx


def _descs(geoms, keep):
    """geoms: list of dicts(pos [V,k>=3] float32 strided, idx or None, transform or None) -> (GeometryDesc * n) with
    DEVICE pointers (torch tensors kept alive in `keep`)."""
    import torch
    from tracerboy_b200.api import GeometryDesc
    d = (GeometryDesc * len(geoms))()
    for i, g in enumerate(geoms):
        tp = torch.from_numpy(np.ascontiguousarray(g["pos"])).cuda()
        keep.append(tp)
        d[i].Positions = tp.data_ptr(); d[i].PositionStrideBytes = g["pos"].shape[1] * 4; d[i].VertexCount = g["pos"].shape[0]
        d[i].GeometryFlags = 1
        if g.get("idx") is not None:
            raw = np.ascontiguousarray(g["idx"]).reshape(-1)
            ti = torch.from_numpy(raw.view(np.uint8)).cuda()
            keep.append(ti)
            d[i].Indices = ti.data_ptr(); d[i].IndexFormat = raw.dtype.itemsize; d[i].IndexCount = raw.size
        if g.get("transform") is not None:
            tt = torch.from_numpy(np.ascontiguousarray(g["transform"], np.float32)).cuda()
            keep.append(tt)
            d[i].Transform3x4 = tt.data_ptr()
    return d


def _host_list(geoms):
    """The same geometry for the host-pointer path (tb_bvh_build), transforms applied with the pinned arithmetic there."""
    out = []
    for g in geoms:
        out.append((np.ascontiguousarray(g["pos"][:, :3]), g.get("idx")))
    return out


@pytest.mark.parametrize("case", ["u32", "u16_strided_transform", "nonindexed_mixed"])
def test_build_into_caller_memory_and_trace_on_caller_stream(case, built):
    import torch
    import tracerboy_b200 as tb
    from tracerboy_b200.api import RAY_DTYPE, HIT_DTYPE
    rng = np.random.default_rng(7)
    if case == "u32":
        p, i = _mesh(rng, 3000, 9000)
        geoms = [dict(pos=p, idx=i)]
    elif case == "u16_strided_transform":
        p, i = _mesh(rng, 2000, 5000)
        p5 = np.concatenate([p, rng.uniform(0, 1, (2000, 2)).astype(np.float32)], 1)  # stride 20: uv after the position
        m = np.array([[1.5, 0.1, 0, 0.3], [0, 0.8, -0.2, -1.0], [0.05, 0, 1.1, 2.0]], np.float32)
        geoms = [dict(pos=p5, idx=i.astype(np.uint16), transform=m), dict(pos=p, idx=i[:1000])]
    else:
        p, i = _mesh(rng, 999, 10)
        q, j = _mesh(rng, 500, 700)
        geoms = [dict(pos=p, idx=None), dict(pos=q, idx=j.astype(np.uint16)), dict(pos=q + 3, idx=j)]
    keep = []
    d = _descs(geoms, keep)
    info = tb.prebuild_info(d, len(geoms))
    ntri = sum((g["idx"].size if g.get("idx") is not None else g["pos"].shape[0]) // 3 for g in geoms)
    assert info.ReferenceLayoutSizeInBytes == 116 * ntri - 16
    assert info.ResultDataMaxSizeInBytes >= info.ReferenceLayoutSizeInBytes + 112 * ntri - 64 and info.ScratchDataSizeInBytes > 100 * ntri
    dst = torch.zeros(info.ResultDataMaxSizeInBytes, dtype=torch.uint8, device="cuda")
    scratch = torch.empty(info.ScratchDataSizeInBytes, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()
    g = tb.TracerBoy(0)
    g.BuildRaytracingAccelerationStructureDevice(d, len(geoms), dst.data_ptr(), dst.numel(), scratch.data_ptr(), scratch.numel(),
                                                 stream.cuda_stream)
    ref_bytes = dst[:info.ReferenceLayoutSizeInBytes].cpu().numpy()

    # the same geometry through the host-pointer path of a second handle (which the oracle tests pin): identical bytes
    h = tb.TracerBoy(0)
    host = []
    for gd in geoms:
        pos = np.ascontiguousarray(gd["pos"][:, :3])
        if gd.get("transform") is not None:
            m = gd["transform"]
            x, y, z = pos[:, 0], pos[:, 1], pos[:, 2]
            pos = np.stack([((m[r, 0] * x + m[r, 1] * y) + m[r, 2] * z) + m[r, 3] for r in range(3)], 1).astype(np.float32)
        host.append((pos, gd.get("idx")))
    h.BuildRaytracingAccelerationStructure(host)
    assert np.array_equal(ref_bytes, h.GetBVH())

    # ray queries against the caller-owned structure, device rays / hits, on the caller's stream
    rays = np.zeros(50000, RAY_DTYPE)
    rays["Origin"] = rng.uniform(-3, 3, (50000, 3)); rays["Direction"] = rng.normal(0, 1, (50000, 3))
    rays["TMin"] = 0.001; rays["TMax"] = 999999.0
    d_rays = torch.from_numpy(rays.view(np.uint8)).cuda()
    d_hits = torch.zeros(50000 * 32, dtype=torch.uint8, device="cuda")
    with torch.cuda.stream(stream):
        g.TraceRaysDevice(dst.data_ptr(), dst.numel(), d_rays.data_ptr(), 50000, d_hits.data_ptr(), stream.cuda_stream)
    stream.synchronize()
    got = d_hits.cpu().numpy().view(HIT_DTYPE)
    want = h.TraceRays(rays)
    for f in got.dtype.names:
        assert np.array_equal(got[f].view(np.uint32), want[f].view(np.uint32)), f
    assert (got["t"] > 0).sum() > 1000

    # another handle recognises the structure from its own trailer (no shared host state)
    k = tb.TracerBoy(0)
    d_hits.zero_()
    k.TraceRaysDevice(dst.data_ptr(), dst.numel(), d_rays.data_ptr(), 50000, d_hits.data_ptr(), None)
    k.Synchronize()
    assert np.array_equal(d_hits.cpu().numpy().view(HIT_DTYPE)["t"].view(np.uint32), want["t"].view(np.uint32))

    # library-owned scratch (NULL) gives the same bytes; wrong sizes are E_INVALIDARG
    dst2 = torch.zeros_like(dst)
    g.BuildRaytracingAccelerationStructureDevice(d, len(geoms), dst2.data_ptr(), dst2.numel(), None, 0, None)
    assert torch.equal(dst2[:info.ReferenceLayoutSizeInBytes], dst[:info.ReferenceLayoutSizeInBytes])
    with pytest.raises(tb.TracerBoyError):
        g.BuildRaytracingAccelerationStructureDevice(d, len(geoms), dst.data_ptr(), info.ReferenceLayoutSizeInBytes, None, 0, None)
    with pytest.raises(tb.TracerBoyError):
        g.BuildRaytracingAccelerationStructureDevice(d, len(geoms), dst.data_ptr(), dst.numel(), scratch.data_ptr(), 1024, None)
    g.ForgetAccelerationStructure(dst.data_ptr())


def test_trace_rays_device_on_the_handles_scene(cornell):
    import torch
    import tracerboy_b200 as tb
    from tracerboy_b200.api import RAY_DTYPE, HIT_DTYPE
    g = tb.TracerBoy(0)
    g.LoadScene(cornell)
    rng = np.random.default_rng(3)
    rays = np.zeros(4096, RAY_DTYPE)
    rays["Origin"] = rng.uniform(-0.5, 0.5, (4096, 3)) + np.array([0, 1, 3]); rays["Direction"] = rng.normal(0, 1, (4096, 3))
    rays["TMin"] = 0.001; rays["TMax"] = 999999.0
    d_rays = torch.from_numpy(rays.view(np.uint8)).cuda()
    d_hits = torch.zeros(4096 * 32, dtype=torch.uint8, device="cuda")
    g.TraceRaysDevice(None, 0, d_rays.data_ptr(), 4096, d_hits.data_ptr(), None)
    g.Synchronize()
    want = g.TraceRays(rays)
    got = d_hits.cpu().numpy().view(HIT_DTYPE)
    for f in got.dtype.names:
        assert np.array_equal(got[f].view(np.uint32), want[f].view(np.uint32)), f


def test_set_material_rejects_dangling_indices(cornell):
    """tb_set_material validates what the kernels index unchecked (the reference leans on D3D robust buffer access)."""
    import tracerboy_b200 as tb
    g = tb.TracerBoy(0)
    g.LoadScene(cornell)
    m, _ = g.GetMaterial(0)
    m.albedoIndex = 12345
    with pytest.raises(tb.TracerBoyError) as e:
        g.SetMaterial(0, m)
    assert e.value.code == -1
    m, _ = g.GetMaterial(0)
    m.Flags |= 0x8  # mix material whose ids (albedo.x / .y) point nowhere
    m.albedo.x = 99.0
    with pytest.raises(tb.TracerBoyError):
        g.SetMaterial(0, m)
    g.Resize(16, 16)
    g.Render(tb.get_default_output_settings(), 1, 0.0)  # the device scene is still the valid one


def test_update_in_place_equals_the_oracles_update(tmp_path, built):
    """PERFORM_UPDATE (GpuBVH2Builder.cpp:165-234): a caller-owned acceleration structure refitted to
    moved vertices, in place, on caller scratch. Bytes of the reference layout and ray queries equal the oracle's
    update (which is pinned against ComputeAABBs.hlsli compiled from the mount with PERFORM_UPDATE); an update with
    unmoved vertices reproduces the build; a different triangle count is E_INVALIDARG."""
    import torch
    import tracerboy_b200 as tb
    from oracle.binding import Oracle
    from tracerboy_b200.api import RAY_DTYPE, HIT_DTYPE
    rng = np.random.default_rng(17)
    # a wavy grid: coherent geometry, so that the refitted tree is still a sensible one
    nx = 70
    gx, gy = np.meshgrid(np.linspace(-1, 1, nx), np.linspace(-1, 1, nx), indexing="ij")
    pos = np.stack([gx, gy, 0.2 * np.sin(4 * gx) * np.cos(3 * gy)], -1).reshape(-1, 3).astype(np.float32)
    idx = np.arange(nx * nx).reshape(nx, nx)
    quads = np.stack([idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, :-1], idx[1:, 1:], idx[:-1, 1:]], -1).reshape(-1, 3).astype(np.uint32)
    # the oracle's view of the same geometry: through the host-pointer build + .tbscene
    h = tb.TracerBoy(0)
    h.BuildRaytracingAccelerationStructure([(pos, quads)])
    path = str(tmp_path / "grid.tbscene")
    h.SaveScene(path)
    o = Oracle(); o.LoadScene(path, 3)
    keep = []
    d = _descs([dict(pos=pos, idx=quads)], keep)
    info = tb.prebuild_info(d, 1)
    assert info.UpdateScratchDataSizeInBytes > 4 * 3 * quads.shape[0]
    dst = torch.zeros(info.ResultDataMaxSizeInBytes, dtype=torch.uint8, device="cuda")
    scratch = torch.empty(max(info.ScratchDataSizeInBytes, info.UpdateScratchDataSizeInBytes), dtype=torch.uint8, device="cuda")
    g = tb.TracerBoy(0)
    g.BuildRaytracingAccelerationStructureDevice(d, 1, dst.data_ptr(), dst.numel(), scratch.data_ptr(), scratch.numel(), None)
    nref = info.ReferenceLayoutSizeInBytes
    built_bytes = dst[:nref].cpu().numpy()
    assert np.array_equal(built_bytes, o.GetBVH())
    # unmoved: identical bytes
    g.UpdateRaytracingAccelerationStructureDevice(d, 1, dst.data_ptr(), dst.numel(), scratch.data_ptr(), scratch.numel(), None)
    assert np.array_equal(dst[:nref].cpu().numpy(), built_bytes)
    # moved vertices (the caller overwrites its own vertex buffer, as an animation would)
    moved = pos.copy()
    moved[:, 2] = 0.2 * np.sin(4 * gx.reshape(-1) + 0.7) * np.cos(3 * gy.reshape(-1) - 0.4) + rng.normal(0, 0.01, pos.shape[0]).astype(np.float32)
    keep[0].copy_(torch.from_numpy(moved))
    g.UpdateRaytracingAccelerationStructureDevice(d, 1, dst.data_ptr(), dst.numel(), None, 0, None)   # library-owned scratch
    o.UpdateBVH(moved)
    want = o.GetBVH()
    got = dst[:nref].cpu().numpy()
    diff = np.flatnonzero(got != want)
    assert diff.size == 0, "updated BVH differs at %d bytes, first at %d" % (diff.size, diff[0])
    assert not np.array_equal(got, built_bytes)
    rays = np.zeros(40000, RAY_DTYPE)
    rays["Origin"] = rng.uniform(-1, 1, (40000, 3)) * [1.2, 1.2, 0.2] + [0, 0, 1.5]
    rays["Direction"] = rng.normal(0, 0.3, (40000, 3)) + [0, 0, -1]
    rays["TMin"] = 0.001; rays["TMax"] = 999999.0
    d_rays = torch.from_numpy(rays.view(np.uint8)).cuda()
    d_hits = torch.zeros(40000 * 32, dtype=torch.uint8, device="cuda")
    g.TraceRaysDevice(dst.data_ptr(), dst.numel(), d_rays.data_ptr(), 40000, d_hits.data_ptr(), None)
    g.Synchronize()
    hg, ho = d_hits.cpu().numpy().view(HIT_DTYPE), o.TraceRays(rays)
    for f in hg.dtype.names:
        assert np.array_equal(hg[f].view(np.uint32), ho[f].view(np.uint32)), f
    assert (hg["t"] > 0).mean() > 0.5
    # every hit lies on the MOVED surface
    hit = hg["t"] > 0
    p = rays["Origin"][hit] + rays["Direction"][hit] * hg["t"][hit][:, None]
    tri = moved[quads[hg["PrimitiveIndex"][hit]]]
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    dist = np.abs(((p - tri[:, 0]) * nrm).sum(1)) / np.maximum(np.linalg.norm(nrm, axis=1), 1e-12)
    assert dist.max() < 1e-4
    with pytest.raises(tb.TracerBoyError):
        d2 = _descs([dict(pos=pos, idx=quads[:-5])], keep)
        g.UpdateRaytracingAccelerationStructureDevice(d2, 1, dst.data_ptr(), dst.numel(), None, 0, None)


def test_two_level_structure_and_query_equal_the_oracle(tmp_path, built):
    """Top-level acceleration structure over instances of three caller-owned bottom-level structures:
    the reference-layout bytes of the top level (header, nodes, 116-byte BVHMetadata per sorted leaf: inverse transform,
    object-to-world, original index) equal the oracle's, except for the 8-byte AccelerationStructure field (a device
    address here, an index there); 100 k two-level ray queries equal the oracle's in every field, InstanceIndex and both
    counters included. Mirrored / scaled / rotated instances, one instance masked out, one bottom level of a single triangle."""
    import torch
    import tracerboy_b200 as tb
    from oracle.binding import Oracle
    from tracerboy_b200.api import RAY_DTYPE, HIT_DTYPE, InstanceDesc
    from test_cpu_tlas import _rand_affine, oracle_tlas, oracle_trace_tlas
    rng = np.random.default_rng(31)
    g = tb.TracerBoy(0)
    meshes = []
    for k, (nv, nt) in enumerate(((400, 900), (150, 260), (3, 1))):
        pos = rng.uniform(-1, 1, (nv, 3)).astype(np.float32) * (1.0 + k)
        idx = rng.integers(0, nv, (nt, 3)).astype(np.uint32) if nt > 1 else np.array([[0, 1, 2]], np.uint32)
        meshes.append((pos, idx))
    keep, blas_dev, blas_bytes = [], [], []
    for k, (pos, idx) in enumerate(meshes):
        # oracle's bytes of the same bottom level: host-pointer build + .tbscene + oracle build
        h = tb.TracerBoy(0)
        h.BuildRaytracingAccelerationStructure([(pos, idx)])
        p = str(tmp_path / ("b%d.tbscene" % k))
        h.SaveScene(p)
        o = Oracle(); o.LoadScene(p, 3)
        blas_bytes.append(np.ascontiguousarray(o.GetBVH()))
        d = _descs([dict(pos=pos, idx=idx)], keep)
        info = tb.prebuild_info(d, 1)
        dst = torch.zeros(info.ResultDataMaxSizeInBytes, dtype=torch.uint8, device="cuda")
        g.BuildRaytracingAccelerationStructureDevice(d, 1, dst.data_ptr(), dst.numel(), None, 0, None)
        assert np.array_equal(dst[:info.ReferenceLayoutSizeInBytes].cpu().numpy(), blas_bytes[-1])
        blas_dev.append(dst)
    n = 60
    mats = _rand_affine(rng, n)
    mats[:, :, 3] = rng.uniform(-25, 25, (n, 3))
    inst_gpu, inst_cpu = (InstanceDesc * n)(), (InstanceDesc * n)()
    for i in range(n):
        for arr in (inst_gpu, inst_cpu):
            for j in range(12):
                arr[i].Transform[j] = float(mats[i].reshape(-1)[j])
            arr[i].InstanceIDAndMask = (500 + i) | ((0 if i == 11 else 0xff) << 24)
            arr[i].InstanceContributionToHitGroupIndexAndFlags = 2 * i
        inst_gpu[i].AccelerationStructure = blas_dev[i % 3].data_ptr()
        inst_cpu[i].AccelerationStructure = i % 3
    tinfo = tb.tlas_prebuild_info(n)
    assert tinfo.ReferenceLayoutSizeInBytes == 16 + 32 * (2 * n - 1) + 116 * n and tinfo.ScratchDataSizeInBytes > 0
    tlas = torch.zeros(tinfo.ResultDataMaxSizeInBytes, dtype=torch.uint8, device="cuda")
    tscratch = torch.empty(tinfo.ScratchDataSizeInBytes, dtype=torch.uint8, device="cuda")
    g.BuildTopLevelAccelerationStructureDevice(inst_gpu, n, tlas.data_ptr(), tlas.numel(), tscratch.data_ptr(), tscratch.numel(), None)
    want, arr = oracle_tlas(blas_bytes, inst_cpu, n)
    got = tlas[:tinfo.ReferenceLayoutSizeInBytes].cpu().numpy()
    off_meta = 16 + 32 * (2 * n - 1)
    gm, wm = got[off_meta:].reshape(n, 116).copy(), want[off_meta:].reshape(n, 116).copy()
    ptrs = gm[:, 56:64].copy().view(np.uint64).reshape(-1)
    assert [int(p) for p in ptrs] == [blas_dev[int(i) % 3].data_ptr() for i in gm[:, 112:116].view(np.uint32).reshape(-1)]
    gm[:, 56:64] = 0; wm[:, 56:64] = 0
    assert np.array_equal(got[:off_meta], want[:off_meta]), "header / nodes"
    assert np.array_equal(gm, wm), "instance metadata"
    R = 100000
    rays = np.zeros(R, RAY_DTYPE)
    rays["Origin"] = rng.uniform(-40, 40, (R, 3))
    tgt = mats[rng.integers(0, n, R), :, 3] + rng.normal(0, 1.5, (R, 3))
    dd = tgt - rays["Origin"]
    rays["Direction"] = dd / np.linalg.norm(dd, axis=1, keepdims=True)
    rays["Direction"][::97, 1] = 0.0           # exactly-zero components (D6) ...
    rays["Direction"][5::1013] = np.nan        # ... and NaN rays (D7)
    rays["TMin"] = 0.001; rays["TMax"] = 999999.0
    rays["TMax"][::50] = 20.0
    d_rays = torch.from_numpy(rays.view(np.uint8)).cuda()
    d_hits = torch.zeros(R * 32, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()
    g.TraceRaysTopLevelDevice(tlas.data_ptr(), tlas.numel(), d_rays.data_ptr(), R, d_hits.data_ptr(), stream.cuda_stream)
    stream.synchronize()
    hg, ho = d_hits.cpu().numpy().view(HIT_DTYPE), oracle_trace_tlas(want, arr, 3, rays)
    for f in hg.dtype.names:
        same = hg[f].view(np.uint32) == ho[f].view(np.uint32)
        assert same.all(), "field %s differs for %d rays, first %d" % (f, (~same).sum(), np.flatnonzero(~same)[0])
    hit = hg["t"] > 0
    assert hit.mean() > 0.2 and len(np.unique(hg["InstanceIndex"][hit])) > 30 and not (hg["InstanceIndex"][hit] == 11).any()
    # another handle recognises the structure from its own bytes
    k = tb.TracerBoy(0)
    d_hits.zero_()
    k.TraceRaysTopLevelDevice(tlas.data_ptr(), tlas.numel(), d_rays.data_ptr(), R, d_hits.data_ptr(), None)
    k.Synchronize()
    assert np.array_equal(d_hits.cpu().numpy().view(HIT_DTYPE)["t"].view(np.uint32), ho["t"].view(np.uint32))
    with pytest.raises(tb.TracerBoyError):
        g.BuildTopLevelAccelerationStructureDevice(inst_gpu, n, tlas.data_ptr(), 1000, tscratch.data_ptr(), tscratch.numel(), None)
    with pytest.raises(tb.TracerBoyError):
        g.BuildTopLevelAccelerationStructureDevice(inst_gpu, n, tlas.data_ptr(), tlas.numel(), tscratch.data_ptr(), 1000, None)


@pytest.mark.parametrize("n", [1, 2, 3, 9, 5000])
def test_top_level_bytes_equal_the_oracle_from_one_to_thousands_of_instances(n, tmp_path, built):
    """The GPU top-level build (k_tlas_load, Morton, radix sort, Karras, k_tlas_emit, k_tlas_refit) against the oracle's
    bytes for 1, 2, 3, 9 and 5 000 instances: a single leaf without hierarchy, the smallest trees, and a tree whose
    instances crowd into a few Morton cells (many equal codes: the sort's tie rule and the Karras tie rule matter) plus
    coincident instances (the same transform twice)."""
    import torch
    import tracerboy_b200 as tb
    from oracle.binding import Oracle
    from tracerboy_b200.api import InstanceDesc
    from test_cpu_tlas import _rand_affine, oracle_tlas
    rng = np.random.default_rng(100 + n)
    g = tb.TracerBoy(0)
    keep, blas_dev, blas_bytes = [], [], []
    for k, (nv, nt) in enumerate(((60, 100), (3, 1))):
        pos = rng.uniform(-1, 1, (nv, 3)).astype(np.float32)
        idx = rng.integers(0, nv, (nt, 3)).astype(np.uint32) if nt > 1 else np.array([[0, 1, 2]], np.uint32)
        h = tb.TracerBoy(0)
        h.BuildRaytracingAccelerationStructure([(pos, idx)])
        p = str(tmp_path / ("b%d.tbscene" % k))
        h.SaveScene(p)
        o = Oracle(); o.LoadScene(p, 3)
        blas_bytes.append(np.ascontiguousarray(o.GetBVH()))
        d = _descs([dict(pos=pos, idx=idx)], keep)
        info = tb.prebuild_info(d, 1)
        dst = torch.zeros(info.ResultDataMaxSizeInBytes, dtype=torch.uint8, device="cuda")
        g.BuildRaytracingAccelerationStructureDevice(d, 1, dst.data_ptr(), dst.numel(), None, 0, None)
        blas_dev.append(dst)
    mats = _rand_affine(rng, n)
    mats[:, :, 3] = rng.uniform(-25, 25, (n, 3))
    if n >= 9:
        mats[n // 2:, :, 3] = rng.uniform(-0.05, 0.05, (n - n // 2, 3)) + 20.0   # half of them within one Morton cell
        mats[4] = mats[3]                                                            # coincident instances
    inst_gpu, inst_cpu = (InstanceDesc * n)(), (InstanceDesc * n)()
    for i in range(n):
        for arr in (inst_gpu, inst_cpu):
            for j in range(12):
                arr[i].Transform[j] = float(mats[i].reshape(-1)[j])
            arr[i].InstanceIDAndMask = i | (0xff << 24)
            arr[i].InstanceContributionToHitGroupIndexAndFlags = i
        inst_gpu[i].AccelerationStructure = blas_dev[i % 2].data_ptr()
        inst_cpu[i].AccelerationStructure = i % 2
    tinfo = tb.tlas_prebuild_info(n)
    tlas = torch.zeros(tinfo.ResultDataMaxSizeInBytes, dtype=torch.uint8, device="cuda")
    tscratch = torch.empty(tinfo.ScratchDataSizeInBytes, dtype=torch.uint8, device="cuda")
    g.BuildTopLevelAccelerationStructureDevice(inst_gpu, n, tlas.data_ptr(), tlas.numel(), tscratch.data_ptr(), tscratch.numel(), None)
    want, _ = oracle_tlas(blas_bytes, inst_cpu, n)
    got = tlas[:tinfo.ReferenceLayoutSizeInBytes].cpu().numpy()
    off_meta = 16 + 32 * (2 * n - 1)
    gm, wm = got[off_meta:].reshape(n, 116).copy(), want[off_meta:].reshape(n, 116).copy()
    gm[:, 56:64] = 0; wm[:, 56:64] = 0
    assert np.array_equal(got[:off_meta], want[:off_meta]), "header / nodes"
    assert np.array_equal(gm, wm), "instance metadata"
    # and the build is repeatable into the same memory
    g.BuildTopLevelAccelerationStructureDevice(inst_gpu, n, tlas.data_ptr(), tlas.numel(), tscratch.data_ptr(), tscratch.numel(), None)
    assert np.array_equal(tlas[:off_meta].cpu().numpy(), got[:off_meta])
