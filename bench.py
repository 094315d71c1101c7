#!/usr/bin/env python
"""bench.py — headline benchmark of the path-tracing hot path (contract in the task brief).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--shard samples|rows]
                  [--scaling weak|strong]

One "step" = one pass of the hot path over one batch: `--spp` samples per pixel of the workload scene
(default: Teapot 1920x1080, 64 spp, 6 bounces = BASELINE.json configs[1]) through tb_render. ONE JSON line on rank 0.

  value     Mrays/s, whole job (all ranks), inputs resident in HBM. Time = device time of every step, CUDA events
            recorded by the library on its own stream around the frames of tb_render and around the reduction
            (tb_comm_reduce), max over ranks; the same clock for every N.
  e2e       same metric through the public API with host buffers: every step does the host->device copy of the
            per-frame inputs (settings + camera, pinned) and the device->host read of the resolved JOB-WIDE image
            (for N > 1 that read runs the NCCL reduction), wall clock bracketed by barriers + synchronisation
  roofline  the dominant kernel (k_extend): algorithmic bytes over its time in the timed mode; what binds it per ncu
  cpu_baseline  the CPU oracle timed on this box's host cores on a bounded sample (N = 1 only)

N > 1 (torchrun): one process per GPU, one TbHandle each, joined by the library's own communicator (tb_comm_init: NCCL
over NVLink; the id is broadcast with torch.distributed, which is plumbing only). --shard samples: frame f on rank
f mod N, reduction = all-gather + fixed rank-order sum; --shard rows: bands of 8 rows, bit-identical to one GPU.
--scaling weak (default): every GPU renders `spp` frames per step (or all frames of its bands); strong: `spp` in total.
--impl reference: the reference's CPU implementation of the path (CPU oracle / oracle/_ref).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries exactly one JSON line (NCCL prints its version banner there otherwise)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # frames in flight: one HW queue per stream, before CUDA starts
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scene spec, width, height, spp, max bounces)
    "teapot": ("teapot", 1920, 1080, 64, 6),                   # configs[1]
    "cornell": ("cornell-box", 512, 512, 16, 4),               # configs[0]
    "dragon": ("dragon", 1920, 1080, 256, 8),                  # configs[2]: the reference's scene.pbrt, variant (tracerboy_b200/build.py)
    "vwvan": ("vw-van", 3840, 2160, 128, 6),                   # configs[3], variant scene (tracerboy_b200/build.py)
    "blobs20m": ("synthetic:blobs?copies=20000&tris=1000&seed=1", 1920, 1080, 1024, 6),   # configs[4]
    "blobs871k": ("synthetic:blobs?copies=1&tris=871000&seed=1", 1920, 1080, 256, 8),     # round-1 stand-in: area light, glass, metal
}


def scene_arg(spec):
    if spec.startswith("synthetic:"):
        return spec
    for d in ("scenes/_cache", "tests/golden"):
        p = os.path.join(ROOT, d, spec + ".tbscene")
        if os.path.exists(p):
            return p
    raise SystemExit("scene cache for '%s' is missing (built by __graft_entry__.build() from the reference mount)" % spec)


def tbscene_for_oracle(spec):
    """The oracle only reads .tbscene; synthetic specs are converted through the host-only C ABI call."""
    import hashlib
    import tempfile
    import tracerboy_b200 as tb
    if not spec.startswith("synthetic:"):
        return scene_arg(spec)
    out = os.path.join(tempfile.gettempdir(), "tb_bench_synth_%s.tbscene" % hashlib.sha1(spec.encode()).hexdigest()[:10])
    if not os.path.exists(out):
        tb.convert_scene(spec, out)
    return out


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            text, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            text, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        for line in text.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_units(workload):
    """What binds k_extend on this workload according to the committed `ncu --set full` capture of the same command
    (profiles/r2_ncu_units.json, written by tools/ncu_units.py from the raw CSV): per-unit utilisation, DRAM / L2 bytes."""
    p = os.path.join(ROOT, "profiles", "r2_ncu_units.json")
    try:
        return json.load(open(p)).get(workload)
    except Exception:
        return None


def cpu_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


CPU_PARTS = {"core (PathTrace / Trace / BRDFs / material model)": "reference text: TracerBoy/kernel.glsl compiled from the mount as host C++ (oracle/_ref/libref_core.so)",
             "ray query, BVH builder, ray-generation glue": "hand restatement (oracle/*.cpp), each stage pinned bit-exact against the reference's own shader text compiled from the mount (tests/test_cpu_oracle.py)"}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (restated core; oracle/_ref when built) on host cores."""
    if rank != 0:
        return
    from oracle import binding
    from oracle.binding import Oracle
    import tracerboy_b200 as tb
    spec, w, h, spp, bounces = WORKLOADS[args.workload]
    # oracle/_ref = the reference's own kernel.glsl compiled as host C++ (built from the mount, travels as a
    # binary); without it the hand-restated core runs ("port"). BVH build, traversal and glue are restated either way.
    kind = "port"
    if binding.reference_core_available():
        binding.use_reference_core(True)
        kind = "reference"
    o = Oracle()
    o.LoadScene(tbscene_for_oracle(spec), 3)
    o.Resize(w, h)
    s = tb.get_default_output_settings()
    s.MaxBounces = bounces
    cores = cpu_cores()  # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1 for its workers)
    sample_spp = max(1, args.ref_spp)
    for _ in range(args.warmup):
        o.Render(s, 1, 0.0, threads=cores)
    c0 = o.Counts()
    t = 0.0
    for _ in range(args.steps):
        t += o.Render(s, sample_spp, 0.0, threads=cores)
    c1 = o.Counts()
    rays = c1["rays"] - c0["rays"]
    value = rays / t / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic camera/seeds on the bundled scene",
        "config": {"workload": "%s %dx%d, %d bounces, CPU sample of %d spp per step (a rate metric: Mrays/s does not depend on the sample count)" % (
            args.workload, w, h, bounces, sample_spp)},
        "samples_per_s": w * h * sample_spp * args.steps / t,
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind, "parts": CPU_PARTS if kind == "reference" else None,
                         "sample": "%d spp of the %dx%d %s workload per step, OpenMP over pixels" % (sample_spp, w, h, args.workload)},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="teapot", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="samples per pixel per step (default: the workload's)")
    ap.add_argument("--ref-spp", type=int, default=2, help="CPU sample size per step for --impl reference")
    ap.add_argument("--cpu-baseline-spp", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fif", type=int, default=0, help="frames in flight (0: the library's automatic policy)")
    ap.add_argument("--shard", default="samples", choices=["samples", "rows"],
                    help="N>1: 'samples' = frame f on rank f mod N (fixed rank-order sum); 'rows' = interleaved bands of 8 rows "
                         "(bit-identical to one GPU)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = every GPU renders `spp` frames per step; strong = `spp` frames per step in total")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import tracerboy_b200 as tb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    spec, w, h, spp, bounces = WORKLOADS[args.workload]
    if args.spp:
        spp = args.spp
    t_load = time.time()
    g = tb.TracerBoy(local_rank)
    g.LoadScene(scene_arg(spec))
    load_s = time.time() - t_load
    build_first_ms = g.GetBVHBuildMilliseconds()  # includes the first-use costs: module load, first touch of the scratch
    g.LoadScene(scene_arg(spec))                  # the same build again: the steady-state builder time
    g.Resize(w, h)
    if args.fif:
        g.SetFramesInFlight(args.fif)
    rows = args.shard == "rows"
    if world > 1:
        # the product's own communicator; torch.distributed only carries the 128-byte id to the other processes
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(tb.comm_get_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        g.CommInit(bytes(uid.cpu().numpy().tobytes()), rank, world, tb.SHARD_ROWS if rows else tb.SHARD_SAMPLES)
    # frames this rank renders per step
    if rows:
        spp_rank = spp                                   # all frames of its own bands: strong scaling by construction
        frames_job = spp
    elif args.scaling == "strong":
        spp_rank = max(1, spp // world)
        frames_job = spp_rank * world
    else:
        spp_rank = spp
        frames_job = spp * world
    strong = world > 1 and (rows or args.scaling == "strong")
    s = tb.get_default_output_settings()
    s.MaxBounces = bounces
    info = g.GetSceneInfo()

    # pinned host buffers for the e2e leg
    host_img = torch.empty((h, w, 3), dtype=torch.float32).pin_memory()
    host_in = torch.empty(ctypes.sizeof(s) + ctypes.sizeof(tb.Camera), dtype=torch.uint8).pin_memory()
    cam = g.GetCamera()

    def step_resident():
        g.InvalidateHistory()
        g.Render(s, spp_rank, 0.0)
        if world > 1:
            g.CommReduce()  # the path's one exchange step (collective): job-wide accumulation buffers on every rank

    def step_e2e():
        # host -> device: this step's inputs (PerFrameConstants sources) from pinned memory
        ctypes.memmove(host_in.data_ptr(), ctypes.addressof(s), ctypes.sizeof(s))
        ctypes.memmove(host_in.data_ptr() + ctypes.sizeof(s), ctypes.addressof(cam), ctypes.sizeof(cam))
        s_in = tb.OutputSettings.from_address(host_in.data_ptr())
        cam_in = tb.Camera.from_address(host_in.data_ptr() + ctypes.sizeof(s))
        g.SetCamera(cam_in)  # also invalidates history, like TracerBoy::Update
        g.Render(s_in, spp_rank, 0.0)
        # device -> host: the resolved image (rgb / w) of the WHOLE JOB (runs the reduction when N > 1), read by the caller
        g.Readback(tb.BufferKind.RESOLVED_RGB, out=host_img.numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        g.ResetRenderStats()
        g.Synchronize()
        red0 = g.CommInfo().TotalReductionMilliseconds if world > 1 else 0.0
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        g.Synchronize()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        barrier()
        st = g.GetRenderStats()
        red_ms = (g.CommInfo().TotalReductionMilliseconds - red0) if world > 1 else 0.0
        # device time: the library's CUDA events around the frames of every tb_render + around every reduction
        return st, st.DeviceMilliseconds + red_ms, wall, red_ms

    for _ in range(max(3, args.warmup)):
        step_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    st, dev_ms, wall, red_ms = timed(step_resident, args.steps)
    clocks = sampler.stop()
    vals = torch.tensor([dev_ms / 1e3, wall, red_ms / 1e3], dtype=torch.float64, device="cuda")
    sums = torch.tensor([float(st.RaysTraced), float(st.KernelLaunches), float(st.BoxesTested), float(st.TrianglesTested)] +
                        [float(x) for x in st.RaysByBounce], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    t_step, wall, red_s = vals[0].item(), vals[1].item(), vals[2].item()
    rays_total, launches, boxes_total, tris_total = (sums[i].item() for i in range(4))
    rays_by_bounce = [sums[4 + b].item() for b in range(32)]
    value = rays_total / t_step / 1e6

    # the same step with primary rays only (MaxBounces = 1: camera ray + its shadow ray): the difference is what the
    # incoherent bounces cost in the timed mode, frames in flight and all
    incoherent = None
    if bounces > 1:
        s1 = tb.get_default_output_settings()
        s1.MaxBounces = 1

        def step_primary():
            g.InvalidateHistory()
            g.Render(s1, spp_rank, 0.0)
        step_primary()
        st1, dev1_ms, _, _ = timed(step_primary, args.steps)
        v1 = torch.tensor([dev1_ms / 1e3], dtype=torch.float64, device="cuda")
        r1 = torch.tensor([float(st1.RaysTraced)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(v1, op=dist.ReduceOp.MAX)
            dist.all_reduce(r1, op=dist.ReduceOp.SUM)
        t_render = t_step - red_s
        if t_render > v1[0].item() and rays_total > r1[0].item():
            incoherent = {"mrays_per_s": (rays_total - r1[0].item()) / (t_render - v1[0].item()) / 1e6,
                          "primary_only_mrays_per_s": r1[0].item() / v1[0].item() / 1e6,
                          "rays_share": (rays_total - r1[0].item()) / rays_total,
                          "how": "(rays(all bounces) - rays(MaxBounces=1)) / (device time(all bounces) - device time(MaxBounces=1)), same timed mode"}

    # e2e leg
    for _ in range(2):
        step_e2e()
    st_e, _, wall_e, _ = timed(step_e2e, args.steps)
    ve = torch.tensor([wall_e], dtype=torch.float64, device="cuda")
    re_ = torch.tensor([float(st_e.RaysTraced)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ve, op=dist.ReduceOp.MAX)
        dist.all_reduce(re_, op=dist.ReduceOp.SUM)
    wall_e, rays_e = ve[0].item(), re_[0].item()
    e2e_value = rays_e / wall_e / 1e6

    # ---- roofline of the dominant kernel
    # (a) shares in the timed mode: per-launch CUDA events on each frame slot's own stream while the frames overlap as
    #     in the timed region (profiling mode 2, no graph replay)
    g.SetProfiling(2)
    g.ResetRenderStats()
    g.InvalidateHistory()
    g.Render(s, spp_rank, 0.0)
    cst = g.GetRenderStats()
    # (b) exclusive times: one frame at a time (profiling mode 1), also per bounce
    g.SetProfiling(1)
    g.ResetRenderStats()
    g.InvalidateHistory()
    g.Render(s, min(spp_rank, 16), 0.0)
    pst = g.GetRenderStats()
    g.SetProfiling(0)
    peak, peak_src = measured_peak_gbs()
    # `achieved`: what the traversal stage achieves in the TIMED mode -- the algorithmic bytes of every ray of the timed
    # region (extension, shadow and walk rays; all k_extend kinds, k_extend_resume, k_walk) over the whole device time of
    # the region, shading and accumulation included: a lower bound of the traversal rate that needs no attribution of
    # overlapped kernels. With 32 frames in flight the launches of different frames overlap and fill each other's tails
    # (the step is ~2x shorter than the sum of its launches timed alone), so a per-launch duration only exists for a
    # launch that has the GPU to itself: that is measured too (profiling mode 1, one frame at a time, CUDA events around
    # every k_extend<EXT_MAIN> launch) and reported as `exclusive`, with the kernel's share of the step for the
    # cross-check against the serialised ncu launch list in profiles/.
    share = pst.ExtendMilliseconds / max(1e-9, pst.ExtendMilliseconds + pst.ShadeMilliseconds + pst.ResumeMilliseconds)
    share_concurrent = cst.ExtendMilliseconds / max(1e-9, cst.ExtendMilliseconds + cst.ShadeMilliseconds + cst.ResumeMilliseconds)
    serial_alg = 32.0 * pst.ExtendBoxesTested + 40.0 * pst.ExtendTrianglesTested + 64.0 * pst.ExtendRays
    exclusive_gbs = serial_alg / max(1e-9, pst.ExtendMilliseconds / 1e3) / 1e9
    alg_step = (32.0 * boxes_total + 40.0 * tris_total + 64.0 * rays_total) / args.steps   # whole job, per step
    achieved = alg_step / (t_step / args.steps) / 1e9 / world                                # per GPU
    units = ncu_units(args.workload)
    bound = "unknown (no ncu capture of this workload in profiles/r2_ncu_units.json)"
    if units:
        pct = {"issue": units.get("issue_active_pct") or 0.0, "l1": units.get("l1tex_throughput_pct") or 0.0,
               "l2": units.get("lts_throughput_pct") or 0.0, "hbm": units.get("dram_throughput_pct") or 0.0}
        bound = max(pct, key=pct.get)
    per_bounce = []
    for b in range(min(bounces, 32)):
        if pst.RaysByBounce[b]:
            per_bounce.append({"bounce": b, "rays": int(pst.RaysByBounce[b]), "ms": pst.BounceMilliseconds[b],
                               "mrays_per_s": pst.RaysByBounce[b] / max(1e-9, pst.BounceMilliseconds[b]) / 1e3,
                               "extend_ms": pst.BounceExtendMilliseconds[b]})
    roofline = {
        "bound": bound, "kernel": "traversal stage (k_extend<MAIN|SHADOW|WALK>, k_extend_resume, k_walk)", "achieved": achieved, "peak": peak,
        "unit": "GB/s", "frac": achieved / peak,
        "traffic": (units or {}).get("dram_bytes_per_launch"),
        "traffic_source": (units or {}).get("source"),
        "algorithmic_bytes_per_step": alg_step / world, "ms_per_step": 1e3 * t_step / args.steps,
        "how": "timed mode, per GPU: (32 B x BoxesTested + 40 B x TrianglesTested + 64 B x rays) of ALL rays of the timed region / device time "
               "of the region (the same CUDA events as `value`), shading included; no attribution of overlapped launches is needed",
        "kernel_share_of_step": share,
        "units_pct_of_peak": units,
        "exclusive": {  # k_extend<EXT_MAIN> alone on the GPU: one frame at a time, CUDA events around every launch on its stream
            "kernel": "k_extend<EXT_MAIN>", "achieved": exclusive_gbs, "frac": exclusive_gbs / peak,
            "algorithmic_bytes_per_launch": serial_alg / max(1, pst.ExtendLaunches),
            "avg_launch_ms": pst.ExtendMilliseconds / max(1, pst.ExtendLaunches), "launches": pst.ExtendLaunches,
            "frames": min(spp_rank, 16), "kernel_share_of_step": share,
            "resume_rounds": {"rays": pst.ResumeRays, "ms": pst.ResumeMilliseconds},
            "note": "includes each launch's tail (a few rays with thousands of node visits), which the timed mode hides behind other frames"},
        "kernel_share_of_stream_time_under_concurrency": share_concurrent,
        "peak_source": peak_src,
        "note": "algorithmic bytes on the reference BVH2 layout (SURVEY 8d). `peak` is the measured HBM copy bandwidth; a BVH that fits the "
                "126 MB L2 is served from L1/L2, so `bound` names the unit ncu shows closest to its own peak (units_pct_of_peak) and frac "
                "is NOT a DRAM utilisation there",
    }

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import binding
        from oracle.binding import Oracle
        kind = "port"
        if binding.reference_core_available():
            binding.use_reference_core(True)
            kind = "reference"
        o = Oracle()
        o.LoadScene(tbscene_for_oracle(spec), 3)
        o.Resize(w, h)
        cores = cpu_cores()
        o.Render(s, 1, 0.0, threads=cores)
        c0 = o.Counts()
        sec = o.Render(s, args.cpu_baseline_spp, 0.0, threads=cores)
        c1 = o.Counts()
        cpu_baseline = {"value": (c1["rays"] - c0["rays"]) / sec / 1e6, "unit": "Mrays/s", "cores": cores,
                        "kind": kind, "parts": CPU_PARTS if kind == "reference" else None,
                        "sample": "%d spp of the %dx%d %s workload (%.1f s), OpenMP over pixels" % (
                            args.cpu_baseline_spp, w, h, args.workload, sec)}

    if rank == 0:
        comm = g.CommInfo() if world > 1 else None
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": 1e3 * t_step / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic camera/seeds on the bundled scene",
            "config": {"workload": "%s %dx%d, %d spp per step %s, %d bounces, NEE on, blue noise on, Time=0" % (
                           args.workload, w, h, frames_job, "in total" if strong else "per GPU" if world > 1 else "", bounces),
                       "triangles": info.NumTriangles, "bvh_depth": g.GetBVHDepth(),
                       "sharding": None if world == 1 else (("bands of 8 rows, band b on rank b mod N; NCCL all-gather of the owned bands (bit-identical to one GPU)" if rows else
                                                            "frame f on rank f mod N; NCCL all-gather + fixed rank-order sum") + ", inside the library (tb_comm_reduce), once per step"),
                       "l2": "inputs exceed L2: %d MB of path state + accumulation buffers are rewritten every sample" % (
                           (w * h * 16 * 14) >> 20),
                       "frames_in_flight": args.fif if args.fif else "auto (memory budget, <= 32)",
                       "timing": "device: CUDA events of tb_render + tb_comm_reduce on the library's stream, max over ranks"},
            "samples_per_s": w * h * frames_job * args.steps / t_step,
            "rays_per_step": rays_total / args.steps,
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "reduce_ms_per_step": 1e3 * red_s / args.steps if world > 1 else None,
            "comm": None if comm is None else {"nccl_version": comm.NcclVersion, "bytes_received_per_reduction": comm.BytesReceivedPerReduction,
                                                   "transport": "peer memory (one kernel per rank over CUDA IPC mappings, NVLink)" if comm.Transport == 1
                                                   else "nccl all-gather + combine kernels"},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(host_in.numel()),
                    "d2h_bytes_per_step": int(host_img.numel() * 4), "ms_per_step": 1e3 * wall_e / args.steps,
                    "image": "resolved rgb of the whole job (after the reduction)" if world > 1 else "resolved rgb"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "rays_by_bounce": [int(x) for x in rays_by_bounce[:bounces]],
            "incoherent": incoherent,
            "per_bounce_exclusive": per_bounce,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "bvh_build_ms": g.GetBVHBuildMilliseconds(), "bvh_build_first_call_ms": build_first_ms,
            "bvh_build_mtris_per_s": g.GetSceneInfo().NumTriangles / max(1e-9, g.GetBVHBuildMilliseconds()) / 1e3, "scene_load_s": load_s,
            # the builder against the HBM roofline: compulsory traffic of its stages per triangle: load 48 + 52,
            # Morton 40 + 8, four radix passes 4 x 32, rearrange 108, Karras ~24, three treelet passes ~3 x 100, refit 96 + 32,
            # traversal layout 116 + 112 = 1064 B
            "bvh_build_roofline": {"bound": "hbm", "algorithmic_bytes_per_triangle": 1064,
                                   "achieved": 1064.0 * info.NumTriangles / max(1e-9, g.GetBVHBuildMilliseconds() / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": 1064.0 * info.NumTriangles / max(1e-9, g.GetBVHBuildMilliseconds() / 1e3) / 1e9 / peak,
                                   "note": "small scenes are bound by the serial climb at the top of the tree (launch / latency), large ones by the "
                                           "treelet search's shared-memory traffic (profiles/r1c_treelet_octets_20m_ncu_full_raw.csv)"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        g.CommDestroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
