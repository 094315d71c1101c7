#!/usr/bin/env python
"""bench.py — headline benchmark of the path-tracing hot path (contract in the task brief).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one batch: `--spp` samples per pixel of the
workload scene at 1920x1080 (default: Teapot, 64 spp, BASELINE.json configs[1]) through
tb_render. Prints ONE JSON line on rank 0.

  value     Mrays/s, whole job (all ranks), inputs resident in HBM, CUDA-event timed
  e2e       same metric through the public API with host buffers: every step does the
            host->device copy of the per-frame inputs (settings + camera, pinned) and the
            device->host read of the resolved image, inside the timed region
  roofline  dominant kernel (k_extend): algorithmic bytes / summed CUDA-event launch time
  cpu_baseline  the CPU oracle timed on this box's host cores on a bounded sample (N=1 only)

N>1 (torchrun): sample-index sharding (frame f on rank f mod N), one NCCL all-reduce of the
float4 accumulation buffer per step (the path's only exchange step); weak scaling.
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
    "teapot": ("teapot", 1920, 1080, 64, 6),
    "cornell": ("cornell-box", 512, 512, 16, 4),
    "dragon": ("synthetic:blobs?copies=1&tris=871000&seed=1", 1920, 1080, 256, 8),
    "vwvan": ("vw-van", 3840, 2160, 128, 6),  # configs[3], variant scene (tracerboy_b200/build.py)
    "blobs20m": ("synthetic:blobs?copies=20000&tris=1000&seed=1", 1920, 1080, 1024, 6),
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
    import tracerboy_b200 as tb
    if not spec.startswith("synthetic:"):
        return scene_arg(spec)
    out = os.path.join(ROOT, "scenes", "_cache", "bench_synth_%08x.tbscene" % (hash(spec) & 0xffffffff))
    os.makedirs(os.path.dirname(out), exist_ok=True)
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


def ncu_traffic():
    """dram bytes per k_extend launch from the committed ncu --set full summary, if any."""
    p = os.path.join(ROOT, "profiles", "extend_dram_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


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
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1 for its workers)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
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
        "config": {"workload": "%s %dx%d, %d bounces, CPU sample of %d spp per step" % (args.workload, w, h, bounces, sample_spp)},
        "samples_per_s": w * h * sample_spp * args.steps / t,
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind,
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
                    help="N>1: 'samples' = frame f on rank f mod N (weak scaling, default); 'rows' = interleaved bands of 8 rows, "
                         "every rank renders all frames of its bands (strong scaling, bit-identical to one GPU)")
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
    build_first_ms = g.GetBVHBuildMilliseconds()  # includes the first-use costs: module load, growth of the memory pool
    g.LoadScene(scene_arg(spec))                  # the same build again: the steady-state builder time
    g.Resize(w, h)
    if args.fif:
        g.SetFramesInFlight(args.fif)
    if args.shard == "rows":
        g.SetRowShard(rank, world)    # bands of 8 rows, band b on rank b mod N
    else:
        g.SetFrameShard(rank, world)  # frame f rendered on rank f mod N
    s = tb.get_default_output_settings()
    s.MaxBounces = bounces
    info = g.GetSceneInfo()

    acc_ptr, acc_bytes = g.DeviceBuffer(tb.BufferKind.ACCUM_RGBW)

    class _Holder:  # zero-copy torch view of the library's float4 accumulation buffer
        __cuda_array_interface__ = {"shape": (h, w, 4), "typestr": "<f4", "data": (acc_ptr, False), "version": 2}
    acc_t = torch.as_tensor(_Holder(), device=torch.device("cuda", local_rank))
    reduced = torch.empty_like(acc_t) if world > 1 else None

    # pinned host buffers for the e2e leg
    host_img = torch.empty((h, w, 3), dtype=torch.float32).pin_memory()
    host_in = torch.empty(ctypes.sizeof(s) + ctypes.sizeof(tb.Camera), dtype=torch.uint8).pin_memory()
    cam = g.GetCamera()

    def step_resident():
        g.InvalidateHistory()
        g.Render(s, spp, 0.0)
        if world > 1:  # the path's one exchange step: sum of the per-rank accumulation buffers
            reduced.copy_(acc_t)
            dist.all_reduce(reduced)

    def step_e2e():
        # host -> device: this step's inputs (PerFrameConstants sources) from pinned memory
        ctypes.memmove(host_in.data_ptr(), ctypes.addressof(s), ctypes.sizeof(s))
        ctypes.memmove(host_in.data_ptr() + ctypes.sizeof(s), ctypes.addressof(cam), ctypes.sizeof(cam))
        s_in = tb.OutputSettings.from_address(host_in.data_ptr())
        cam_in = tb.Camera.from_address(host_in.data_ptr() + ctypes.sizeof(s))
        g.SetCamera(cam_in)  # also invalidates history, like TracerBoy::Update
        g.Render(s_in, spp, 0.0)
        if world > 1:
            reduced.copy_(acc_t)
            dist.all_reduce(reduced)
        # device -> host: the resolved image (rgb / w), read by the caller
        g.Readback(tb.BufferKind.RESOLVED_RGB, out=host_img.numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g.ResetRenderStats()
        g.Synchronize()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        g.Synchronize()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        barrier()
        st = g.GetRenderStats()
        # device time of the library's own stream (CUDA events recorded by tb_render) ...
        dev_ms = st.DeviceMilliseconds
        # ... and the wall clock bracketed by synchronisations (includes copies / collectives)
        return st, dev_ms, wall

    for _ in range(max(3, args.warmup)):
        step_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    st, dev_ms, wall = timed(step_resident, args.steps)
    clocks = sampler.stop()
    # value: whole-step time on the device (max over ranks); tb_render's events bracket the
    # kernels, the wall clock additionally covers the all-reduce when N > 1
    t_step = wall if world > 1 else dev_ms / 1e3
    vals = torch.tensor([t_step, float(st.RaysTraced), float(st.KernelLaunches)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = vals.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = vals.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        t_step, rays_total, launches = tmax[0].item(), tsum[1].item(), tsum[2].item()
    else:
        rays_total, launches = float(st.RaysTraced), float(st.KernelLaunches)
    value = rays_total / t_step / 1e6

    # e2e leg
    for _ in range(2):
        step_e2e()
    st_e, _, wall_e = timed(step_e2e, args.steps)
    ve = torch.tensor([wall_e, float(st_e.RaysTraced)], dtype=torch.float64, device="cuda")
    if world > 1:
        a = ve.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b = ve.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
        wall_e, rays_e = a[0].item(), b[1].item()
    else:
        rays_e = float(st_e.RaysTraced)
    e2e_value = rays_e / wall_e / 1e6

    # roofline of the dominant kernel, measured live with per-launch CUDA events (profiling mode)
    g.SetProfiling(True)
    g.ResetRenderStats()
    g.InvalidateHistory()
    g.Render(s, min(spp, 16), 0.0)
    pst = g.GetRenderStats()
    g.SetProfiling(False)
    # rays that finish inside k_extend (the rest are suspended and finish in k_extend_resume, timed separately)
    alg_bytes = 32.0 * pst.ExtendBoxesTested + 40.0 * pst.ExtendTrianglesTested + 64.0 * pst.ExtendRays
    peak, peak_src = measured_peak_gbs()
    ext_s = pst.ExtendMilliseconds / 1e3
    achieved = alg_bytes / ext_s / 1e9 if ext_s > 0 else 0.0
    # DRAM bytes per launch from the committed ncu --set full capture (Teapot only: the capture is of that workload).
    # The captured launch is a bounce-0 launch; scaled by rays to the average launch `achieved` is quoted on.
    traffic = ncu_traffic() if args.workload == "teapot" else None
    traffic_per_launch = None
    if traffic:
        cap = traffic.get("launches", [{}])[0]
        if cap.get("rays"):
            per_ray = (cap["dram_read"] + cap["dram_write"]) / cap["rays"]
            traffic_per_launch = per_ray * pst.ExtendRays / max(1, pst.ExtendLaunches)
        else:
            traffic_per_launch = traffic["dram_bytes_per_launch"]
    roofline = {
        "bound": "hbm", "kernel": "k_extend", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic_per_launch,
        "traffic_source": traffic["source"] if traffic else None,
        "algorithmic_bytes_per_launch": alg_bytes / max(1, pst.ExtendLaunches),
        "avg_launch_ms": pst.ExtendMilliseconds / max(1, pst.ExtendLaunches), "launches": pst.ExtendLaunches,
        "kernel_share_of_step": pst.ExtendMilliseconds / max(1e-9, pst.ExtendMilliseconds + pst.ShadeMilliseconds + pst.ResumeMilliseconds),
        "resume_rounds": {"rays": pst.ResumeRays, "ms": pst.ResumeMilliseconds,
                          "algorithmic_bytes": 32.0 * pst.ResumeBoxesTested + 40.0 * pst.ResumeTrianglesTested + 64.0 * pst.ResumeRays},
        # all rays of the timed region (every kernel, frames in flight overlapped) over the whole step time
        "whole_step_algorithmic_GBps": world * (32.0 * st.BoxesTested + 40.0 * st.TrianglesTested + 64.0 * st.RaysTraced) / t_step / 1e9,
        "peak_source": peak_src,
        "note": "algorithmic bytes = 32 B x BoxesTested + 40 B x TrianglesTested + 64 B ray/hit (reference BVH2 layout); "
                "a BVH that fits the 126 MB L2 is served from L2, so frac can exceed DRAM-only expectations",
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
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        o.Render(s, 1, 0.0, threads=cores)
        c0 = o.Counts()
        sec = o.Render(s, args.cpu_baseline_spp, 0.0, threads=cores)
        c1 = o.Counts()
        cpu_baseline = {"value": (c1["rays"] - c0["rays"]) / sec / 1e6, "unit": "Mrays/s", "cores": cores,
                        "kind": kind, "sample": "%d spp of the %dx%d %s workload (%.1f s), OpenMP over pixels" % (
                            args.cpu_baseline_spp, w, h, args.workload, sec)}

    if rank == 0:
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": 1e3 * t_step / args.steps, "higher_is_better": True,
            "scaling": "strong" if (args.shard == "rows" and world > 1) else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic camera/seeds on the bundled scene",
            "config": {"workload": "%s %dx%d, %d spp per step per GPU, %d bounces, NEE on, blue noise on, Time=0" % (
                           args.workload, w, h, spp, bounces),
                       "triangles": info.NumTriangles, "sharding": ("bands of 8 rows, band b on rank b mod N" if args.shard == "rows" else "frame f on rank f mod N") +
                                   "; all-reduce of the accumulation buffer per step",
                       "l2": "inputs exceed L2: %d MB of path state + accumulation buffers are rewritten every sample" % (
                           (w * h * 16 * 14) >> 20),
                       "frames_in_flight": args.fif if args.fif else "auto (memory budget, <= 32)"},
            "samples_per_s": (1 if args.shard == "rows" else world) * w * h * spp * args.steps / t_step,
            "rays_per_step": rays_total / args.steps,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(host_in.numel()),
                    "d2h_bytes_per_step": int(host_img.numel() * 4), "ms_per_step": 1e3 * wall_e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "bvh_build_ms": g.GetBVHBuildMilliseconds(), "bvh_build_first_call_ms": build_first_ms,
            "bvh_build_mtris_per_s": g.GetSceneInfo().NumTriangles / max(1e-9, g.GetBVHBuildMilliseconds()) / 1e3, "scene_load_s": load_s,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
