#!/usr/bin/env python
"""bench.py -- avatar train-step frames/s (forward + backward Gaussian rasterisation), BASELINE.json's metric.

One "step" = every rank rasterises F frames of the workload (forward + backward; `--lanes` of them in flight on
separate CUDA streams, gradients of the frames summed into one per-rank bucket) followed, when N > 1, by ONE NCCL
all-reduce of that bucket.  Frames are independent, so ranks share nothing else: weak scaling.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2] [--frames F] [--lanes S] [--impl b200|reference]

Legs (all in one JSON line, printed by rank 0):
  value     device-resident: inputs already in HBM, C ABI driven through FrameLanes, the step captured in a CUDA graph
  e2e       the public plugin API (GaussianRenderer -> GaussianRasterizer autograd, L1 loss, backward) with HOST buffers,
            copies inside the timed region, the whole step in one CUDA graph.  --e2e-upload per-step (default): the
            Gaussian parameter set goes host->device once per step (all frames of a step render it) plus one target
            image per frame; the step's summed gradients and per-frame losses come back.  per-frame: every rasteriser
            call uploads its full inputs and downloads its own gradients.
  roofline  dominant kernel's algorithmic bytes / its live CUDA-event duration (in-library profiler, frames one at a time)
  cpu_baseline  the CPU oracle (oracle/, "port") timed on the host cores on a bounded sample of the same workload (N = 1)
  five_render   (--five-render) ExAvatar's five renders per training frame on FiveRenderPlan
`--impl reference` times that CPU oracle through the same GaussianRenderer call as the reference arm.

Timing: W >= 3 warm-up steps; L2 is flushed (256 MiB memset) before every timed step, outside the per-step CUDA
event pairs; per-rank time = sum of per-step event durations; max over ranks.  SM clocks / throttle reasons: NVML polled
by a thread during the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from exavatar_release_b200.camera import look_at_cam_param  # noqa: E402
from exavatar_release_b200.synthetic import WORKLOADS, make_assets, make_grad_image  # noqa: E402

METRIC = "avatar train-step frames/sec (fwd+bwd raster)"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=[k for k in WORKLOADS if k.startswith("C")])
    ap.add_argument("--frames", type=int, default=8, help="frames per rank per step")
    ap.add_argument("--lanes", type=int, default=4, help="frames in flight per rank (CUDA streams; 1 = serial)")
    ap.add_argument("--no-graph", action="store_true", help="do not capture the step in a CUDA graph")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-upload", default="per-step", choices=["per-step", "per-frame"],
                    help="e2e leg: copy the Gaussian set host->device once per step (the frames of a step share it, as in "
                         "the device-resident leg; result = losses + the step's summed gradients) or once per frame "
                         "(every rasteriser call gets fresh host inputs and returns its own gradients)")
    ap.add_argument("--five-render", action="store_true",
                    help="extra leg: ExAvatar's five-render training frame (model.py:81-162) on FiveRenderPlan; needs a "
                         "workload with both populations (C2, C4)")
    ap.add_argument("--trace-e2e", default=None, help="write a chrome trace (CUPTI via torch.profiler) of one e2e step here")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU budget of the cpu_baseline leg")
    return ap.parse_args()


def frame_yaw(global_frame: int) -> float:
    return -20.0 + 40.0 * ((global_frame % 8) / 7.0)  # 8 distinct cameras, yaw +-20 deg


def config_dict(args, wl, extra=None):
    c = {"workload": wl.name, "frames_per_rank_per_step": args.frames, "P": wl.n_avatar + wl.n_scene,
         "image": f"{wl.width}x{wl.height}", "sh_degree": wl.sh_degree, "backward": wl.backward,
         "parallelism": f"frames sharded over {args.gpus} rank(s), one gradient all-reduce per step" if args.gpus > 1
         else "single GPU", "l2": "256 MiB L2 flush before every timed step (outside the event pairs)"}
    if extra:
        c.update(extra)
    return c


# ---------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle behind the reference-facing call
# ---------------------------------------------------------------------------------------------------------------
def cpu_frame_fn(wl_name, seed=0):
    """Returns a closure running one frame (fwd [+bwd]) of the workload on the CPU oracle through GaussianRenderer."""
    from oracle import oracle as O
    from exavatar_release_b200.renderer import GaussianRenderer, render_settings

    wl = WORKLOADS[wl_name]
    O.set_num_threads(os.cpu_count() or 1)
    assets = make_assets(wl_name, seed=seed)
    gi = make_grad_image(wl_name, seed)
    renderer = GaussianRenderer(rasterizer_cls=O.OracleRasterizer, settings_cls=O.OracleSettings)
    bg = torch.ones(3)
    use_sh = wl.sh_degree > 0

    def frame(i):
        cam = look_at_cam_param(frame_yaw(i), (wl.height, wl.width))
        leaves = {k: v.clone().requires_grad_(wl.backward) for k, v in assets.items()}
        if use_sh:  # C3: colours from SH inside the rasteriser
            st = render_settings((wl.height, wl.width), cam, bg, O.OracleSettings)._replace(sh_degree=wl.sh_degree)
            m2 = torch.zeros(leaves["mean_3d"].shape[0], 3, requires_grad=wl.backward)
            img = O.OracleRasterizer(st)(means3D=leaves["mean_3d"], means2D=m2, opacities=leaves["opacity"],
                                         shs=leaves["shs"], scales=leaves["scale"], rotations=leaves["rotation"])[0]
        else:
            img = renderer(leaves, (wl.height, wl.width), cam, bg)["img"]
        if wl.backward:
            (img * gi).sum().backward()
        return float(img.detach().sum())

    # "all the host threads it can use": OpenMP scaling of the oracle saturates (atomics in the backward), so pick the
    # fastest thread count among a few candidates instead of blindly using every core
    ncpu = os.cpu_count() or 1
    best, best_t = ncpu, None
    for cand in sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu}):
        O.set_num_threads(cand)
        t0 = time.perf_counter()
        frame(0)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = cand, dt
    O.set_num_threads(best)
    return frame, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = WORKLOADS[args.workload]
    frame, threads = cpu_frame_fn(args.workload)
    # bounded sample: one frame per step
    for i in range(args.warmup):
        frame(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        frame(i)
    dt = time.perf_counter() - t0
    fps = args.steps / dt
    sample = f"1 frame of {wl.name} per step ({'fwd+bwd' if wl.backward else 'fwd'}), {args.steps} steps, {threads} threads"
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, wl, {"frames_per_rank_per_step": 1, "parallelism": f"{threads} host threads (OpenMP)",
                                              "l2": "n/a (CPU)"}),
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  The region lasts tens of milliseconds, far too
    short for `nvidia-smi -lms`, so a thread polls NVML directly (~1 kHz); nvidia-smi is the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        import threading
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.window = "timed region"
        self._stop = threading.Event()
        self._thread = None
        self.proc, self.path = None, None
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            bits = {"hw_slowdown": getattr(N, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(N, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(N, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(N, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}

            def poll():
                while not self._stop.is_set():
                    try:
                        self.samples.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
                        r = int(N.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        for name, bit in bits.items():
                            if r & bit:
                                self.reasons.add(name)
                    except Exception:
                        pass
                    time.sleep(0.001)

            self._thread = threading.Thread(target=poll, daemon=True)
            self._thread.start()
            return
        except Exception:
            self._thread = None
        exe = shutil.which("nvidia-smi")
        if exe is None:
            return
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        self.f = open(self.path, "w")
        self.proc = subprocess.Popen([exe, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                      "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
            if not self.samples:
                return None
            sm = sorted(self.samples)
            return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                    "samples": len(sm), "source": "nvml thread", "window": self.window}
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        rows = [r.strip().split(", ") for r in open(self.path) if r.strip()]
        os.unlink(self.path)
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi -lms 20"}


# ---------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    from exavatar_release_b200 import _lib as L
    from exavatar_release_b200 import rasterizer as RZ
    from exavatar_release_b200.plan import FrameLanes
    from exavatar_release_b200.renderer import GaussianRenderer, render_settings

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = L.load()
    wl = WORKLOADS[args.workload]
    F, K, Wm = args.frames, args.steps, max(args.warmup, 3)
    P, H, Wd = wl.n_avatar + wl.n_scene, wl.height, wl.width
    N = H * Wd
    use_sh = wl.sh_degree > 0
    M = (wl.sh_degree + 1) ** 2 if use_sh else 0
    bg = torch.ones(3, device=dev)

    # one Gaussian set per rank (the replicated parameters), F cameras per step
    assets = make_assets(args.workload, seed=0, device=dev)
    gimgs = [make_grad_image(args.workload, seed=f, device=dev) for f in range(F)]
    cams = [look_at_cam_param(frame_yaw(rank * F + f), (H, Wd), device=dev) for f in range(F)]
    settings = []
    for c in cams:
        st = render_settings((H, Wd), c, bg)
        settings.append(st._replace(sh_degree=wl.sh_degree) if use_sh else st)

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # ---- capacity: learn the duplicate counts through the autograd front-end (exact mode) ----
    def public_frame(f, leaves=None, grad=True):
        lv = leaves or {k: v.detach().requires_grad_(wl.backward and grad) for k, v in assets.items()}
        m2 = torch.zeros(P, 3, device=dev, requires_grad=wl.backward and grad)
        rast = RZ.GaussianRasterizer(settings[f])
        img = rast(means3D=lv["mean_3d"], means2D=m2, opacities=lv["opacity"], shs=lv["shs"] if use_sh else None,
                   colors_precomp=None if use_sh else lv["rgb"], scales=lv["scale"], rotations=lv["rotation"])[0]
        return img, lv, m2

    dups = []
    with torch.no_grad():
        for f in range(F):
            img, _, _ = public_frame(f, grad=False)
            dups.append(RZ._state(dev).predicted[(P, Wd, H)])
    cap = int(max(dups) * 1.1) + 4096

    S = max(1, min(args.lanes, F))
    lanes = FrameLanes(S, P, Wd, H, cap, dev, sh_coeffs=M)
    plan = lanes.plans[0]
    scenes = [lanes.scene(f, settings[f], assets) for f in range(F)]
    bucket, views = lanes.bucket, lanes.lane_views[0]

    def step_body():
        lanes.step(scenes, gimgs, backward=wl.backward)

    graph = None
    if not args.no_graph:
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            step_body()  # warm caches / attribute calls before capture
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        l0 = lib.b2r_launch_count()
        with torch.cuda.graph(graph):
            step_body()
        launches_per_step = lib.b2r_launch_count() - l0  # kernels of this library recorded into the graph

    def step():
        if graph is not None:
            graph.replay()
        else:
            step_body()
        if world > 1 and wl.backward:
            dist.all_reduce(bucket)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def timed(fn, steps, count_launches=False):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        l0 = lib.b2r_launch_count()
        t0 = time.perf_counter()
        for s in range(steps):
            flush_buf.zero_()
            ev[s][0].record()
            fn()
            ev[s][1].record()
        barrier()
        wall = time.perf_counter() - t0
        ms = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), wall, lib.b2r_launch_count() - l0

    # ---- leg 1: device-resident value ----
    for _ in range(Wm):
        step()
    clocks = ClockSampler(local) if rank == 0 else None
    ms_total, wall, launches_eager = timed(step, K)
    if clocks is not None and len(clocks.samples) < 20:
        # The timed region lasts tens of milliseconds and one NVML query takes a few: keep the identical step running
        # (untimed) until the sampler has seen enough of this load
        t_end = time.perf_counter() + 0.5
        while time.perf_counter() < t_end and len(clocks.samples) < 40:
            for _ in range(4):
                step()
            torch.cuda.synchronize(dev)
        clocks.window = "timed region, then the same step replayed untimed until >= 20 NVML samples were taken"
    clk = clocks.stop() if clocks else None
    st_last = lanes.status()
    if st_last["overflow"]:
        raise SystemExit("bench.py: duplicate capacity overflowed; results invalid")
    launches = launches_eager if graph is None else K * launches_per_step
    fps = world * F * K / (ms_total * 1e-3)

    # ---- leg 2: per-kernel durations (same step, eager, in-library events) ----
    lib.b2r_profile_enable(1)
    import ctypes as C
    ms_arr = (C.c_double * 9)()
    cnt_arr = (C.c_uint64 * 9)()
    lib.b2r_profile_read(ms_arr, cnt_arr, 1)
    cons_f, cons_b, ndups = [], [], []
    prof_steps = min(K, 5)
    for s in range(prof_steps):
        flush_buf.zero_()
        for f in range(F):
            plan.forward(scenes[f])
            if wl.backward:
                plan.backward(scenes[f], gimgs[f], views, accumulate=(f > 0))
            if s == 0:
                stt = plan.status()
                cons_f.append(stt["consumed_fwd"]); cons_b.append(stt["consumed_bwd"]); ndups.append(stt["num_dups"])
    lib.b2r_profile_read(ms_arr, cnt_arr, 1)
    lib.b2r_profile_enable(0)
    per_kernel = {lib.b2r_kernel_name(i).decode(): {"ms_avg": (ms_arr[i] / cnt_arr[i]) if cnt_arr[i] else 0.0,
                                                    "launches": int(cnt_arr[i])} for i in range(9)}
    tiles = ((Wd + 15) // 16) * ((H + 15) // 16)
    # the composites run four quarter-tile CTAs per tile, each counting what it staged: /4 = per-tile list entries
    Cf = sum(cons_f) / len(cons_f) / 4.0
    Cb = sum(cons_b) / len(cons_b) / 4.0 if wl.backward else 0.0
    # algorithmic bytes per launch (SURVEY section 8d / BASELINE.md section 4)
    algo = {"composite_fwd": 44.0 * Cf + 24.0 * N + 8.0 * tiles, "composite_bwd": 84.0 * Cb + 20.0 * N}
    dom = max(("composite_fwd", "composite_bwd"), key=lambda k: per_kernel[k]["ms_avg"])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ach = algo[dom] / (per_kernel[dom]["ms_avg"] * 1e-3) / 1e9 if per_kernel[dom]["ms_avg"] > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload, {}).get(dom)
    except Exception:
        pass
    frame_kernel_ms = sum(v["ms_avg"] * v["launches"] for v in per_kernel.values()) / (prof_steps * F)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak if peak else None, "traffic": traffic,
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_launch": algo[dom], "kernel_ms_avg": per_kernel[dom]["ms_avg"],
                "binding_bound": "instruction issue (ALU/SFU/shared memory), not HBM -- see DESIGN.md section 5",
                "per_kernel_ms": {k: round(v["ms_avg"], 5) for k, v in per_kernel.items() if v["launches"]},
                "sum_kernel_ms_per_frame": frame_kernel_ms,
                "consumed_fwd_per_frame": Cf, "consumed_bwd_per_frame": Cb, "dups_per_frame": sum(ndups) / len(ndups)}

    # ---- leg 3: end to end through the public API with host buffers ----
    e2e = None
    if not args.no_e2e:
        host_assets = {k: v.cpu().pin_memory() for k, v in assets.items() if (k != "rgb" or not use_sh)}
        if use_sh:
            host_assets.pop("rgb", None)
        host_targets = [torch.rand(3, H, Wd).pin_memory() for _ in range(F)]
        host_grads = {k: torch.empty_like(v).pin_memory() for k, v in host_assets.items()}
        host_m2 = torch.empty(P, 3).pin_memory()
        host_loss = torch.empty(F).pin_memory()
        renderer = GaussianRenderer()
        h2d = F * (sum(v.numel() * 4 for v in host_assets.values()) + 3 * N * 4)
        d2h = F * ((sum(v.numel() * 4 for v in host_grads.values()) + P * 12 + 4) if wl.backward else 3 * N * 4)
        host_img = torch.empty(3, H, Wd).pin_memory()

        # The user's whole step -- pinned H2D of every frame's inputs, GaussianRenderer forward, loss, autograd backward,
        # D2H of loss + gradients -- is captured once in a CUDA graph through the PUBLIC API (fixed-capacity mode of the
        # rasteriser: no polling, see rasterizer.set_fixed_capacity) and replayed per step; inside the graph the
        # copies of frame f+1 / f-1 run on forked streams beside frame f's kernels.  Host cost per step: one launch.
        RZ.set_fixed_capacity(cap)
        h2d_s, d2h_s = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        lane_s = [torch.cuda.Stream(dev) for _ in range(S)] if S > 1 else None
        keep_alive = []  # nothing allocated inside the capture may be recycled across the forked streams

        def upload(f, cur):
            h2d_s.wait_stream(cur)
            with torch.cuda.stream(h2d_s):
                lv = {k: v.to(dev, non_blocking=True) for k, v in host_assets.items()}
                tgt = host_targets[f].to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(h2d_s)
            keep_alive.append((lv, tgt))
            return lv, tgt, ev

        def e2e_frame(f, lv, tgt):
            lv = {k: v.requires_grad_(wl.backward) for k, v in lv.items()}
            if use_sh:
                img, _, m2 = public_frame(f, leaves=lv)
            else:
                o = renderer(lv, (H, Wd), cams[f], bg, raster_settings=settings[f])
                img, m2 = o["img"], o["mean_2d"]
            if wl.backward:
                loss = (img - tgt).abs().mean()
                loss.backward()
                outs = [(host_loss[f:f + 1], loss.detach().reshape(1)), (host_m2, m2.grad)]
                outs += [(host_grads[k], lv[k].grad) for k in host_grads]
            else:
                outs = [(host_img, img.detach())]
            keep_alive.append((lv, img, m2, outs))
            return outs

        def e2e_body_per_frame():
            cur = torch.cuda.current_stream(dev)
            if lane_s:
                for ls in lane_s:
                    ls.wait_stream(cur)
            d2h_s.wait_stream(cur)
            nxt = upload(0, cur)
            for f in range(F):
                lv, tgt, ev = nxt
                if f + 1 < F:
                    nxt = upload(f + 1, cur)
                fs = lane_s[f % S] if lane_s else cur  # frame f runs on lane f mod S
                fs.wait_event(ev)
                with torch.cuda.stream(fs):
                    outs = e2e_frame(f, lv, tgt)
                    done = torch.cuda.Event()
                    done.record(fs)
                with torch.cuda.stream(d2h_s):
                    d2h_s.wait_event(done)
                    for dst, src in outs:
                        dst.copy_(src, non_blocking=True)
            cur.wait_stream(d2h_s)  # the step ends when its last result is in host memory (joins the forked streams)
            cur.wait_stream(h2d_s)
            if lane_s:
                for ls in lane_s:
                    cur.wait_stream(ls)

        def e2e_body_per_step():
            """One training step as ExAvatar runs it (train.py:35-57): the frames of the batch render the SAME parameter
            set, the loss gradients of all frames are summed into the parameters' .grad, the optimiser would read those.
            Host -> device: the parameter set once, one target image per frame.  Device -> host: the summed gradients
            and the per-frame losses."""
            cur = torch.cuda.current_stream(dev)
            streams = lane_s if lane_s else [cur]
            for st_ in (lane_s or []):
                st_.wait_stream(cur)
            d2h_s.wait_stream(cur)
            h2d_s.wait_stream(cur)
            with torch.cuda.stream(h2d_s):
                params = {k: v.to(dev, non_blocking=True) for k, v in host_assets.items()}
                ev_p = torch.cuda.Event()
                ev_p.record(h2d_s)
                tgts, ev_t = [], []
                for f in range(F):
                    tgts.append(host_targets[f].to(dev, non_blocking=True))
                    e = torch.cuda.Event()
                    e.record(h2d_s)
                    ev_t.append(e)
            keep_alive.append((params, tgts))
            leaves = []
            for st_ in streams:  # lane-private leaf views of the one uploaded parameter set
                st_.wait_event(ev_p)
                with torch.cuda.stream(st_):
                    leaves.append({k: v.detach().requires_grad_(wl.backward) for k, v in params.items()})
            losses = []
            for f in range(F):
                fs = streams[f % len(streams)]
                lv = leaves[f % len(streams)]
                with torch.cuda.stream(fs):
                    if use_sh:
                        img, _, m2 = public_frame(f, leaves=lv)
                    else:
                        o = renderer(lv, (H, Wd), cams[f], bg, raster_settings=settings[f])
                        img, m2 = o["img"], o["mean_2d"]
                    if wl.backward:
                        fs.wait_event(ev_t[f])  # the target image is only needed here: its upload overlaps the forward
                        loss = torch.nn.functional.l1_loss(img, tgts[f])
                        loss.backward()  # accumulates into this lane's leaves
                        losses.append(loss.detach().reshape(1))
                        keep_alive.append((img, m2, loss))
                    else:
                        done = torch.cuda.Event()
                        done.record(fs)
                        keep_alive.append((img, m2))
                        with torch.cuda.stream(d2h_s):
                            d2h_s.wait_event(done)
                            host_img.copy_(img.detach(), non_blocking=True)
            for st_ in (lane_s or []):
                cur.wait_stream(st_)
            if wl.backward:
                total = {k: leaves[0][k].grad for k in host_grads}
                for lv in leaves[1:]:
                    if lv[next(iter(host_grads))].grad is not None:
                        total = {k: total[k] + lv[k].grad for k in host_grads}
                lvec = torch.cat(losses)
                keep_alive.append((leaves, total, lvec))
                done = torch.cuda.Event()
                done.record(cur)
                with torch.cuda.stream(d2h_s):
                    d2h_s.wait_event(done)
                    host_loss.copy_(lvec, non_blocking=True)
                    for k in host_grads:
                        host_grads[k].copy_(total[k], non_blocking=True)
            cur.wait_stream(d2h_s)
            cur.wait_stream(h2d_s)

        per_step = args.e2e_upload == "per-step"
        e2e_body = e2e_body_per_step if per_step else e2e_body_per_frame
        if per_step:
            h2d = sum(v.numel() * 4 for v in host_assets.values()) + F * 3 * N * 4
            d2h = (sum(v.numel() * 4 for v in host_grads.values()) + F * 4) if wl.backward else F * 3 * N * 4

        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                e2e_body()
                torch.cuda.synchronize(dev)  # buffers cross streams: do not recycle them while a lane may still read
                keep_alive.clear()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        e2e_graph = None
        if not args.no_graph:
            try:
                e2e_graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(e2e_graph):
                    e2e_body()
            except Exception as exc:  # report, fall back to eager replay of the same body
                print(f"bench.py: e2e graph capture failed ({type(exc).__name__}: {exc}); timing the eager step", file=sys.stderr)
                e2e_graph = None
                torch.cuda.synchronize(dev)

        def e2e_step():
            if e2e_graph is not None:
                e2e_graph.replay()
            else:
                e2e_body()
                torch.cuda.synchronize(dev)
                keep_alive.clear()
            if world > 1 and wl.backward:
                dist.all_reduce(bucket)  # same collective as the device-resident leg

        for _ in range(3):
            e2e_step()
        ke = max(3, min(K, 20))
        ms_e, _, _ = timed(e2e_step, ke)
        if args.trace_e2e and rank == 0:  # after the timed region: one more step under the profiler
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                e2e_step()
                torch.cuda.synchronize(dev)
            prof.export_chrome_trace(args.trace_e2e)
        if RZ.overflowed():
            raise SystemExit("bench.py: e2e leg overflowed its fixed duplicate capacity; results invalid")
        RZ.set_fixed_capacity(None)
        e2e = {"value": world * F * ke / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "upload": args.e2e_upload + (": parameter set once per step + one target image per frame up; summed "
                                            "gradients + per-frame losses down" if per_step else
                                            ": every frame uploads the full Gaussian set + target and downloads its gradients"),
               "api": "GaussianRenderer.forward -> GaussianRasterizer (autograd) + L1 loss + backward, pinned host buffers; "
                      + ("whole step captured in a CUDA graph, copies on forked streams" if e2e_graph is not None
                         else "eager, copies on side streams"),
               "steps": ke}

    # ---- optional leg: the five-render training frame of avatar/main/model.py:81-162 ----
    five = None
    if args.five_render and wl.backward and wl.n_avatar and wl.n_scene and not use_sh:
        from exavatar_release_b200.plan import RENDERS, FiveRenderPlan
        from exavatar_release_b200.synthetic import make_population_assets
        scene_a, human_a, refined_a = make_population_assets(args.workload, seed=0, device=dev)
        bg_rand = torch.tensor([0.3, 0.7, 0.2], device=dev)
        st_w = [render_settings((H, Wd), c, bg) for c in cams]
        st_r = [render_settings((H, Wd), c, bg_rand) for c in cams]
        g5 = [{r: make_grad_image(args.workload, seed=10 * f + j, device=dev) for j, r in enumerate(RENDERS)} for f in range(F)]
        probe = FiveRenderPlan(wl.n_scene, wl.n_avatar, Wd, H, {r: 8_000_000 for r in RENDERS}, dev)
        probe.set_scene(scene_a)
        need = {r: 0 for r in RENDERS}
        for f in range(F):
            probe.frame(f, st_w[f], st_r[f], scene_a, human_a, refined_a, g5[f], accumulate=False)
            torch.cuda.synchronize(dev)
            for r in RENDERS:
                need[r] = max(need[r], probe.plans[r].status()["num_dups"])
        del probe
        fplan = FiveRenderPlan(wl.n_scene, wl.n_avatar, Wd, H, {r: int(need[r] * 1.1) + 4096 for r in RENDERS}, dev)

        def five_body():
            fplan.set_scene(scene_a)
            for f in range(F):
                fplan.frame(f, st_w[f], st_r[f], scene_a, human_a, refined_a, g5[f], accumulate=(f > 0))
            return fplan.reduce()

        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            five_body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g5graph = torch.cuda.CUDAGraph()
        l0 = lib.b2r_launch_count()
        with torch.cuda.graph(g5graph):
            five_out = five_body()
        five_launches = lib.b2r_launch_count() - l0
        for _ in range(3):
            g5graph.replay()
        k5 = max(3, min(K, 10))
        ms5, _, _ = timed(g5graph.replay, k5)
        if fplan.overflowed():
            raise SystemExit("bench.py: five-render leg overflowed its duplicate capacity")
        five = {"value": world * F * k5 / (ms5 * 1e-3), "unit": "frames/s (5 renders, fwd+bwd, per frame)",
                "renders_per_s": 5 * world * F * k5 / (ms5 * 1e-3), "launches_per_step": int(five_launches),
                "P_scene": wl.n_scene, "P_human": wl.n_avatar, "dups_per_render": need,
                "pattern": "scene | human (random bg) | cat(scene.detach(), human) | human_refined | "
                           "cat(scene.detach(), human_refined); five streams per frame, one CUDA graph per step"}

    # ---- leg 4: CPU baseline on the host cores (rank 0) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # reported at N = 1 only (the other ranks would idle)
        frame, threads = cpu_frame_fn(args.workload)
        frame(0)
        n, t0 = 0, time.perf_counter()
        while True:
            frame(n); n += 1
            if time.perf_counter() - t0 > args.cpu_seconds or n >= 64:
                break
        dt = time.perf_counter() - t0
        cpu = {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{n} frame(s) of {wl.name} on the CPU oracle (OpenMP, {threads} threads), {dt:.1f} s"}

    if rank == 0:
        line = {"metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": config_dict(args, wl, {"cuda_graph": graph is not None, "dup_capacity": cap, "lanes": S,
                                                 "frames_per_rank_per_step": F}),
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "five_render": five,
                "wall_s_timed_region": wall}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
