"""Kernel timeline of one bench step under FrameLanes (CUPTI through torch.profiler; no nsys in this image).

  python tools/lanes_timeline.py --workload C2 --lanes 4 > gpurun_out/lanes_timeline.txt
Prints per-kernel durations under concurrency, how many kernels overlap over time, and the critical chain of one lane.
"""
import argparse
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from exavatar_release_b200 import rasterizer as RZ  # noqa: E402
from exavatar_release_b200.camera import look_at_cam_param  # noqa: E402
from exavatar_release_b200.plan import FrameLanes  # noqa: E402
from exavatar_release_b200.renderer import render_settings  # noqa: E402
from exavatar_release_b200.synthetic import WORKLOADS, make_assets, make_grad_image  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--lanes", type=int, default=4)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--graph", action="store_true")
    a = ap.parse_args()
    wl = WORKLOADS[a.workload]
    dev = torch.device("cuda:0")
    P, H, W = wl.n_avatar + wl.n_scene, wl.height, wl.width
    M = (wl.sh_degree + 1) ** 2 if wl.sh_degree > 0 else 0
    assets = make_assets(a.workload, seed=0, device=dev)
    bg = torch.ones(3, device=dev)
    F = a.frames
    sts = []
    for f in range(F):
        st = render_settings((H, W), look_at_cam_param(-20.0 + 40.0 * (f % 8) / 7.0, (H, W), device=dev), bg)
        sts.append(st._replace(sh_degree=wl.sh_degree) if M else st)
    gis = [make_grad_image(a.workload, f, device=dev) for f in range(F)]
    lanes = FrameLanes(a.lanes, P, W, H, 40_000_000 if a.workload in ("C3", "C5") else 12_000_000, dev, sh_coeffs=M)
    scenes = [lanes.scene(f, sts[f], assets) for f in range(F)]

    def step():
        lanes.step(scenes, gis, backward=wl.backward)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    run = step
    if a.graph:
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        run = g.replay
        run()
        torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        run()
        torch.cuda.synchronize()
    path = os.path.join(tempfile.gettempdir(), "lanes_trace.json")
    prof.export_chrome_trace(path)
    ev = json.load(open(path))["traceEvents"]
    ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e]
    if not ks:
        print("no kernel events captured")
        return
    t0 = min(e["ts"] for e in ks)
    t1 = max(e["ts"] + e["dur"] for e in ks)
    print(f"step span {t1 - t0:.1f} us, {len(ks)} GPU activities, graph={a.graph}, lanes={a.lanes}, frames={F}")
    by = {}
    for e in ks:
        n = e["name"].split("(")[0].replace("void ", "").replace("b2r::", "")[:40]
        by.setdefault(n, []).append(e["dur"])
    print("kernel                                    count   mean us    min    max   total")
    for n, d in sorted(by.items(), key=lambda kv: -sum(kv[1])):
        print(f"{n:40s} {len(d):6d} {sum(d) / len(d):9.1f} {min(d):6.1f} {max(d):6.1f} {sum(d):8.1f}")
    print("sum of kernel durations", round(sum(e["dur"] for e in ks), 1), "us; / span =",
          round(sum(e["dur"] for e in ks) / (t1 - t0), 2), "average overlap")
    # concurrency histogram
    import numpy as np
    grid = np.linspace(t0, t1, 401)[:-1]
    conc = np.zeros(len(grid), int)
    for e in ks:
        conc += (grid >= e["ts"]) & (grid < e["ts"] + e["dur"])
    print("fraction of span with k kernels running:", {int(k): round(float((conc == k).mean()), 3) for k in np.unique(conc)})
    # one lane's chain
    streams = sorted({e["args"].get("stream") for e in ks if "args" in e})
    print("streams:", streams)
    for s in streams[:2]:
        ch = sorted([e for e in ks if e.get("args", {}).get("stream") == s], key=lambda e: e["ts"])
        print(f"-- stream {s}: first 24 activities (start us, dur us, gap before)")
        prev = None
        for e in ch[:24]:
            gap = (e["ts"] - prev) if prev is not None else 0.0
            print(f"   {e['ts'] - t0:8.1f} {e['dur']:7.1f} {gap:7.1f}  {e['name'][:60]}")
            prev = e["ts"] + e["dur"]


if __name__ == "__main__":
    main()
