"""CUPTI kernel timeline (torch.profiler) of one bench step of the five-render pattern: S MergedFivePlan engines, F training
frames, one CUDA-graph replay.  Prints per-kernel durations under concurrency, the overlap histogram and the idle gaps.

  python tools/five_timeline.py --lanes 3 --frames 8 > gpurun_out/five_timeline.txt
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from exavatar_release_b200 import plan as PL  # noqa: E402
from exavatar_release_b200.camera import look_at_cam_param  # noqa: E402
from exavatar_release_b200.renderer import render_settings  # noqa: E402
from exavatar_release_b200.synthetic import WORKLOADS, make_grad_image, make_population_assets  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C4")
    ap.add_argument("--lanes", type=int, default=3)
    ap.add_argument("--frames", type=int, default=8)
    a = ap.parse_args()
    wl = WORKLOADS[a.workload]
    dev = torch.device("cuda:0")
    H, W, S, F = wl.height, wl.width, a.lanes, a.frames
    scene, human, refined = make_population_assets(a.workload, seed=0, device=dev)
    cams = [look_at_cam_param(-20.0 + 40.0 * (f % 8) / 7.0, (H, W), device=dev) for f in range(F)]
    st_w = [render_settings((H, W), c, torch.ones(3, device=dev)) for c in cams]
    st_r = [render_settings((H, W), c, torch.tensor([0.3, 0.7, 0.2], device=dev)) for c in cams]
    g5 = [{r: make_grad_image(a.workload, 10 * f + j, device=dev) for j, r in enumerate(PL.RENDERS)} for f in range(F)]
    engines = [PL.MergedFivePlan(wl.n_scene, wl.n_avatar, W, H, {"A": 900_000, "B": 900_000}, dev) for _ in range(S)]
    streams = [torch.cuda.Stream(dev) for _ in range(S)]

    def body():
        cur = torch.cuda.current_stream(dev)
        for s in range(S):
            streams[s].wait_stream(cur)
            with torch.cuda.stream(streams[s]):
                e = engines[s]
                e.set_scene(scene)
                for j, f in enumerate(range(s, F, S)):
                    e.frame(("f", f), st_w[f], st_r[f], scene, human, refined, g5[f], accumulate=(j > 0))
        for s in range(S):
            cur.wait_stream(streams[s])

    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        body()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    assert not any(e.overflowed() for e in engines)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        g.replay()
        torch.cuda.synchronize()
    path = os.path.join(tempfile.gettempdir(), "five_trace.json")
    prof.export_chrome_trace(path)
    ev = json.load(open(path))["traceEvents"]
    ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e]
    t0 = min(e["ts"] for e in ks)
    t1 = max(e["ts"] + e["dur"] for e in ks)
    span = t1 - t0
    print(f"step span {span:.1f} us = {F / span * 1e6:.0f} training frames/s, {len(ks)} GPU activities, lanes={S}, frames={F}")
    by = {}
    for e in ks:
        n = e["name"].split("(")[0].replace("void ", "").replace("b2r::", "")[:44]
        by.setdefault(n, []).append(e["dur"])
    print(f"{'kernel':44s} {'count':>6s} {'mean us':>9s} {'min':>7s} {'max':>7s} {'total':>9s}")
    for n, d in sorted(by.items(), key=lambda kv: -sum(kv[1])):
        print(f"{n:44s} {len(d):6d} {sum(d) / len(d):9.1f} {min(d):7.1f} {max(d):7.1f} {sum(d):9.1f}")
    tot = sum(e["dur"] for e in ks)
    print(f"sum of activity durations {tot:.0f} us; / span = {tot / span:.2f} average overlap")
    grid = np.linspace(t0, t1, 2001)[:-1]
    conc = np.zeros(len(grid), int)
    comp = np.zeros(len(grid), int)
    for e in ks:
        m = (grid >= e["ts"]) & (grid < e["ts"] + e["dur"])
        conc += m
        if "composite" in e["name"]:
            comp += m
    print("fraction of span with k activities running:", {int(k): round(float((conc == k).mean()), 3) for k in np.unique(conc)})
    print("fraction of span with k COMPOSITE kernels running:", {int(k): round(float((comp == k).mean()), 3) for k in np.unique(comp)})


if __name__ == "__main__":
    main()
