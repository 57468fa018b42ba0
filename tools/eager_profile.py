"""cProfile of the eager drop-in path: N training frames of the five-call pattern (model.py:117-162) through
GaussianRenderer -> GaussianRasterizer (autograd), forward + backward, no CUDA graph.  Shows where the host time goes.

  python tools/eager_profile.py [--frames 40] [--fused [--graph]]
"""
import argparse
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from exavatar_release_b200 import GaussianRenderer  # noqa: E402
from exavatar_release_b200.camera import look_at_cam_param  # noqa: E402
from exavatar_release_b200.synthetic import WORKLOADS, make_population_assets  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C4")
    ap.add_argument("--frames", type=int, default=40)
    ap.add_argument("--fused", action="store_true", help="TrainingFrameRenderer instead of five GaussianRenderer calls")
    ap.add_argument("--presettings", action="store_true",
                    help="with --fused: raster settings built ahead of the loop (a data-loader worker), so the frame has no "
                         "camera D2H synchronisation")
    ap.add_argument("--graph", action="store_true", help="with --fused: TrainingFrameRenderer(use_graph=True)")
    ap.add_argument("--single", action="store_true", help="one render per frame of the whole workload (e.g. --workload C2)")
    ap.add_argument("--same-camera", action="store_true", help="reuse one camera tensor set (the per-camera cache hits)")
    a = ap.parse_args()
    wl = WORKLOADS[a.workload]
    dev = torch.device("cuda:0")
    H, W = wl.height, wl.width
    if a.single:
        from exavatar_release_b200.synthetic import make_assets
        whole = make_assets(a.workload, seed=0, device=dev)
        scene = human = refined = whole
    else:
        scene, human, refined = make_population_assets(a.workload, seed=0, device=dev)
    cam0 = look_at_cam_param(-20.0, (H, W), device=dev)
    bg_r = torch.tensor([0.3, 0.7, 0.2], device=dev)
    tgt = torch.rand(3, H, W, device=dev)
    R = GaussianRenderer()
    fused = None
    if a.fused:
        from exavatar_release_b200 import TrainingFrameRenderer
        fused = TrainingFrameRenderer(wl.n_scene, wl.n_avatar, (H, W), dev, {"A": 900_000, "B": 900_000}, use_graph=a.graph)
    cat = lambda x, y: {k: torch.cat((x[k].detach(), y[k])) for k in x}

    pre = None
    if a.presettings:
        from exavatar_release_b200.rasterizer import GaussianRasterizationSettings
        from exavatar_release_b200.renderer import render_settings
        white = torch.ones(3, device=dev)
        pre = [render_settings((H, W), look_at_cam_param(-20.0 + j * 5.0, (H, W), device=dev), white,
                               GaussianRasterizationSettings) for j in range(8)]

    def frame(i):
        cam = cam0 if a.same_camera else look_at_cam_param(-20.0 + (i % 8) * 5.0, (H, W), device=dev)  # new tensors per frame
        if a.single:
            lv1 = {k: v.detach().requires_grad_() for k, v in whole.items()}
            torch.nn.functional.l1_loss(R(lv1, (H, W), cam)["img"], tgt).backward()
            return
        lv = {n: {k: v.detach().requires_grad_() for k, v in s.items()} for n, s in
              (("scene", scene), ("human", human), ("refined", refined))}
        if fused is not None:
            out = fused(lv["scene"], lv["human"], lv["refined"], cam, bg_r,
                        raster_settings=None if pre is None else pre[i % 8])
            imgs = [out[r]["img"] for r in out]
        else:
            imgs = [R(lv["scene"], (H, W), cam)["img"], R(lv["human"], (H, W), cam, bg_r)["img"],
                    R(cat(lv["scene"], lv["human"]), (H, W), cam)["img"], R(lv["refined"], (H, W), cam, bg_r)["img"],
                    R(cat(lv["scene"], lv["refined"]), (H, W), cam)["img"]]
        loss = sum(torch.nn.functional.l1_loss(im, tgt) for im in imgs)
        loss.backward()

    for i in range(5):
        frame(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(a.frames):
        frame(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"eager: {a.frames / dt:.1f} training frames/s, {dt / a.frames * 1e3:.3f} ms per frame ({('fused+graph' if a.graph else 'fused') if a.fused else 'five calls'})")
    pr = cProfile.Profile()
    pr.enable()
    for i in range(a.frames):
        frame(i)
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(28)
    st.sort_stats("tottime").print_stats(22)


if __name__ == "__main__":
    main()
