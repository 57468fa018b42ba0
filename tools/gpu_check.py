"""Development diagnostic: CUDA rasteriser vs CPU oracle on seeded workloads, printing per-tensor errors.

Run on a GPU box:  python tools/gpu_check.py [workload ...]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from util import workload_settings  # noqa: E402
from exavatar_release_b200 import rasterizer as RZ  # noqa: E402
from exavatar_release_b200.synthetic import WORKLOADS, make_assets, make_grad_image  # noqa: E402
from oracle import oracle as O  # noqa: E402


def err(name, x, y):
    x = np.asarray(x, np.float64); y = np.asarray(y, np.float64)
    if y.size == 0:
        return
    ninf = np.abs(y).max()
    d = np.abs(x - y)
    i = np.unravel_index(np.argmax(d), d.shape)
    rel = (d / np.maximum(np.abs(y), 1e-2 * ninf + 1e-30)).max()
    print(f"   {name:10s} |y|inf={ninf:10.4g} max|d|={d.max():10.3g} at {i} (x={x[i]:.6g} y={y[i]:.6g})  "
          f"d/|y|inf={d.max() / (ninf + 1e-30):8.2e}  elem-rel(floor 1e-2)={rel:8.2e}  n(d>1e-4|y|inf)={(d > 1e-4 * ninf).sum()}",
          flush=True)


def run(name, yaw=12.0, seed=0, da=False):
    wl = WORKLOADS[name]
    dev = torch.device("cuda:0")
    assets = make_assets(name, seed=seed)
    st_cpu = workload_settings(name, yaw=yaw, bg=(0.2, 0.6, 0.9))
    st_gpu = workload_settings(name, yaw=yaw, bg=(0.2, 0.6, 0.9), device=dev, settings_cls=RZ.GaussianRasterizationSettings)
    use_sh = wl.sh_degree > 0
    if use_sh:
        st_cpu = st_cpu._replace(sh_degree=wl.sh_degree)
        st_gpu = st_gpu._replace(sh_degree=wl.sh_degree)
    print(f"== {wl.name} yaw={yaw} seed={seed} P={assets['mean_3d'].shape[0]}", flush=True)
    t0 = time.time()
    oc, orad, od, oa, octx = O.forward(st_cpu, assets["mean_3d"], assets["opacity"], assets.get("shs") if use_sh else None,
                                       None if use_sh else assets["rgb"], assets["scale"], assets["rotation"])
    print(f"   oracle fwd {time.time() - t0:.2f}s  D={octx.num_dups} consumed_fwd={octx.consumed_fwd}", flush=True)

    g = {k: v.to(dev).requires_grad_() for k, v in assets.items()}
    m2 = torch.zeros(g["mean_3d"].shape[0], 3, device=dev, requires_grad=True)
    rast = RZ.GaussianRasterizer(st_gpu)
    color, radii, depth, alpha = rast(means3D=g["mean_3d"], means2D=m2, opacities=g["opacity"],
                                      shs=g["shs"] if use_sh else None, colors_precomp=None if use_sh else g["rgb"],
                                      scales=g["scale"], rotations=g["rotation"])
    torch.cuda.synchronize()
    cx = color.grad_fn.cx if hasattr(color.grad_fn, "cx") else None
    print("   radii equal:", bool((radii.cpu().numpy() == orad).all()), " mismatches:", int((radii.cpu().numpy() != orad).sum()))
    err("color", color.detach().cpu().numpy(), oc)
    err("depth", depth.detach().cpu().numpy(), od)
    err("alpha", alpha.detach().cpu().numpy(), oa)
    gi = make_grad_image(name, seed)
    gd = ga = None
    loss = (color * gi.to(dev)).sum()
    if da:
        gen = torch.Generator().manual_seed(77)
        gd = torch.randn(1, wl.height, wl.width, generator=gen)
        ga = torch.randn(1, wl.height, wl.width, generator=gen)
        loss = loss + (depth * gd.to(dev)).sum() + (alpha * ga.to(dev)).sum()
    loss.backward()
    torch.cuda.synchronize()
    t0 = time.time()
    og = O.backward(octx, gi.numpy(), None if gd is None else gd.numpy()[0], None if ga is None else ga.numpy()[0])
    print(f"   oracle bwd {time.time() - t0:.2f}s consumed_bwd={octx.consumed_bwd}", flush=True)
    err("d_means3D", g["mean_3d"].grad.cpu().numpy(), og["means3D"])
    err("d_means2D", m2.grad.cpu().numpy(), og["means2D"])
    err("d_opacity", g["opacity"].grad.cpu().numpy(), og["opacities"])
    err("d_scales", g["scale"].grad.cpu().numpy(), og["scales"])
    err("d_rots", g["rotation"].grad.cpu().numpy(), og["rotations"])
    if use_sh:
        err("d_shs", g["shs"].grad.cpu().numpy(), og["shs"])
    else:
        err("d_colors", g["rgb"].grad.cpu().numpy(), og["colors"])


if __name__ == "__main__":
    names = sys.argv[1:] or ["T0", "T1", "T2", "C1"]
    for n in names:
        run(n, da=(n in ("T1", "T2")))
