"""Per-CTA timeline of the composite kernels (needs a library built with B2R_NVCC_EXTRA=-DB2R_CTA_TRACE).

  B2R_NVCC_EXTRA=-DB2R_CTA_TRACE python -c "from exavatar_release_b200 import build_ext; build_ext.build(force=True)"
  python tools/cta_trace.py --workload C2 > gpurun_out/cta_trace.txt

Prints, for the forward and the backward composite of one frame: kernel span, per-CTA duration percentiles, the
longest CTAs with their list lengths, resident-CTA count over time and per-SM busy fraction -- what decides whether
the kernel is bounded by its tail (longest chains) or by throughput.
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from exavatar_release_b200.camera import look_at_cam_param  # noqa: E402
from exavatar_release_b200.plan import FramePlan, grad_bucket  # noqa: E402
from exavatar_release_b200.renderer import render_settings  # noqa: E402
from exavatar_release_b200.synthetic import WORKLOADS, make_assets, make_grad_image  # noqa: E402


def report(name, tr):
    t0, t1, sm, n = tr[:, 0].astype(np.int64), tr[:, 1].astype(np.int64), tr[:, 2].astype(np.int64), tr[:, 3].astype(np.int64)
    ok = t1 > 0
    t0, t1, sm, n = t0[ok], t1[ok], sm[ok], n[ok]
    base = t0.min()
    s, e = (t0 - base) / 1e3, (t1 - base) / 1e3
    d = e - s
    span = e.max()
    print(f"== {name}: {ok.sum()} CTAs, span {span:.1f} us, sum of CTA durations {d.sum():.0f} us, "
          f"mean residency {d.sum() / span:.0f} CTAs")
    print("   CTA duration us pct 50/90/99/max:", np.round(np.percentile(d, [50, 90, 99, 100]), 1))
    top = np.argsort(-d)[:8]
    print("   longest CTAs (blockIdx, start, end, dur, list n, sm):",
          [(int(i), round(float(s[i]), 1), round(float(e[i]), 1), round(float(d[i]), 1), int(n[i]), int(sm[i])) for i in top])
    last = np.argsort(-e)[:8]
    print("   last to finish (blockIdx, start, end, dur, list n):",
          [(int(i), round(float(s[i]), 1), round(float(e[i]), 1), round(float(d[i]), 1), int(n[i])) for i in last])
    grid = np.linspace(0, span, 21)
    res = [(int(((s <= g) & (e > g)).sum())) for g in grid]
    print("   resident CTAs at 0%,5%..100% of span:", res)
    busy = []
    for m in np.unique(sm):
        k = sm == m
        busy.append(e[k].max())
    busy = np.array(busy)
    print(f"   per-SM time of last CTA end us: min {busy.min():.1f} median {np.median(busy):.1f} max {busy.max():.1f}; "
          f"SMs used {len(busy)}")
    starts = np.sort(s)
    print("   start time of CTA #k (k=0,1000,2000,3000,last):", [round(float(starts[min(k, len(starts) - 1)]), 1) for k in (0, 1000, 2000, 3000, len(starts) - 1)])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C2")
    a = ap.parse_args()
    wl = WORKLOADS[a.workload]
    dev = torch.device("cuda:0")
    P, H, W = wl.n_avatar + wl.n_scene, wl.height, wl.width
    M = (wl.sh_degree + 1) ** 2 if wl.sh_degree > 0 else 0
    assets = make_assets(a.workload, seed=0, device=dev)
    st = render_settings((H, W), look_at_cam_param(5.0, (H, W), device=dev), torch.ones(3, device=dev))
    if M:
        st = st._replace(sh_degree=wl.sh_degree)
    gi = make_grad_image(a.workload, 0, device=dev)
    plan = FramePlan(P, W, H, 40_000_000 if a.workload in ("C3", "C5") else 12_000_000, dev, sh_coeffs=M)
    sc = plan.scene(0, st, assets)
    _, views = grad_bucket(P, dev, M)
    lib = plan.lib
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    tf = torch.zeros(tiles * 4, 4, dtype=torch.int64, device=dev)
    tb = torch.zeros(max(tiles * 4, 65536), 4, dtype=torch.int64, device=dev)  # the backward grid strides over its items: one row per CTA (its last item)
    ts = torch.zeros(tiles * 4, 4, dtype=torch.int64, device=dev)
    for fn, t in (("b2r_debug_trace_fwd", tf), ("b2r_debug_trace_bwd", tb), ("b2r_debug_trace_sort", ts)):
        f = getattr(lib, fn)
        f.argtypes = [C.c_void_p]
        f.restype = C.c_int
        assert f(t.data_ptr()) == 0
    for _ in range(3):
        tf.zero_(); tb.zero_(); ts.zero_()
        plan.forward(sc)
        if wl.backward:
            plan.backward(sc, gi, views)
    torch.cuda.synchronize()
    report("sort_mixed (n > 0: CTA class, n < 0: -n of the last warp of a warp-class CTA)", ts.cpu().numpy())
    report("composite_fwd", tf.cpu().numpy())
    if wl.backward:
        report("composite_bwd", tb.cpu().numpy())


if __name__ == "__main__":
    main()
