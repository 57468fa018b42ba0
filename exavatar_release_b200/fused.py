"""`TrainingFrameRenderer`: ExAvatar's five renders of a training frame as ONE autograd call.

`Model.forward` renders every frame five times with the same camera (avatar/main/model.py:117-162):

    scene_render                = gaussian_renderer(scene_asset, ...)
    human_render                = gaussian_renderer(human_asset, ..., bg)
    scene_human_render          = gaussian_renderer(cat(scene_asset.detach(), human_asset), ...)
    human_render_refined        = gaussian_renderer(human_asset_refined, ..., bg)
    scene_human_render_refined  = gaussian_renderer(cat(scene_asset.detach(), human_asset_refined), ...)

Through the drop-in `GaussianRasterizer` those are five full projection / binning / sort / composite pipelines.  This
module renders the same five images from TWO projection + binning passes (`plan.MergedFivePlan`: views of
cat(scene, human) and cat(scene, human_refined), human-free tiles of the human-only and combined views skipped) and
back-propagates into the three asset dicts exactly what `loss.backward()` leaves there: the scene render's gradient in the
scene assets, human-only + combined render in the human assets (the scene part of the combined renders is detached, as in
the reference), likewise for the refined set.  INTEGRATION.md shows the replacement of model.py:117-162.

    frame = TrainingFrameRenderer(P_scene, P_human, (H, W), device, dup_capacity)
    out = frame(scene_asset, human_asset, human_asset_refined, cam_param, bg_human)
    out["scene"]["img"], out["human"]["mask"], out["scene_human"]["img"], ...      # same keys as GaussianRenderer
    out["scene"]["mean_2d"].grad                                                   # after backward (train.py:51)

The duplicate capacity is fixed per instance (every buffer is resident; nothing is polled or synchronised, so the call
is capturable in a CUDA graph); `overflowed()` reports if a frame needed more (its lists were truncated, never corrupt).
One frame may be in flight per instance: run backward (or drop the outputs) before the next call.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import nn

from .plan import RENDERS, MergedFivePlan, _views_of
from .rasterizer import GaussianRasterizationSettings, _f32c
from .renderer import render_settings

_KEYS = ("mean_3d", "opacity", "scale", "rotation", "rgb")
_GRAD_OF = {"mean_3d": "means3D", "opacity": "opacities", "scale": "scales", "rotation": "rotations", "rgb": "colors"}


class _FrameFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, settings, settings_h, scene_m2d, *tensors):
        plan: MergedFivePlan = mod.plan
        scene, human, refined = (dict(zip(_KEYS, (_f32c(t.detach(), k) for k, t in zip(_KEYS, tensors[i * 5:i * 5 + 5]))))
                                 for i in range(3))
        plan.set_scene(scene)
        mod._frame_no += 1
        plan.forward_frame(None, settings, settings_h, scene, human, refined)  # no descriptor cache: cameras change
        outs = []
        for r in RENDERS:  # fresh tensors: the plan's image buffers are overwritten by the next frame
            pk = "A" if r in plan.VIEWS["A"] else "B"
            color, depth, alpha = plan.passes[pk].img[plan.VIEWS[pk].index(r)]
            outs += [color.clone(), depth.clone(), alpha.clone()]
        radii_a, radii_b = plan.passes["A"].radii.clone(), plan.passes["B"].radii.clone()
        ctx.mod = mod
        ctx.shapes = [t.shape for t in tensors]
        ctx.m2d_shape = scene_m2d.shape
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(radii_a, radii_b)
        return (*outs, radii_a, radii_b)

    @staticmethod
    def backward(ctx, *grads):
        mod = ctx.mod
        plan: MergedFivePlan = mod.plan
        g = grads[:15]
        gc = {r: (None if g[3 * i] is None else _f32c(g[3 * i], "grad_color")) for i, r in enumerate(RENDERS)}
        gd = {r: (None if g[3 * i + 1] is None else _f32c(g[3 * i + 1], "grad_depth")) for i, r in enumerate(RENDERS)}
        ga = {r: (None if g[3 * i + 2] is None else _f32c(g[3 * i + 2], "grad_alpha")) for i, r in enumerate(RENDERS)}
        dev = plan.device
        flat_a = torch.empty(plan.PER * plan.P, dtype=torch.float32, device=dev)
        flat_b = torch.empty(plan.PER * plan.Ph, dtype=torch.float32, device=dev)
        _, va = _views_of(flat_a, plan.P)
        _, vb = _views_of(flat_b, plan.Ph)
        plan.backward_frame(gc, va, vb, g_depths=gd, g_alphas=ga, densify=mod.densify)
        Ps = plan.Ps
        out = [None, None, None, va["means2D"][:Ps].reshape(ctx.m2d_shape)]
        for part in (lambda v: v[:Ps], lambda v: v[Ps:]):
            for k in _KEYS:
                out.append(part(va[_GRAD_OF[k]]))
        for k in _KEYS:
            out.append(vb[_GRAD_OF[k]])
        for i, shp in enumerate(ctx.shapes):
            out[4 + i] = out[4 + i].reshape(shp)
        return tuple(out)


class TrainingFrameRenderer(nn.Module):
    def __init__(self, P_scene: int, P_human: int, img_shape, device, dup_capacity: Optional[Dict[str, int]] = None):
        super().__init__()
        self.img_shape = (int(img_shape[0]), int(img_shape[1]))
        self.plan = MergedFivePlan(P_scene, P_human, self.img_shape[1], self.img_shape[0], dup_capacity, device)
        self.densify = None  # optional {'grad_accum','count','radius_max'} (P_scene) tensors updated by the backward
        self._frame_no = 0

    def overflowed(self) -> bool:
        return self.plan.overflowed()

    def forward(self, scene_asset, human_asset, human_asset_refined, cam_param, bg_human, bg=None, raster_settings=None,
                raster_settings_human=None):
        """Asset dicts as `GaussianRenderer.forward` takes them (mean_3d, opacity, scale, rotation, rgb); `bg_human` is the
        background of the two human-only renders (model.py:72), `bg` of the others (white by default, module.py:592).
        Returns {render name: {img, depthmap, mask, radius, is_vis[, mean_2d]}} for the five renders of plan.RENDERS."""
        dev = scene_asset["mean_3d"].device
        if bg is None:
            bg = torch.ones(3, dtype=torch.float32, device=dev)
        st = raster_settings or render_settings(self.img_shape, cam_param, bg, GaussianRasterizationSettings)
        st_h = raster_settings_human or st._replace(bg=bg_human)
        Ps = scene_asset["mean_3d"].shape[0]
        mean_2d = torch.zeros((Ps, 3), dtype=torch.float32, device=dev, requires_grad=True)  # module.py:626-629
        flat = [a[k] for a in (scene_asset, human_asset, human_asset_refined) for k in _KEYS]
        res = _FrameFn.apply(self, st, st_h, mean_2d, *flat)
        radii_a, radii_b = res[15], res[16]
        radius = {"scene": radii_a[:Ps], "human": radii_a[Ps:], "scene_human": radii_a, "human_refined": radii_b[Ps:],
                  "scene_human_refined": radii_b}
        out = {}
        for i, r in enumerate(RENDERS):
            out[r] = {"img": res[3 * i], "depthmap": res[3 * i + 1], "mask": res[3 * i + 2], "radius": radius[r],
                      "is_vis": radius[r] > 0}
        out["scene"]["mean_2d"] = mean_2d
        return out
