"""Host-side mirror of ExAvatar's render boundary.

`GaussianRenderer.forward` restates /root/reference/avatar/common/nets/module.py:592-647 (the only
caller of the rasteriser) without its hard-coded `.cuda()` calls, so the same code drives the B200
rasteriser on a GPU box and the CPU oracle in tests.  Argument meaning, the settings tuple
(module.py:609-622), the dummy `mean_2d` leaf (module.py:626-629) and the returned dict
(module.py:642-647) are the reference's.

`render_settings` additionally caches the per-camera setup: the reference rebuilds the matrices with
~8 host<->device syncs per call; a caller that renders five asset sets with one
camera (avatar/main/model.py:130-162) can build the settings once and reuse them.
"""
from __future__ import annotations

import torch
from torch import nn

from .camera import get_fov, get_proj_matrix, get_view_matrix
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer


def render_settings(img_shape, cam_param, bg, settings_cls=GaussianRasterizationSettings):
    """module.py:604-622: fov, transposed view / full-projection matrices, camera position, settings tuple."""
    fov = get_fov(cam_param["focal"], cam_param["princpt"], img_shape)
    view_matrix = get_view_matrix(cam_param["R"], cam_param["t"]).permute(1, 0)
    proj_matrix = get_proj_matrix(cam_param["focal"], cam_param["princpt"], img_shape, 0.01, 100, 1.0).permute(1, 0)
    full_proj_matrix = torch.mm(view_matrix, proj_matrix)
    cam_pos = view_matrix.inverse()[3, :3]
    return settings_cls(
        image_height=img_shape[0],
        image_width=img_shape[1],
        tanfovx=float(torch.tan(fov[0] / 2)),
        tanfovy=float(torch.tan(fov[1] / 2)),
        bg=bg,
        scale_modifier=1.0,
        viewmatrix=view_matrix,
        projmatrix=full_proj_matrix,
        sh_degree=0,  # colours are precomputed by the caller (module.py:618)
        campos=cam_pos,
        prefiltered=False,
        debug=False,
    )


class GaussianRenderer(nn.Module):
    """Same call as module.py:588-647; `rasterizer_cls` / `settings_cls` let tests substitute the CPU oracle."""

    def __init__(self, rasterizer_cls=GaussianRasterizer, settings_cls=GaussianRasterizationSettings):
        super().__init__()
        self.rasterizer_cls = rasterizer_cls
        self.settings_cls = settings_cls

    def forward(self, gaussian_assets, img_shape, cam_param, bg=None, raster_settings=None):
        mean_3d = gaussian_assets["mean_3d"]
        if bg is None:  # reference default: white (module.py:592)
            bg = torch.ones(3, dtype=torch.float32, device=mean_3d.device)
        if raster_settings is None:
            raster_settings = render_settings(img_shape, cam_param, bg, self.settings_cls)
        rasterizer = self.rasterizer_cls(raster_settings=raster_settings)

        # screen-space positions: a zero leaf whose .grad is read after backward (train.py:51, model.py:285)
        mean_2d = torch.zeros((mean_3d.shape[0], 3), dtype=torch.float32, device=mean_3d.device)
        mean_2d.requires_grad = True
        mean_2d.retain_grad()

        render_img, radius, render_depthmap, render_mask = rasterizer(
            means3D=mean_3d,
            means2D=mean_2d,
            shs=None,
            colors_precomp=gaussian_assets["rgb"],
            opacities=gaussian_assets["opacity"],
            scales=gaussian_assets["scale"],
            rotations=gaussian_assets["rotation"],
            cov3D_precomp=None)

        return {"img": render_img,
                "depthmap": render_depthmap,
                "mask": render_mask,
                "mean_2d": mean_2d,
                "is_vis": radius > 0,
                "radius": radius}
