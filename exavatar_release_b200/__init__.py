"""exavatar_release_b200 -- B200-native differentiable Gaussian rasteriser for ExAvatar's render path.

Public surface (mirrors what /root/reference/avatar/common/nets/module.py:11 imports):
    GaussianRasterizationSettings, GaussianRasterizer
plus the host-side mirror of the caller (`GaussianRenderer`, module.py:588-647), `TrainingFrameRenderer` (the five
renders of a training frame, avatar/main/model.py:117-162, as one autograd call), synthetic workloads and the
frame-sharding helper used by bench.py.
"""
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians  # noqa: F401
from .renderer import GaussianRenderer, render_settings  # noqa: F401


def __getattr__(name):  # TrainingFrameRenderer pulls in the plan machinery; loaded on first use
    if name == "TrainingFrameRenderer":
        from .fused import TrainingFrameRenderer
        return TrainingFrameRenderer
    raise AttributeError(name)


__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians", "GaussianRenderer",
           "render_settings", "TrainingFrameRenderer"]
