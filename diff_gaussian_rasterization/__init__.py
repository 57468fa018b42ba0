"""Import-compatible alias for the vanilla package name (commented alternative at
/root/reference/avatar/common/nets/module.py:10)."""
from exavatar_release_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                              rasterize_gaussians)
