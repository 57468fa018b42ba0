"""Real spherical-harmonics colour (degree <= 3), the polynomial ExAvatar evaluates for scene Gaussians.

Restates what `eval_sh` computes at /root/reference/avatar/common/utils/transforms.py:112-167 (called from
`SceneGaussian.forward`, avatar/common/nets/module.py:258-266) so that the caller-side path "SH -> rgb in PyTorch, then
`colors_precomp`" can be compared with the in-kernel SH path of the rasteriser (`shs=` + `sh_degree`, SURVEY.md section
8f-4).  Device-agnostic and differentiable; values are pinned against the reference by tests/golden/sh.npz.
"""
from __future__ import annotations

import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435)


def sh_basis(deg: int, dirs: torch.Tensor) -> torch.Tensor:
    """(P, (deg+1)^2) basis values for unit directions `dirs` (P, 3)."""
    x, y, z = dirs[:, 0], dirs[:, 1], dirs[:, 2]
    b = [torch.full_like(x, C0)]
    if deg > 0:
        b += [-C1 * y, C1 * z, -C1 * x]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        b += [C2[0] * xy, C2[1] * yz, C2[2] * (2.0 * zz - xx - yy), C2[3] * xz, C2[4] * (xx - yy)]
    if deg > 2:
        b += [C3[0] * y * (3 * xx - yy), C3[1] * xy * z, C3[2] * y * (4 * zz - xx - yy),
              C3[3] * z * (2 * zz - 3 * xx - 3 * yy), C3[4] * x * (4 * zz - xx - yy), C3[5] * z * (xx - yy),
              C3[6] * x * (xx - 3 * yy)]
    return torch.stack(b, dim=1)


def sh_to_rgb(deg: int, shs: torch.Tensor, means: torch.Tensor, campos: torch.Tensor) -> torch.Tensor:
    """`shs` (P, M, 3) -> rgb (P, 3) = max(0, sum_k basis_k(dir) * shs[:, k] + 0.5), dir = normalize(mean - campos)
    (module.py:261-266)."""
    dirs = torch.nn.functional.normalize(means - campos.reshape(1, 3), p=2, dim=1)
    basis = sh_basis(deg, dirs)
    rgb = (basis[:, :, None] * shs[:, : basis.shape[1], :]).sum(1)
    return torch.clamp_min(rgb + 0.5, 0.0)
