"""Independent dense fp64 restatement of the rasteriser, differentiated by torch.autograd.

TEST INFRASTRUCTURE ONLY (small cases).  Purpose: cross-validate the hand-derived backward of
oracle/gs_oracle.c.  The forward is written with plain torch ops following SURVEY.md App. A.1-A.3;
the backward is autograd's, with the four places where the published backward is NOT the autograd
derivative of its forward (App. A.6) encoded explicitly:
  (i)   min(0.99, .) clamp is straight-through,
  (ii)  conic gradient uses 1/(det^2 + 1e-7),
  (iii) a frustum-clamped t.x / t.y passes no gradient,
  (iv)  all masks are constants.
`means2D` enters as a dummy added to the NDC position so that its autograd gradient is the
NDC-scaled screen-space gradient ExAvatar reads for densification (module.py:626-629).
"""
import math

import torch

F32 = lambda v: float(torch.tensor(v, dtype=torch.float32))
K_NEAR, K_DIL, K_AMAX, K_AMIN, K_TMIN, K_FRU, K_EPS, K_EIG = (F32(0.2), F32(0.3), F32(0.99), F32(1.0 / 255.0),
                                                              F32(0.0001), F32(1.3), F32(0.0000001), F32(0.1))
C0 = F32(0.28209479177387814)
C1 = F32(0.4886025119029199)
C2 = [F32(v) for v in (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)]
C3 = [F32(v) for v in (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
                       -0.4570457994644658, 1.445305721320277, -0.5900435899266435)]


class _Conic(torch.autograd.Function):
    """conic = (c, -b, a)/det with the published backward (App. A.5 / A.6(ii))."""

    @staticmethod
    def forward(ctx, a, b, c):
        det = a * c - b * b
        ctx.save_for_backward(a, b, c)
        return c / det, -b / det, a / det

    @staticmethod
    def backward(ctx, gx, gy, gz):
        a, b, c = ctx.saved_tensors
        denom = a * c - b * b
        d2 = 1.0 / (denom * denom + K_EPS)
        # gy here is the TRUE dL/dconic_y (autograd); the published kernels carry half of it and double it back
        # in this formula (App. A.4 "not doubled here"), so the two agree.
        da = d2 * (-c * c * gx + b * c * gy + (denom - a * c) * gz)
        dc = d2 * (-a * a * gz + a * b * gy + (denom - a * c) * gx)
        db = d2 * (2 * b * c * gx - (denom + 2 * b * b) * gy + 2 * a * b * gz)
        return da, db, dc


def _sh_rgb(deg, sh, dirs):
    """transforms.py:112-167 polynomial on (P,M,3) layout."""
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    r = C0 * sh[:, 0]
    if deg > 0:
        r = r - C1 * y * sh[:, 1] + C1 * z * sh[:, 2] - C1 * x * sh[:, 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            r = (r + C2[0] * xy * sh[:, 4] + C2[1] * yz * sh[:, 5] + C2[2] * (2.0 * zz - xx - yy) * sh[:, 6]
                 + C2[3] * xz * sh[:, 7] + C2[4] * (xx - yy) * sh[:, 8])
            if deg > 2:
                r = (r + C3[0] * y * (3 * xx - yy) * sh[:, 9] + C3[1] * xy * z * sh[:, 10]
                     + C3[2] * y * (4 * zz - xx - yy) * sh[:, 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12]
                     + C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + C3[5] * z * (xx - yy) * sh[:, 14]
                     + C3[6] * x * (xx - 3 * yy) * sh[:, 15])
    return r


def render(settings, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
           cov3D_precomp=None):
    """Returns (color (3,H,W), radii (P), depth (1,H,W), alpha (1,H,W)); all float64, autograd-connected."""
    dd = torch.float64
    H, W = int(settings.image_height), int(settings.image_width)
    P = means3D.shape[0]
    Vt = settings.viewmatrix.to(dd).contiguous()      # stored transposed: Vt[c, r] = V[r, c]
    PVt = settings.projmatrix.to(dd).contiguous()
    V = Vt.t()
    PV = PVt.t()
    tfx, tfy = float(settings.tanfovx), float(settings.tanfovy)
    fx, fy = W / (2.0 * tfx), H / (2.0 * tfy)
    bg = settings.bg.to(dd)
    mod = float(settings.scale_modifier)

    p = means3D.to(dd)
    ones = torch.ones(P, 1, dtype=dd)
    ph1 = torch.cat([p, ones], 1)
    pv = ph1 @ V.t()[:, :3]                            # (P,3) view-space
    hom = ph1 @ PV.t()                                 # (P,4)
    pw = 1.0 / (hom[:, 3] + K_EPS)
    ndc = hom[:, :2] * pw[:, None] + means2D.to(dd)[:, :2]
    vis = pv[:, 2] > K_NEAR

    if cov3D_precomp is not None:
        c6 = cov3D_precomp.to(dd)
        Sig = torch.stack([torch.stack([c6[:, 0], c6[:, 1], c6[:, 2]], -1),
                           torch.stack([c6[:, 1], c6[:, 3], c6[:, 4]], -1),
                           torch.stack([c6[:, 2], c6[:, 4], c6[:, 5]], -1)], 1)
    else:
        q = rotations.to(dd)
        r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        R = torch.stack([torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)], -1),
                         torch.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)], -1),
                         torch.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1)], 1)
        s = mod * scales.to(dd)
        N = R * s[:, None, :]
        Sig = N @ N.transpose(1, 2)

    tz = pv[:, 2]
    limx, limy = K_FRU * tfx, K_FRU * tfy
    txtz, tytz = pv[:, 0] / tz, pv[:, 1] / tz
    cx = (txtz < -limx) | (txtz > limx)
    cy = (tytz < -limy) | (tytz > limy)
    tx = torch.where(cx, (txtz.clamp(-limx, limx) * tz).detach(), pv[:, 0])
    ty = torch.where(cy, (tytz.clamp(-limy, limy) * tz).detach(), pv[:, 1])
    zero = torch.zeros_like(tz)
    J = torch.stack([torch.stack([fx / tz, zero, -(fx * tx) / (tz * tz)], -1),
                     torch.stack([zero, fy / tz, -(fy * ty) / (tz * tz)], -1)], 1)   # (P,2,3)
    A = J @ V[:3, :3]
    cov2 = A @ Sig @ A.transpose(1, 2)
    a = cov2[:, 0, 0] + K_DIL
    b = cov2[:, 0, 1]
    c = cov2[:, 1, 1] + K_DIL
    det = a * c - b * b
    vis = vis & (det != 0)
    safe = lambda t: torch.where(vis, t, torch.ones_like(t))
    con_x, con_y, con_z = _Conic.apply(safe(a), torch.where(vis, b, torch.zeros_like(b)), safe(c))
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=K_EIG))
    radius = torch.ceil(3.0 * torch.sqrt(lam.detach().clamp(min=0))).to(torch.int64)
    pix_x = ((ndc[:, 0] + 1.0) * W - 1.0) * 0.5
    pix_y = ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5
    gx, gy = (W + 15) // 16, (H + 15) // 16
    trunc = lambda t: torch.trunc(t).to(torch.int64)
    rf = radius.to(dd)
    x0 = trunc((pix_x.detach() - rf) / 16).clamp(0, gx)
    y0 = trunc((pix_y.detach() - rf) / 16).clamp(0, gy)
    x1 = trunc((pix_x.detach() + rf + 15) / 16).clamp(0, gx)
    y1 = trunc((pix_y.detach() + rf + 15) / 16).clamp(0, gy)
    vis = vis & (((x1 - x0) * (y1 - y0)) > 0)
    radii = torch.where(vis, radius, torch.zeros_like(radius)).to(torch.int32)

    if shs is not None:
        d = p - settings.campos.to(dd)[None]
        dirs = d / d.norm(dim=1, keepdim=True)
        rgb = torch.clamp_min(_sh_rgb(int(settings.sh_degree), shs.to(dd), dirs) + 0.5, 0.0)
    else:
        rgb = colors_precomp.to(dd)
    op = opacities.to(dd).reshape(P)

    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    pxf, pyf = xs.to(dd), ys.to(dd)
    txi, tyi = xs // 16, ys // 16
    order = sorted([i for i in range(P) if bool(vis[i])], key=lambda i: (float(pv[i, 2].detach()), i))
    T = torch.ones(H, W, dtype=dd)
    done = torch.zeros(H, W, dtype=torch.bool)
    Cc = torch.zeros(3, H, W, dtype=dd)
    Dp = torch.zeros(H, W, dtype=dd)
    Aa = torch.zeros(H, W, dtype=dd)
    for i in order:
        in_tile = (txi >= x0[i]) & (txi < x1[i]) & (tyi >= y0[i]) & (tyi < y1[i])
        dx = pix_x[i] - pxf
        dy = pix_y[i] - pyf
        power = -0.5 * (con_x[i] * dx * dx + con_z[i] * dy * dy) - con_y[i] * dx * dy
        G = torch.exp(torch.clamp(power, max=0.0))
        a_raw = op[i] * G
        alpha = a_raw + (torch.clamp(a_raw, max=K_AMAX) - a_raw).detach()
        ok = in_tile & (power.detach() <= 0) & (alpha.detach() >= K_AMIN) & ~done
        test = T * (1 - alpha)
        stop = ok & (test.detach() < K_TMIN)
        done = done | stop
        use = ok & ~stop
        w = torch.where(use, alpha * T, torch.zeros_like(T))
        Cc = Cc + rgb[i][:, None, None] * w[None]
        Dp = Dp + pv[i, 2] * w
        Aa = Aa + w
        T = torch.where(use, test, T)
    color = Cc + T[None] * bg[:, None, None]
    return color, radii, Dp[None], Aa[None]
