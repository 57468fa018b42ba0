"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Exposes the oracle behind the same call signature as the rasteriser ExAvatar uses
(`GaussianRasterizer(raster_settings)(means3D=..., means2D=..., ...)`,
/root/reference/avatar/common/nets/module.py:609-640) so tests read like the caller, and
`bench.py --impl reference` / `cpu_baseline` can time it on host cores.

Nothing under exavatar_release_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional

import numpy as np
import torch

from . import build as _build

_LIBS = {}


def _lib(variant: str):
    if variant in _LIBS:
        return _LIBS[variant]
    _build.build()
    lib = C.CDLL(_build.lib_path(variant))
    real = C.c_float if variant == "f32" else C.c_double
    rp = C.c_void_p
    lib.gso_forward.restype = C.c_void_p
    lib.gso_forward.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, rp, rp, rp, rp, rp, rp, rp, real, rp, rp, rp,
                                real, real, rp, rp, rp, rp, rp]
    lib.gso_backward.restype = None
    lib.gso_backward.argtypes = [C.c_void_p] + [rp] * 11
    lib.gso_free.restype = None
    lib.gso_free.argtypes = [C.c_void_p]
    lib.gso_mark_visible.restype = None
    lib.gso_mark_visible.argtypes = [C.c_int, rp, rp, rp]
    lib.gso_fragility.restype = None
    lib.gso_fragility.argtypes = [C.c_void_p, real, real, rp, rp]
    for name in ("gso_num_dups", "gso_consumed_fwd", "gso_consumed_bwd"):
        getattr(lib, name).restype = C.c_int64
        getattr(lib, name).argtypes = [C.c_void_p]
    for name in ("gso_xy", "gso_depth", "gso_conic_opacity", "gso_cov3D", "gso_rgb", "gso_rect", "gso_tiles_touched",
                 "gso_list", "gso_ranges", "gso_final_T", "gso_n_contrib"):
        getattr(lib, name).restype = C.c_void_p
        getattr(lib, name).argtypes = [C.c_void_p]
    lib.gso_num_threads.restype = C.c_int
    lib.gso_set_num_threads.argtypes = [C.c_int]
    _LIBS[variant] = lib
    return lib


def set_num_threads(n: int) -> None:
    for v in ("f32", "f64"):
        _lib(v).gso_set_num_threads(int(n))


def num_threads() -> int:
    return int(_lib("f32").gso_num_threads())


def _np(t: Optional[torch.Tensor], dtype) -> Optional[np.ndarray]:
    if t is None or t.numel() == 0:
        return None
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(dtype, copy=False))


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleContext:
    """One forward pass's saved state (geometry, sorted per-tile lists, per-pixel T / n_contrib)."""

    def __init__(self, variant, handle, P, W, H, M, keep):
        self.variant, self.handle, self.P, self.W, self.H, self.M = variant, handle, P, W, H, M
        self._keep = keep
        self.dtype = np.float32 if variant == "f32" else np.float64

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib(self.variant).gso_free(self.handle)
                self.handle = None
        except Exception:  # interpreter shutdown
            pass

    def _arr(self, name, ctype, shape):
        lib = _lib(self.variant)
        p = getattr(lib, name)(self.handle)
        n = int(np.prod(shape))
        if n == 0 or not p:
            return np.zeros(shape, dtype=ctype)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(np.ctypeslib.as_ctypes_type(ctype))), shape=(n,)).reshape(shape).copy()

    @property
    def num_dups(self):
        return int(_lib(self.variant).gso_num_dups(self.handle))

    @property
    def consumed_fwd(self):
        return int(_lib(self.variant).gso_consumed_fwd(self.handle))

    @property
    def consumed_bwd(self):
        return int(_lib(self.variant).gso_consumed_bwd(self.handle))

    def xy(self): return self._arr("gso_xy", self.dtype, (self.P, 2))
    def depth(self): return self._arr("gso_depth", self.dtype, (self.P,))
    def conic_opacity(self): return self._arr("gso_conic_opacity", self.dtype, (self.P, 4))
    def cov3D(self): return self._arr("gso_cov3D", self.dtype, (self.P, 6))
    def rgb(self): return self._arr("gso_rgb", self.dtype, (self.P, 3))
    def rect(self): return self._arr("gso_rect", np.int32, (self.P, 4))
    def tiles_touched(self): return self._arr("gso_tiles_touched", np.uint32, (self.P,))
    def sorted_ids(self): return self._arr("gso_list", np.uint32, (self.num_dups,))
    def ranges(self):
        tn = ((self.W + 15) // 16) * ((self.H + 15) // 16)
        return self._arr("gso_ranges", np.int64, (tn, 2))
    def final_T(self): return self._arr("gso_final_T", self.dtype, (self.H, self.W))
    def n_contrib(self): return self._arr("gso_n_contrib", np.uint32, (self.H, self.W))


def forward(settings, means3D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
            cov3D_precomp=None, variant="f32"):
    """Returns (color (3,H,W), radii (P) int32, depth (1,H,W), alpha (1,H,W), OracleContext) as numpy/torch-on-CPU."""
    lib = _lib(variant)
    dt = np.float32 if variant == "f32" else np.float64
    P = int(means3D.shape[0])
    H, W = int(settings.image_height), int(settings.image_width)
    m3 = _np(means3D, dt)
    sh = _np(shs, dt)
    M = 0 if sh is None else int(sh.shape[1])
    cp = _np(colors_precomp, dt)
    op = _np(opacities, dt)
    sc = _np(scales, dt)
    ro = _np(rotations, dt)
    cv = _np(cov3D_precomp, dt)
    # matrices arrive as transposed views (module.py:605-607): flat row-major storage of the
    # given tensor is what the rasteriser reads, i.e. element (r,c) of the maths matrix at [4c+r].
    view = _np(settings.viewmatrix.contiguous(), dt).reshape(-1)
    proj = _np(settings.projmatrix.contiguous(), dt).reshape(-1)
    campos = _np(settings.campos, dt).reshape(-1)
    bg = _np(settings.bg, dt).reshape(-1)
    if P > 0 and (sh is None) == (cp is None):
        raise Exception("Please provide excatly one of either SHs or precomputed colors!")
    color = np.zeros((3, H, W), dt)
    depth = np.zeros((1, H, W), dt)
    alpha = np.zeros((1, H, W), dt)
    radii = np.zeros((P,), np.int32)
    if P == 0:  # the published binding returns a zero image without launching anything [EXT]
        return color, radii, depth, alpha, None
    real = C.c_float if variant == "f32" else C.c_double
    h = lib.gso_forward(P, W, H, int(settings.sh_degree), M, _ptr(m3), _ptr(sh), _ptr(cp), _ptr(op), _ptr(sc), _ptr(ro),
                        _ptr(cv), real(float(settings.scale_modifier)), _ptr(view), _ptr(proj), _ptr(campos),
                        real(float(settings.tanfovx)), real(float(settings.tanfovy)), _ptr(bg), _ptr(color), _ptr(depth),
                        _ptr(alpha), _ptr(radii))
    ctx = OracleContext(variant, h, P, W, H, M, (m3, sh, cp, op, sc, ro, cv))
    return color, radii, depth, alpha, ctx


def backward(ctx: OracleContext, dL_dcolor, dL_ddepth=None, dL_dalpha=None):
    """Returns dict of numpy gradients: means3D, means2D, shs, colors, opacities, scales, rotations, cov3D."""
    lib = _lib(ctx.variant)
    dt = ctx.dtype
    P, M = ctx.P, ctx.M
    gc = np.ascontiguousarray(np.asarray(dL_dcolor, dtype=dt))
    gd = None if dL_ddepth is None else np.ascontiguousarray(np.asarray(dL_ddepth, dtype=dt))
    ga = None if dL_dalpha is None else np.ascontiguousarray(np.asarray(dL_dalpha, dtype=dt))
    out = {
        "means3D": np.zeros((P, 3), dt), "means2D": np.zeros((P, 3), dt), "shs": np.zeros((P, M, 3), dt),
        "colors": np.zeros((P, 3), dt), "opacities": np.zeros((P, 1), dt), "scales": np.zeros((P, 3), dt),
        "rotations": np.zeros((P, 4), dt), "cov3D": np.zeros((P, 6), dt),
    }
    lib.gso_backward(ctx.handle, _ptr(gc), _ptr(gd), _ptr(ga), _ptr(out["means3D"]), _ptr(out["means2D"]),
                     _ptr(out["shs"]) if M > 0 else None, _ptr(out["colors"]), _ptr(out["opacities"]), _ptr(out["scales"]),
                     _ptr(out["rotations"]), _ptr(out["cov3D"]))
    return out


def fragility(ctx: OracleContext, eps_alpha=2e-5, eps_T=2e-3):
    """(pixel mask (H,W) bool, Gaussian mask (P) bool): where a discrete composite decision sits on its threshold.

    eps_T is wider than eps_alpha on purpose: T is a product of (1 - alpha) factors, and for the near-opaque splats
    ExAvatar produces (opacity == 1, module.py:565) 1 - alpha cancels catastrophically -- a 1e-6 relative error in
    alpha ~ 0.99 is a 1e-4 relative error in that factor, in ANY fp32 implementation."""
    lib = _lib(ctx.variant)
    real = C.c_float if ctx.variant == "f32" else C.c_double
    pm = np.zeros((ctx.H, ctx.W), np.uint8)
    gm = np.zeros((max(ctx.P, 1),), np.uint8)
    lib.gso_fragility(ctx.handle, real(eps_alpha), real(eps_T), _ptr(pm), _ptr(gm))
    return pm.astype(bool), gm[: ctx.P].astype(bool)


def mark_visible(positions, viewmatrix, variant="f32"):
    dt = np.float32 if variant == "f32" else np.float64
    p = _np(positions, dt)
    v = _np(viewmatrix.contiguous(), dt).reshape(-1)
    out = np.zeros((p.shape[0],), np.uint8)
    _lib(variant).gso_mark_visible(p.shape[0], _ptr(p), _ptr(v), _ptr(out))
    return torch.from_numpy(out.astype(bool))


class OracleSettings(NamedTuple):
    """Same 12 fields, same order as the call site module.py:609-622."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class _OracleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp, settings, variant):
        none = lambda t: None if t is None or t.numel() == 0 else t
        color, radii, depth, alpha, octx = forward(settings, means3D, opacities, none(shs), none(colors_precomp),
                                                   none(scales), none(rotations), none(cov3D_precomp), variant)
        ctx.octx = octx
        ctx.tdtype = means3D.dtype
        ctx.shapes = (shs.shape if shs is not None else None,)
        td = means3D.dtype
        outs = (torch.from_numpy(color).to(td), torch.from_numpy(radii), torch.from_numpy(depth).to(td),
                torch.from_numpy(alpha).to(td))
        ctx.mark_non_differentiable(outs[1])
        return outs

    @staticmethod
    def backward(ctx, g_color, g_radii, g_depth, g_alpha):
        o = ctx.octx
        z = lambda g: None if g is None else g.detach().cpu().numpy().reshape(o.H, o.W)
        gr = backward(o, g_color.detach().cpu().numpy(), z(g_depth), z(g_alpha))
        td = ctx.tdtype
        t = lambda k: torch.from_numpy(gr[k]).to(td)
        return (t("means3D"), t("means2D"), t("shs") if o.M > 0 else None, t("colors"), t("opacities"), t("scales"),
                t("rotations"), t("cov3D"), None, None)


class OracleRasterizer(torch.nn.Module):
    """CPU stand-in with the call signature of `GaussianRasterizer` (module.py:623,632-640)."""

    def __init__(self, raster_settings, variant="f32"):
        super().__init__()
        self.raster_settings = raster_settings
        self.variant = variant

    def markVisible(self, positions):
        with torch.no_grad():
            return mark_visible(positions, self.raster_settings.viewmatrix, self.variant)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                (scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        e = torch.empty(0, dtype=means3D.dtype)
        return _OracleFn.apply(means3D, means2D, e if shs is None else shs, e if colors_precomp is None else colors_precomp,
                               opacities, e if scales is None else scales, e if rotations is None else rotations,
                               e if cov3D_precomp is None else cov3D_precomp, self.raster_settings, self.variant)
