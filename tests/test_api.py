"""Python surface of the drop-in (no GPU needed): names, settings tuple, validation, alias packages, no CPU fallback,
and -- where /root/reference exists -- the reference's own `GaussianRenderer.forward` source executed verbatim
against this package's interface."""
import ast
import os
import sys
import types

import numpy as np
import pytest
import torch

from exavatar_release_b200 import GaussianRasterizationSettings, GaussianRasterizer
from exavatar_release_b200.synthetic import make_assets
from util import kat_settings, pack, splat

REF_MODULE = "/root/reference/avatar/common/nets/module.py"


def test_settings_tuple_matches_call_site():
    # field names and order of module.py:609-622
    assert GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix", "sh_degree",
        "campos", "prefiltered", "debug")
    st = kat_settings(settings_cls=GaussianRasterizationSettings)
    assert isinstance(st, tuple) and st.image_height == 32 and st.debug is False


def test_alias_packages_resolve_to_the_same_classes():
    import diff_gaussian_rasterization as vanilla
    import diff_gaussian_rasterization_depth as depth
    assert depth.GaussianRasterizer is GaussianRasterizer and vanilla.GaussianRasterizer is GaussianRasterizer
    assert depth.GaussianRasterizationSettings is GaussianRasterizationSettings


def test_argument_validation_same_messages_as_upstream():
    st = kat_settings(settings_cls=GaussianRasterizationSettings)
    a = pack([splat((0, 0, 2.0))])
    m2 = torch.zeros(1, 3)
    r = GaussianRasterizer(raster_settings=st)
    with pytest.raises(Exception, match="Please provide excatly one of either SHs or precomputed colors!"):
        r(means3D=a["means3D"], means2D=m2, opacities=a["opacities"], scales=a["scales"], rotations=a["rotations"])
    with pytest.raises(Exception, match="Please provide excatly one of either SHs or precomputed colors!"):
        r(means3D=a["means3D"], means2D=m2, opacities=a["opacities"], shs=torch.zeros(1, 16, 3),
          colors_precomp=a["colors_precomp"], scales=a["scales"], rotations=a["rotations"])
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(means3D=a["means3D"], means2D=m2, opacities=a["opacities"], colors_precomp=a["colors_precomp"])
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(means3D=a["means3D"], means2D=m2, opacities=a["opacities"], colors_precomp=a["colors_precomp"],
          scales=a["scales"], rotations=a["rotations"], cov3D_precomp=torch.zeros(1, 6))


def test_product_path_has_no_cpu_fallback():
    st = kat_settings(settings_cls=GaussianRasterizationSettings)
    a = pack([splat((0, 0, 2.0))])
    r = GaussianRasterizer(raster_settings=st)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        r(means3D=a["means3D"], means2D=torch.zeros(1, 3), opacities=a["opacities"], colors_precomp=a["colors_precomp"],
          scales=a["scales"], rotations=a["rotations"])


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "exavatar_release_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "gs_oracle" not in src, f


def test_missing_library_raises(monkeypatch):
    from exavatar_release_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libb200raster.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


@pytest.mark.skipif(not os.path.exists(REF_MODULE), reason="reference tree not present (GPU box)")
def test_reference_renderer_source_runs_unmodified_against_our_interface(monkeypatch):
    """Extract `class GaussianRenderer` from the reference file and execute it as-is.  Its imports are satisfied by
    this repo (camera helpers, settings / rasteriser classes with the oracle behind them on this CPU box)."""
    from exavatar_release_b200 import camera
    from oracle import oracle as O

    src = open(REF_MODULE).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "GaussianRenderer")
    code = ast.get_source_segment(src, cls)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self, raising=False)
    ns = {"torch": torch, "nn": torch.nn, "get_fov": camera.get_fov, "get_view_matrix": camera.get_view_matrix,
          "get_proj_matrix": camera.get_proj_matrix, "GaussianRasterizationSettings": O.OracleSettings,
          "GaussianRasterizer": O.OracleRasterizer}
    exec(compile(code, REF_MODULE, "exec"), ns)
    ref_renderer = ns["GaussianRenderer"]()

    assets = {k: v.requires_grad_() for k, v in make_assets("T0", seed=5).items()}
    cam = {"R": torch.eye(3), "t": torch.zeros(3), "focal": torch.tensor([93.76, 93.76]), "princpt": torch.tensor([32.0, 32.0])}
    bg = torch.tensor([0.2, 0.4, 0.6])
    out_ref = ref_renderer(assets, (64, 64), cam, bg)

    from exavatar_release_b200.renderer import GaussianRenderer
    mine = GaussianRenderer(rasterizer_cls=O.OracleRasterizer, settings_cls=O.OracleSettings)
    assets2 = {k: v.detach().clone().requires_grad_() for k, v in assets.items()}
    out_mine = mine(assets2, (64, 64), cam, bg)
    for k in ("img", "depthmap", "mask", "radius", "is_vis"):
        assert torch.equal(out_ref[k], out_mine[k]), k
    out_ref["img"].sum().backward()
    out_mine["img"].sum().backward()
    assert torch.equal(out_ref["mean_2d"].grad, out_mine["mean_2d"].grad)
    assert torch.equal(assets["mean_3d"].grad, assets2["mean_3d"].grad)
