"""The C-ABI shared library: loads without a GPU, exports every symbol include/b200raster.h declares, struct layouts
agree with the ctypes mirror, argument validation returns the documented codes before any CUDA call."""
import ctypes as C
import os
import re

import pytest

from exavatar_release_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200raster.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2r_[a-z_0-9A-Z]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(L.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in b200raster.h but not exported"
    # and the binding knows all of them
    bound = {s[0] for s in L.SYMBOLS}
    assert set(names) <= bound, set(names) - bound


def test_struct_layouts_and_version():
    lib = L.load()
    assert lib.b2r_abi_version() == 3 == L.ABI_VERSION
    for idx, cls in enumerate((L.B2RScene, L.B2RStatus, L.B2RWorkspace, L.B2RForwardOutputs, L.B2RBackwardArgs, L.B2RView)):
        assert lib.b2r_sizeof(idx) == C.sizeof(cls)
    assert lib.b2r_sizeof(99) == 0
    assert C.sizeof(L.B2RStatus) == 64


def test_size_queries():
    lib = L.load()
    a = lib.b2r_ctx_bytes(1000, 64, 64)
    b = lib.b2r_ctx_bytes(2000, 64, 64)
    c = lib.b2r_ctx_bytes(1000, 128, 128)
    assert 0 < a < b and a < c and a % 256 == 0
    assert lib.b2r_scratch_bytes(1000, 64, 64, 0) >= 8
    assert lib.b2r_scratch_bytes(1000, 64, 64, 1 << 20) >= 8 << 20
    assert lib.b2r_backward_scratch_bytes(1000) >= 48000
    assert lib.b2r_ctx_bytes(0, 16, 16) > 0
    # checkpoint store: a segment table + 6 KB per 512-entry cut; grows with the duplicate capacity and the tile count
    c0 = lib.b2r_checkpoint_bytes(64, 64, 0)
    c1 = lib.b2r_checkpoint_bytes(64, 64, 1 << 20)
    assert 0 < c0 < c1 and c1 >= (1 << 20) // 512 * 6144
    assert lib.b2r_checkpoint_bytes(512, 512, 0) > c0


def test_error_codes_without_touching_cuda():
    lib = L.load()
    assert lib.b2r_strerror(0) == b"ok"
    assert b"invalid" in lib.b2r_strerror(-1)
    assert b"workspace" in lib.b2r_strerror(-2)
    sc = L.B2RScene()
    ws = L.B2RWorkspace()
    out = L.B2RForwardOutputs()
    # null scene / zero-sized image
    assert lib.b2r_forward(None, C.byref(ws), C.byref(out), None) == -1
    sc.P, sc.width, sc.height = 10, 0, 32
    assert lib.b2r_forward(C.byref(sc), C.byref(ws), C.byref(out), None) == -1
    sc.width, sc.tanfovx, sc.tanfovy = 32, 0.5, 0.5
    assert lib.b2r_forward(C.byref(sc), C.byref(ws), C.byref(out), None) == -1  # matrices missing
    fake = 0x1000  # never dereferenced on the host
    sc.bg = sc.viewmatrix = sc.projmatrix = sc.campos = fake
    sc.means3D = sc.opacities = fake
    assert lib.b2r_forward(C.byref(sc), C.byref(ws), C.byref(out), None) == -1  # neither shs nor colours
    sc.colors_precomp = fake
    sc.shs = fake
    assert lib.b2r_forward(C.byref(sc), C.byref(ws), C.byref(out), None) == -1  # both
    sc.shs = None
    assert lib.b2r_forward(C.byref(sc), C.byref(ws), C.byref(out), None) == -1  # no covariance source
    sc.scales = sc.rotations = fake
    assert lib.b2r_forward(C.byref(sc), C.byref(ws), C.byref(out), None) == -1  # ctx missing
    ws.ctx, ws.ctx_bytes = fake, 16
    assert lib.b2r_forward(C.byref(sc), C.byref(ws), C.byref(out), None) == -2  # ctx too small
    assert lib.b2r_mark_visible(-1, None, None, None, None) == -1
    assert lib.b2r_launch_count() == 0  # nothing was launched by any of the above


def test_kernel_names():
    lib = L.load()
    names = [lib.b2r_kernel_name(i).decode() for i in range(9)]
    assert names[0] == "project" and names[5] == "composite_fwd" and names[6] == "composite_bwd"
    assert lib.b2r_kernel_name(99) == b"?"


def test_error_codes_of_the_round1_additions_without_touching_cuda():
    """Validation of the fields added in ABI v2 (fused skinning, detached prefix, SH row limit, id width) happens on the
    host before any launch, so it is testable without a GPU."""
    lib = L.load()
    fake = 0x1000
    sc = L.B2RScene()
    sc.P, sc.width, sc.height, sc.tanfovx, sc.tanfovy = 10, 32, 32, 0.5, 0.5
    sc.bg = sc.viewmatrix = sc.projmatrix = sc.campos = fake
    sc.opacities = sc.colors_precomp = sc.scales = sc.rotations = fake
    ws = L.B2RWorkspace()
    ws.ctx, ws.ctx_bytes = fake, 16  # too small on purpose: a scene that validates reaches the workspace check (-2)
    out = L.B2RForwardOutputs()
    fwd = lambda: lib.b2r_forward(C.byref(sc), C.byref(ws), C.byref(out), None)
    assert fwd() == -1                      # neither means3D nor skinning
    sc.means3D = fake
    assert fwd() == -2
    # fused skinning replaces means3D, but needs all of its inputs and a sane joint count
    sc.means3D = None
    sc.skin_xyz = fake
    assert fwd() == -1
    sc.skin_weights = sc.skin_joint_mats = sc.skin_trans = fake
    sc.skin_J = 0
    assert fwd() == -1
    sc.skin_J = 65
    assert fwd() == -1
    sc.skin_J = 55
    assert fwd() == -2
    sc.skin_cam_Rinv = fake                 # camera rotation without its translation
    assert fwd() == -1
    sc.skin_cam_t = fake
    assert fwd() == -2
    # SH rows are staged through shared memory: at most 16 coefficients, degree <= 3, enough coefficients for the degree
    sc.colors_precomp = None
    sc.shs = fake
    sc.sh_degree, sc.sh_coeffs = 3, 16
    assert fwd() == -2
    sc.sh_coeffs = 9
    assert fwd() == -1
    sc.sh_degree, sc.sh_coeffs = 1, 17
    assert fwd() == -1
    sc.sh_degree, sc.sh_coeffs = 4, 25
    assert fwd() == -1
    sc.shs, sc.colors_precomp, sc.sh_degree, sc.sh_coeffs = None, fake, 0, 0
    # the splat record carries the Gaussian id in 29 bits
    sc.P = 1 << 29
    assert fwd() == -1
    sc.P = 10
    # backward: colour gradient and scratch are mandatory, the detached prefix cannot exceed P
    args = L.B2RBackwardArgs()
    ws.ctx_bytes = lib.b2r_ctx_bytes(10, 32, 32)
    bwd = lambda scratch, nbytes: lib.b2r_backward(C.byref(sc), C.byref(ws), C.byref(args), scratch, nbytes, None)
    assert bwd(fake, 1 << 20) == -1         # no dL_dcolor
    args.dL_dcolor = fake
    assert bwd(None, 1 << 20) == -1
    assert bwd(fake, 8) == -2               # scratch too small
    args.first_row = 11
    assert bwd(fake, 1 << 20) == -1


def test_error_codes_of_the_abi_v3_entry_points_without_touching_cuda():
    """Views, the split pipeline stages, the checkpoint store and the posed-position gradient are validated on the host."""
    lib = L.load()
    fake = 0x1000
    sc = L.B2RScene()
    sc.P, sc.width, sc.height, sc.tanfovx, sc.tanfovy = 10, 32, 32, 0.5, 0.5
    sc.bg = sc.viewmatrix = sc.projmatrix = sc.campos = fake
    sc.means3D = sc.opacities = sc.colors_precomp = sc.scales = sc.rotations = fake
    ws = L.B2RWorkspace()
    ws.ctx, ws.ctx_bytes = fake, lib.b2r_ctx_bytes(10, 32, 32)
    ws.dup_ids, ws.dup_capacity = fake, 1000
    out = L.B2RForwardOutputs()
    view = L.B2RView()
    comp = lambda: lib.b2r_forward_composite(C.byref(sc), C.byref(ws), C.byref(view), C.byref(out), None)
    assert comp() == -1                      # no output images
    out.color = out.depth = out.alpha = fake
    view.id_begin, view.id_end = 4, 2
    assert comp() == -1                      # empty / inverted Gaussian range
    view.id_begin, view.id_end = 0, 11
    assert comp() == -1                      # range beyond P
    view.id_begin, view.id_end = 2, 10
    view.final_T = fake
    assert comp() == -1                      # final_T and n_contrib come as a pair
    view.final_T = None
    # an undersized checkpoint store is rejected before anything is launched
    ws.checkpoints, ws.checkpoint_bytes = fake, 16
    assert comp() == -2
    ws.checkpoint_bytes = lib.b2r_checkpoint_bytes(32, 32, 1000)
    view.checkpoints, view.checkpoint_bytes = fake, 16
    assert comp() == -2                      # ... and so is a view's own store
    # binning needs the scratch
    assert lib.b2r_forward_bin(C.byref(sc), C.byref(ws), None) == -1
    ws.scratch, ws.scratch_bytes = fake, 8
    assert lib.b2r_forward_bin(C.byref(sc), C.byref(ws), None) == -2
    # backward stages
    args = L.B2RBackwardArgs()
    view = L.B2RView()
    view.id_begin, view.id_end = 0, 10
    bc = lambda scratch, n: lib.b2r_backward_composite(C.byref(sc), C.byref(ws), C.byref(view), C.byref(args), scratch, n, None)
    bp = lambda scratch, n: lib.b2r_backward_project(C.byref(sc), C.byref(ws), C.byref(args), scratch, n, None)
    assert bc(fake, 1 << 20) == -1           # no dL_dcolor
    args.dL_dcolor = fake
    assert bc(fake, 8) == -2 and bp(fake, 8) == -2
    assert bp(None, 1 << 20) == -1
    args.first_row = 11
    assert bc(fake, 1 << 20) == -1 and bp(fake, 1 << 20) == -1
    args.first_row = 0
    args.dL_dposed = fake                    # a posed-position gradient only makes sense with fused skinning
    assert bp(fake, 1 << 20) == -1
    assert lib.b2r_backward(C.byref(sc), C.byref(ws), C.byref(args), fake, 1 << 20, None) == -1


def test_compiled_torch_binding_builds_and_loads():
    """csrc_torch/b2r_torch.cpp (the eager call's host path as a C++ autograd Function) builds in-tree against this
    interpreter's torch, links to libb200raster.so and reports the same ABI version.  No compute here (no GPU)."""
    from exavatar_release_b200 import build_ext, rasterizer
    path = build_ext.build_torch_ext()
    assert os.path.exists(path)
    ext = rasterizer._compiled_binding()
    assert ext and ext.abi_version() == L.ABI_VERSION
    import torch
    a = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="CUDA tensor"):  # no CPU fallback in the compiled route either
        ext.rasterize(a, a, None, a, torch.zeros(4, 1), a, torch.zeros(4, 4), None, 16, 16, 1.0, 1.0, torch.zeros(3), 1.0,
                      torch.eye(4), torch.eye(4), 0, torch.zeros(3), True, True, 1.25, True)
