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
    assert lib.b2r_abi_version() == 2
    for idx, cls in enumerate((L.B2RScene, L.B2RStatus, L.B2RWorkspace, L.B2RForwardOutputs, L.B2RBackwardArgs)):
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
