"""Row-interval tile culling (exavatar_release_b200/csrc/common.cuh: row_cull_setup / row_kept_columns) restated in numpy
float32 and checked against the per-tile region test (region_max_p2) it replaced: on random splats -- sizes from 0.4 to
60 px, anisotropy up to 50:1, any orientation, centres inside and just outside the image -- the rows must keep EVERY tile
the per-tile test keeps (a missed tile would drop visible contributions) and only marginally more (the 0.01 px pad).

CPU-only: this pins the closed form; that the CUDA kernels implement it is covered by the GPU parity tests (culled lists
are ordered subsets of the oracle's, images and gradients match, projection count == scatter count or slots stay empty).
"""
import numpy as np

f32 = np.float32
TILE = 16
ROW_PAD = f32(0.01)
LOG2E = 1.4426950408889634


def region_max_p2(sx, sy, A2, B2, C2, x0, y0, x1, y1):
    """common.cuh region_max_p2: maximum of the concave exponent over the continuous rect [x0,x1] x [y0,y1]."""
    lx, hx = sx - x1, sx - x0
    ly, hy = sy - y1, sy - y0
    in_x = (lx <= 0) & (hx >= 0)
    in_y = (ly <= 0) & (hy >= 0)
    best = np.full(np.broadcast(sx, x0).shape, -np.inf, dtype=f32)
    ex = np.where(lx > 0, lx, hx)
    with np.errstate(all="ignore"):
        dy = np.clip(-B2 * ex / (f32(2) * C2), ly, hy)
        vx = A2 * ex * ex + B2 * ex * dy + C2 * dy * dy
        ey = np.where(ly > 0, ly, hy)
        dx = np.clip(-B2 * ey / (f32(2) * A2), lx, hx)
        vy = A2 * dx * dx + B2 * dx * ey + C2 * ey * ey
    best = np.where(~in_x, np.maximum(best, vx), best)
    best = np.where(~in_y, np.maximum(best, vy), best)
    return np.where(in_x & in_y, f32(0), best)


def row_kept_columns(sx, sy, A2, B2, C2, thr2, tys, x0, x1, W, H):
    """common.cuh row_cull_setup + row_kept_columns for the tile rows `tys`; returns (tlo, thi), empty when tlo > thi."""
    with np.errstate(all="ignore"):
        D = B2 * B2 - f32(4) * A2 * C2
        k = f32(4) * thr2 / (-D)
        dy_ext = np.sqrt(A2 * k)
        dx_top = np.sqrt(C2 * k)
        dys = -B2 / (f32(2) * C2) * dx_top
        inv2A = f32(1) / (f32(2) * A2)
        fA4t = f32(4) * A2 * thr2
        y0 = (tys * TILE).astype(f32)
        y1 = np.minimum(y0 + (TILE - 1), H - 1).astype(f32)
        lo = np.maximum(sy - y1, -dy_ext)
        hi = np.minimum(sy - y0, dy_ext)
        empty = ~(lo <= hi)
        dyr = np.minimum(np.maximum(dys, lo), hi)
        dyl = np.minimum(np.maximum(-dys, lo), hi)
        disc_r = np.maximum(D * dyr * dyr + fA4t, f32(0))
        disc_l = np.maximum(D * dyl * dyl + fA4t, f32(0))
        dx_hi = (-B2 * dyr - np.sqrt(disc_r)) * inv2A
        dx_lo = (-B2 * dyl + np.sqrt(disc_l)) * inv2A
        xl = sx - dx_hi - ROW_PAD
        xr = sx - dx_lo + ROW_PAD
        empty |= ~(xl <= f32(W - 1))
        tlo = np.maximum(x0, np.ceil((xl - f32(TILE - 1)) * f32(1.0 / TILE)))
        thi = np.minimum(x1 - 1, np.floor(xr * f32(1.0 / TILE)))
    tlo = np.where(empty, x0, tlo)
    thi = np.where(empty, x0 - 1, thi)
    return tlo.astype(np.int64), thi.astype(np.int64)


def _random_splats(n, W, H, seed):
    rng = np.random.default_rng(seed)
    sx = rng.uniform(-20, W + 20, n).astype(f32)
    sy = rng.uniform(-20, H + 20, n).astype(f32)
    s1 = np.exp(rng.uniform(np.log(0.4), np.log(60), n))
    s2 = s1 * np.exp(rng.uniform(np.log(0.02), 0, n))
    th = rng.uniform(0, np.pi, n)
    c, s = np.cos(th), np.sin(th)
    cxx = c * c * s1 ** 2 + s * s * s2 ** 2 + 0.3
    cyy = s * s * s1 ** 2 + c * c * s2 ** 2 + 0.3
    cxy = c * s * (s1 ** 2 - s2 ** 2)
    det = cxx * cyy - cxy ** 2
    A2 = (-0.5 * LOG2E * cyy / det).astype(f32)
    B2 = (-LOG2E * (-cxy / det)).astype(f32)
    C2 = (-0.5 * LOG2E * cxx / det).astype(f32)
    op = rng.uniform(0.002, 1, n).astype(f32)  # includes opacities below 1/255: thr2 > 0, nothing is kept
    thr2 = (-np.log2(255 * op) - 0.02).astype(f32)
    lam = 0.5 * (cxx + cyy) + np.sqrt(np.maximum(0.1, (0.5 * (cxx + cyy)) ** 2 - det))
    rad = np.ceil(3 * np.sqrt(lam)).astype(np.int64)
    return sx, sy, A2, B2, C2, thr2, rad


def _compare(W, H, n, seed):
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    sx, sy, A2, B2, C2, thr2, rad = _random_splats(n, W, H, seed)
    x0 = np.clip(((sx - rad) / TILE).astype(np.int64), 0, gx)
    y0 = np.clip(((sy - rad) / TILE).astype(np.int64), 0, gy)
    x1 = np.clip(((sx + rad + TILE - 1) / TILE).astype(np.int64), 0, gx)
    y1 = np.clip(((sy + rad + TILE - 1) / TILE).astype(np.int64), 0, gy)
    exact_pairs = row_pairs = missed = 0
    for i in range(n):
        if (x1[i] - x0[i]) * (y1[i] - y0[i]) == 0:
            continue
        txs, tys = np.arange(x0[i], x1[i]), np.arange(y0[i], y1[i])
        TX, TY = np.meshgrid(txs, tys)
        rx0, ry0 = (TX * TILE).astype(f32), (TY * TILE).astype(f32)
        rx1 = np.minimum(rx0 + (TILE - 1), W - 1).astype(f32)
        ry1 = np.minimum(ry0 + (TILE - 1), H - 1).astype(f32)
        exact = ~(region_max_p2(sx[i], sy[i], A2[i], B2[i], C2[i], rx0, ry0, rx1, ry1) < thr2[i])
        tlo, thi = row_kept_columns(sx[i], sy[i], A2[i], B2[i], C2[i], thr2[i], tys, x0[i], x1[i], W, H)
        rows = (TX >= tlo[:, None]) & (TX <= thi[:, None])
        exact_pairs += int(exact.sum())
        row_pairs += int(rows.sum())
        missed += int((exact & ~rows).sum())
    return exact_pairs, row_pairs, missed


def test_rows_keep_every_tile_the_region_test_keeps():
    for (W, H, n, seed) in ((512, 512, 12000, 0), (500, 300, 6000, 1)):  # the second size has partial last tiles
        exact_pairs, row_pairs, missed = _compare(W, H, n, seed)
        assert exact_pairs > 50000
        assert missed == 0
        assert row_pairs - exact_pairs <= 2e-3 * exact_pairs  # only the pad: borderline tiles


def test_degenerate_thresholds():
    one = lambda v: np.array([v], dtype=f32)
    tys = np.arange(0, 4)
    # opacity below 1/255 (thr2 > 0) and opacity 0 (thr2 = +inf): no tile
    for t in (0.5, np.inf):
        tlo, thi = row_kept_columns(one(30), one(30), one(-0.1), one(0.0), one(-0.1), one(t), tys, 0, 4, 64, 64)
        assert (tlo > thi).all()
    # a sharp splat in the middle of tile (1, 1): exactly that tile
    tlo, thi = row_kept_columns(one(24), one(24), one(-2.0), one(0.0), one(-2.0), one(-6.0), tys, 0, 4, 64, 64)
    assert [(int(a), int(b)) for a, b in zip(tlo, thi)] == [(0, -1), (1, 1), (0, -1), (0, -1)]
