"""The line kernel emits pixel k of a segment in closed form instead of walking the reference's
Bresenham loop (rust/src/rasterization/burners.rs:54-84).  This restates both in Python and checks
them against each other exhaustively on a small lattice, including the window clipping formulas of
rz_kernels.cuh (clip_linear / clip_stepped)."""
import itertools


def reference_walk(ix0, iy0, ix1, iy1):
    dx, dy = abs(ix1 - ix0), -abs(iy1 - iy0)
    sx = 1 if ix0 < ix1 else -1
    sy = 1 if iy0 < iy1 else -1
    err = dx + dy
    out = []
    while ix0 != ix1 or iy0 != iy1:
        out.append((ix0, iy0))
        e2 = 2 * err
        if e2 >= dy:
            err += dy
            ix0 += sx
        if e2 <= dx:
            err += dx
            iy0 += sy
    return out


def closed_form(ix0, iy0, ix1, iy1):
    dx, dy = abs(ix1 - ix0), abs(iy1 - iy0)
    sx = 1 if ix0 < ix1 else -1
    sy = 1 if iy0 < iy1 else -1
    xmajor = dx >= dy
    dmaj, dmin = (dx, dy) if xmajor else (dy, dx)
    out = []
    for k in range(dmaj):
        q = (2 * dmin * k + dmaj) // (2 * dmaj)
        out.append((ix0 + sx * k, iy0 + sy * q) if xmajor else (ix0 + sx * q, iy0 + sy * k))
    return out, (dmaj, dmin, sx, sy, xmajor)


def ceil_div(a, b):
    return -((-a) // b)


def clip(ix0, iy0, par, c_lo, c_hi, r_lo, r_hi):
    """k-range whose pixels fall in cols [c_lo,c_hi) x rows [r_lo,r_hi) — mirrors line_setup()."""
    dmaj, dmin, sx, sy, xmajor = par
    klo, khi = 0, dmaj - 1

    def linear(c0, s, lo, hi, klo, khi):
        if s > 0:
            return max(klo, lo - c0), min(khi, hi - 1 - c0)
        return max(klo, c0 - hi + 1), min(khi, c0 - lo)

    def stepped(c0, s, lo, hi, klo, khi):
        qa, qb = (lo - c0, hi - 1 - c0) if s > 0 else (c0 - hi + 1, c0 - lo)
        if dmin == 0:
            return (klo, klo - 1) if (qa > 0 or qb < 0) else (klo, khi)
        if qa > 0:
            klo = max(klo, ceil_div(2 * dmaj * qa - dmaj, 2 * dmin))
        if qb < 0:
            return klo, klo - 1
        if qb < dmin:
            khi = min(khi, ceil_div(2 * dmaj * (qb + 1) - dmaj, 2 * dmin) - 1)
        return klo, khi

    if xmajor:
        klo, khi = linear(ix0, sx, c_lo, c_hi, klo, khi)
        klo, khi = stepped(iy0, sy, r_lo, r_hi, klo, khi)
    else:
        klo, khi = linear(iy0, sy, r_lo, r_hi, klo, khi)
        klo, khi = stepped(ix0, sx, c_lo, c_hi, klo, khi)
    return klo, khi


def test_closed_form_equals_reference_walk_exhaustive():
    R = range(-9, 10)
    for ix1, iy1 in itertools.product(R, R):
        got, _ = closed_form(0, 0, ix1, iy1)
        assert got == reference_walk(0, 0, ix1, iy1), (ix1, iy1)
    for x0, y0, x1, y1 in [(3, -7, 40, 5), (-20, 11, 13, -30), (5, 5, 5, -60), (-3, 2, 97, 3),
                           (1000, -1000, -999, 1001), (7, 7, 7, 7), (0, 0, 64, 32), (0, 0, 33, 64)]:
        assert closed_form(x0, y0, x1, y1)[0] == reference_walk(x0, y0, x1, y1)
        assert closed_form(x1, y1, x0, y0)[0] == reference_walk(x1, y1, x0, y0)


def test_window_clipping_selects_exactly_the_inside_pixels():
    windows = [(0, 6, 0, 6), (2, 5, 1, 3), (0, 1, 0, 1), (3, 9, -2, 2)]
    R = range(-8, 9)
    for (ix0, iy0) in [(0, 0), (4, -3), (-5, 7)]:
        for ix1, iy1 in itertools.product(R, R):
            px, par = closed_form(ix0, iy0, ix1, iy1)
            for c_lo, c_hi, r_lo, r_hi in windows:
                klo, khi = clip(ix0, iy0, par, c_lo, c_hi, r_lo, r_hi)
                inside = [k for k, (x, y) in enumerate(px) if c_lo <= x < c_hi and r_lo <= y < r_hi]
                got = list(range(klo, khi + 1)) if khi >= klo else []
                assert got == inside, (ix0, iy0, ix1, iy1, c_lo, c_hi, r_lo, r_hi)
