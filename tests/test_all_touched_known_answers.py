"""Known answers for `all_touched=True`, derived BY HAND from the reference's walk
(rust/src/rasterization/burners.rs:94-247) and its two-pass polygon burn (burn_geometry.rs:214-238, writers.rs:15-60).
The reference pins all_touched only against live GDAL (python/test/test_many.py:237-262), which is not available
offline; these cases pin the oracle (CPU, every run) and the CUDA path (GPU runs) to the arithmetic written out below
instead of to each other.

Grid: 8 x 8, extent (0, 0, 8, 8), res 1  =>  pixel x = X, pixel y = 8 - Y (edges.rs:119-122).  Every coordinate is a
dyadic fraction, so every step below is exact in f64.  `P(x, y)` converts a pixel-space point to world coordinates.
Pixels are written (row, col).

Branches covered: vertical (with the 1e-4 end trim, both sides of it), horizontal (both directions), sloped walk
with slope 1/2, 2 and 1 (row steps and column steps), x- and y-clipping on all four sides, the polygon's line pass
+ fill pass with and without the PixelCache, per-part dedup for sum/count.  NOT covered (not hand-derivable without
an f64 emulator): the TOLERANCE = 1e-9 nudge of burners.rs:224-237 (a walk that lands within 1e-9 of a row border)
and negative slopes that hit it; those stay "oracle = literal code reading".
"""
import numpy as np
import pytest

import oracle
from oracle.wkt2wkb import wkt_to_wkb

N = 8
KW = dict(out_shape=(N, N), extent=(0, 0, N, N))


def P(x, y):
    return f"{x} {N - y}"


def line(*pts):
    return "LINESTRING (" + ", ".join(P(*p) for p in pts) + ")"


def raster(pixels, value=1, dtype="uint8", repeat=None):
    a = np.zeros((N, N), dtype)
    for r, c in pixels:
        a[r, c] = value
    for (r, c), k in (repeat or {}).items():
        a[r, c] = k
    return a


# ---------------------------------------------------------------------------------------------------------------------
# the cases: (name, WKT, expected pixel list)
# ---------------------------------------------------------------------------------------------------------------------
CASES = []

# A. vertical (|dx| < 0.01, burners.rs:125-149): ix = floor(2.3) = 2, iy = floor(1.2) = 1,
#    iy_end = floor(5.0 - 1e-4) = floor(4.9999) = 4: the endpoint sits exactly on the border of row 5 and the trim keeps
#    row 5 out.
CASES.append(("vertical_end_on_border", line((2.3, 1.2), (2.3, 5.0)), [(1, 2), (2, 2), (3, 2), (4, 2)]))
# A'. same, end at y = 5.0002: iy_end = floor(5.0002 - 0.0001) = floor(5.0001) = 5: row 5 is in.  Reversed vertex order
#     (burners.rs:127-129 swaps).
CASES.append(("vertical_past_the_trim", line((2.3, 5.0002), (2.3, 1.2)), [(1, 2), (2, 2), (3, 2), (4, 2), (5, 2)]))
# B. horizontal (|dy| < 0.01, :152-176): iy = floor(3.5) = 3, ix = floor(1.5) = 1, ix_end = floor(6.0 - 1e-4) = 5.
CASES.append(("horizontal", line((1.5, 3.5), (6.0, 3.5)), [(3, c) for c in range(1, 6)]))
#    ... and right-to-left: the swap at :118-121 makes it the same walk.
CASES.append(("horizontal_reversed", line((6.0, 3.5), (1.5, 3.5)), [(3, c) for c in range(1, 6)]))
# C. slope 1 through pixel corners (0,0) -> (4,4) (:178-241).  At (k,k): write (k,k); sx = floor(k+1) - k = 1, sy = 1,
#    floor(k + 1) != k  =>  slope > 0 branch: sy = (k+1) - k = 1, sx = sy / 1 = 1  =>  (k+1, k+1).  Stops at df_x = 4
#    (not < df_x_end).  Only the diagonal cells: touching a corner does not burn the neighbours.
CASES.append(("diagonal_through_corners", line((0, 0), (4, 4)), [(0, 0), (1, 1), (2, 2), (3, 3)]))
# D. slope 1/2: (0.5, 0.25) -> (4.5, 2.25).
#    (0.5,0.25) write (0,0); sx = .5, sy = .25, floor(.5) = 0 = iy          -> (1.0, 0.5)
#    (1.0,0.5)  write (0,1); sx = 1,  sy = .5,  floor(1.0) = 1 != 0: sy = 1 - .5 = .5, sx = .5/.5 = 1 -> (2.0, 1.0)
#    (2.0,1.0)  write (1,2); sx = 1,  sy = .5,  floor(1.5) = 1 = iy         -> (3.0, 1.5)
#    (3.0,1.5)  write (1,3); sx = 1,  sy = .5,  floor(2.0) = 2 != 1: sy = .5, sx = 1 -> (4.0, 2.0)
#    (4.0,2.0)  write (2,4) (4.0 < 4.5); sx = 1, sy = .5, floor(2.5) = 2    -> (5.0, 2.5) stop
CASES.append(("slope_half", line((0.5, 0.25), (4.5, 2.25)), [(0, 0), (0, 1), (1, 2), (1, 3), (2, 4)]))
# E. slope 2: (0.5, 0.5) -> (2.5, 4.5), inv_slope 0.5.
#    (0.5,0.5)  write (0,0); sx = .5,  sy = 1,   floor(1.5) = 1 != 0: sy = 1 - .5 = .5, sx = .25 -> (0.75, 1.0)
#    (0.75,1.0) write (1,0); sx = .25, sy = .5,  floor(1.5) = 1 = iy       -> (1.0, 1.5)
#    (1.0,1.5)  write (1,1); sx = 1,   sy = 2,   floor(3.5) = 3 != 1: sy = 2 - 1.5 = .5, sx = .25 -> (1.25, 2.0)
#    (1.25,2.0) write (2,1); sx = .75, sy = 1.5, floor(3.5) = 3 != 2: sy = 3 - 2 = 1,   sx = .5  -> (1.75, 3.0)
#    (1.75,3.0) write (3,1); sx = .25, sy = .5,  floor(3.5) = 3 = iy       -> (2.0, 3.5)
#    (2.0,3.5)  write (3,2); sx = 1,   sy = 2,   floor(5.5) = 5 != 3: sy = 4 - 3.5 = .5, sx = .25 -> (2.25, 4.0)
#    (2.25,4.0) write (4,2); sx = .75, sy = 1.5, floor(5.5) = 5 != 4: sy = 1, sx = .5 -> (2.75, 5.0) stop (>= 2.5)
CASES.append(("slope_two", line((0.5, 0.5), (2.5, 4.5)), [(0, 0), (1, 0), (1, 1), (2, 1), (3, 1), (3, 2), (4, 2)]))
# F. clipped on all four sides: (-1, -4) -> (9, 16), slope 20/10 = 2, inv_slope 0.5 (:182-207).
#    x start: df_y = -4 + (0 - -1)*2 = -2, df_x = 0.      x end: df_y_end = 16 + (8 - 9)*2 = 14, df_x_end = 8.
#    y start: df_y = -2 < 0: df_x = 0 + (0 - -2)*.5 = 1, df_y = 0.   y end: 14 > 8: df_x_end = 8 + (8 - 14)*.5 = 5.
#    walk from (1,0) while df_x < 5:
#    (1,0) w (0,1) -> sy = 1, sx = .5 -> (1.5,1) w (1,1) -> (2,2) w (2,2) -> (2.5,3) w (3,2) -> (3,4) w (4,3)
#    -> (3.5,5) w (5,3) -> (4,6) w (6,4) -> (4.5,7) w (7,4) -> (5,8) stop.
CASES.append(("clipped_four_sides", line((-1, -4), (9, 16)),
              [(0, 1), (1, 1), (2, 2), (3, 2), (4, 3), (5, 3), (6, 4), (7, 4)]))
# F'. slope 1 entering through the top-left corner region: (-2, -3) -> (10, 9).
#    x start: df_y = -3 + 2 = -1, df_x = 0;  x end: df_y_end = 9 + (8 - 10) = 7, df_x_end = 8.
#    y start: df_y < 0: df_x = 0 + 1 = 1, df_y = 0.  walk (1,0),(2,1),...: writes (r, r+1) for r = 0..6, stops at (8,7).
CASES.append(("clipped_slope_one", line((-2, -3), (10, 9)), [(r, r + 1) for r in range(7)]))
# a segment wholly outside is dropped by extract_line (edges.rs:124-132): max_x = -0.5 < 0
CASES.append(("outside", line((-3, 2), (-0.5, 6)), []))


@pytest.mark.parametrize("name,wkt,pixels", CASES, ids=[c[0] for c in CASES])
def test_oracle_walk_matches_hand_derivation(name, wkt, pixels):
    got = oracle.rusterize([wkt_to_wkb(wkt)], fun="any", dtype="uint8", background=0, all_touched=True, **KW)
    assert np.array_equal(got[0], raster(pixels)), name
    # the stream order is the walk's order (writers.rs:93-99)
    sp = oracle.rusterize([wkt_to_wkb(wkt)], fun="last", dtype="uint8", background=0, all_touched=True, encoding="sparse", **KW)
    assert list(zip(sp["rows"].tolist(), sp["cols"].tolist())) == pixels, name


# ---------------------------------------------------------------------------------------------------------------------
# polygon: line pass + fill pass
# ---------------------------------------------------------------------------------------------------------------------
# Square with pixel-space corners (1.5,1.5) (4.5,1.5) (4.5,4.5) (1.5,4.5).
#   standard fill (burners.rs:261-320): vertical edges x = 1.5 and x = 4.5, ystart = ceil(1.5 - .5) = 1,
#   yend = ceil(4.5 - .5) = 4: rows 1..3; cols [floor(1.5 + .5), floor(4.5 + .5)) = 2..4        -> 9 pixels.
#   line pass over the closed ring (top, right, bottom, left):
#     top    (1.5,1.5)->(4.5,1.5) horizontal: row 1, cols floor(1.5) .. floor(4.5 - 1e-4) = 1..4
#     right  (4.5,1.5)->(4.5,4.5) vertical:   col 4, rows 1 .. floor(4.4999) = 1..4
#     bottom (4.5,4.5)->(1.5,4.5) horizontal: row 4, cols 1..4
#     left   (1.5,4.5)->(1.5,1.5) vertical:   col 1, rows 1..4
#   union = the 4 x 4 block rows 1..4 x cols 1..4; fill \ walk = {(2,2),(2,3),(3,2),(3,3)}.
SQUARE = "POLYGON ((" + ", ".join(P(*p) for p in [(1.5, 1.5), (4.5, 1.5), (4.5, 4.5), (1.5, 4.5), (1.5, 1.5)]) + "))"
BLOCK = [(r, c) for r in range(1, 5) for c in range(1, 5)]
FILL = [(r, c) for r in range(1, 4) for c in range(2, 5)]
WALK_WRITES = ([(1, c) for c in range(1, 5)] + [(r, 4) for r in range(1, 5)] + [(4, c) for c in range(1, 5)] +
               [(r, 1) for r in range(1, 5)])
# with the PixelCache (sum / count): first visits only, then the fill pixels the walk has not touched
DEDUP_STREAM = [(1, 1), (1, 2), (1, 3), (1, 4), (2, 4), (3, 4), (4, 4), (4, 1), (4, 2), (4, 3), (2, 1), (3, 1),
                (2, 2), (2, 3), (3, 2), (3, 3)]


def test_oracle_polygon_two_passes():
    g = [wkt_to_wkb(SQUARE)]
    std = oracle.rusterize(g, fun="count", dtype="uint8", background=0, all_touched=False, **KW)
    assert np.array_equal(std[0], raster(FILL))
    # count / sum: REQUIRES_DEDUP (prelude.rs:116-118): every pixel of the block exactly once
    cnt = oracle.rusterize(g, fun="count", dtype="uint8", background=0, all_touched=True, **KW)
    assert np.array_equal(cnt[0], raster(BLOCK))
    sm = oracle.rusterize(g, burn=7, fun="sum", dtype="float32", background=np.nan, all_touched=True, **KW)
    exp = np.full((N, N), np.nan, np.float32)
    for r, c in BLOCK:
        exp[r, c] = 7
    assert np.array_equal(sm[0], exp, equal_nan=True)
    sp = oracle.rusterize(g, burn=7, fun="sum", dtype="float32", background=np.nan, all_touched=True, encoding="sparse", **KW)
    assert list(zip(sp["rows"].tolist(), sp["cols"].tolist())) == DEDUP_STREAM
    # last (no cache, square pixels): all 16 walk writes, corners twice, then the 9 fill writes
    sp2 = oracle.rusterize(g, fun="last", dtype="uint8", background=0, all_touched=True, encoding="sparse", **KW)
    assert list(zip(sp2["rows"].tolist(), sp2["cols"].tolist())) == WALK_WRITES + FILL


# An L-shaped path that visits pixel (2,3) twice: (0.5,2.5) -> (3.5,2.5) -> (3.5,0.5).
#   horizontal: row 2, cols 0 .. floor(3.4999) = 0..3;  vertical (swapped to go down): col 3, rows 0 .. floor(2.4999) = 0..2.
#   sum / count dedup per part (burn_geometry.rs:179): (2,3) counts once; `last`-style functions do not care.
ELL = line((0.5, 2.5), (3.5, 2.5), (3.5, 0.5))
ELL_PIXELS = [(2, 0), (2, 1), (2, 2), (2, 3), (0, 3), (1, 3)]


def test_oracle_line_dedup_under_all_touched():
    g = [wkt_to_wkb(ELL)]
    cnt = oracle.rusterize(g, fun="count", dtype="uint8", background=0, all_touched=True, **KW)
    assert np.array_equal(cnt[0], raster(ELL_PIXELS))
    sp = oracle.rusterize(g, burn=5, fun="sum", dtype="int32", background=0, all_touched=True, encoding="sparse", **KW)
    assert list(zip(sp["rows"].tolist(), sp["cols"].tolist())) == ELL_PIXELS and set(sp["data"].tolist()) == {5}
    # two separate geometries covering the same pixel do add up (the cache is per geometry)
    two = oracle.rusterize([wkt_to_wkb(line((0.5, 2.5), (3.5, 2.5))), wkt_to_wkb(line((3.5, 2.5), (3.5, 0.5)))],
                           fun="count", dtype="uint8", background=0, all_touched=True, **KW)
    assert np.array_equal(two[0], raster(ELL_PIXELS, repeat={(2, 3): 2}))


# ---------------------------------------------------------------------------------------------------------------------
# the same answers from the CUDA path (dense and sparse, both polygon engines are irrelevant here: all_touched jobs
# take the record pipeline)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_matches_hand_derivations():
    from rusterize_b200 import core

    def burn(wkts, fun, dtype, bg, burn_v=1, sparse=False):
        g = core.Geoms.from_wkb([wkt_to_wkb(w) for w in wkts])
        ri = core.raster_info(None, shape=(N, N), extent=(0, 0, N, N))
        if sparse:
            sp = core.rasterize_sparse(g, ri, fun, dtype, burn_v, background=bg, all_touched=True)
            return list(zip(sp["rows"].tolist(), sp["cols"].tolist()))
        return core.rasterize_dense(g, ri, fun, dtype, burn_v, background=bg, all_touched=True)[0][0]

    for name, wkt, pixels in CASES:
        assert np.array_equal(burn([wkt], "any", "uint8", 0), raster(pixels)), name
        assert burn([wkt], "last", "uint8", 0, sparse=True) == pixels, name
    # all cases in one job, as separate geometries: count adds up where their pixel sets overlap
    total = np.zeros((N, N), np.uint8)
    for _, _, pixels in CASES:
        total += raster(pixels)
    assert np.array_equal(burn([c[1] for c in CASES], "count", "uint8", 0), total)
    assert np.array_equal(burn([SQUARE], "count", "uint8", 0), raster(BLOCK))
    assert burn([SQUARE], "sum", "float32", np.nan, 7, sparse=True) == DEDUP_STREAM
    assert burn([SQUARE], "last", "uint8", 0, sparse=True) == WALK_WRITES + FILL
    assert np.array_equal(burn([ELL], "count", "uint8", 0), raster(ELL_PIXELS))
    assert burn([ELL], "sum", "int32", 0, 5, sparse=True) == ELL_PIXELS
