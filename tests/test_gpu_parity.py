"""GPU parity: the CUDA burn path, called through the C ABI (rusterize_b200.core -> librz_b200.so),
against the CPU oracle on identical inputs.  Bit-exact for every dtype and pixel function —
including floating-point `sum`, because the fill kernel replays writes in the reference's order."""
import numpy as np
import pytest
from PIL import Image

import oracle
import synth
from cases import GEOMS, GEOMS_EXPLODED, R_INFO, VALUES, VALUES_EXPLODED, rmat, sq
from rusterize_b200 import core

pytestmark = pytest.mark.gpu
GOLDEN = __import__("os").path.join(__import__("os").path.dirname(__file__), "golden")


def both(geoms, fun="last", dtype="float64", burn=1, bg=0, by=None, field_valid=None, rows=None, tile_bytes=0,
         all_touched=False, flags=0, **kw):
    og = oracle.Geoms.from_any(geoms)
    ori = oracle.raster_info(og, **kw)
    exp, names = oracle.rasterize_dense(og, ori, fun, dtype, burn, field_valid, by, bg, all_touched)
    g = core.Geoms.from_any(geoms)
    ri = core.raster_info(g, **kw)
    band, nb = None, 1
    if by is not None:
        band, bn = core.group_keys(by)
        nb = len(bn)
        assert bn == names
    got, st = core.rasterize_dense(g, ri, fun, dtype, burn, field_valid, band, nb, bg, all_touched, rows=rows,
                                   tile_bytes=tile_bytes, flags=flags)
    if rows is not None:
        exp = exp[:, rows[0]:rows[1]]
    return exp, got, st


def assert_same(exp, got):
    assert exp.shape == got.shape and exp.dtype == got.dtype
    assert np.array_equal(exp, got, equal_nan=exp.dtype.kind == "f"), f"{np.sum(exp != got)} pixels differ"


def test_golden_tifs():
    exp, got, _ = both(GEOMS, "sum", "uint8", VALUES, 0, resolution=(1, 1))
    assert_same(exp, got)
    assert np.array_equal(got[0], np.array(Image.open(f"{GOLDEN}/standard_output_sum.tif")))
    exp, got, _ = both(GEOMS_EXPLODED, "sum", "uint8", VALUES_EXPLODED, 0, shape=(47, 319))
    assert_same(exp, got)
    assert np.array_equal(got[0], np.array(Image.open(f"{GOLDEN}/standard_output_sum_custom_shape.tif")))


@pytest.mark.parametrize("fun,vals,expect", [
    ("sum", [5, 7], 12), ("min", [5, 7], 5), ("max", [5, 7], 7), ("first", [5, 7], 5), ("last", [5, 7], 7),
    ("count", [1, 1], 2), ("any", [5, 7], 1)])
def test_r_known_answers(fun, vals, expect):
    two = [sq(0, 0, 4, 4), sq(0, 0, 4, 4)]
    exp, got, _ = both(two, fun, "float64", np.array(vals, float), 0.0, shape=(4, 4), extent=(0, 0, 4, 4))
    assert_same(exp, got)
    assert np.array_equal(got[0], np.full((4, 4), float(expect)))


def test_r_geometry_cases():
    kw = dict(shape=(4, 4), extent=(0, 0, 4, 4))
    _, got, _ = both(["LINESTRING (0 0, 4 4)"], burn=9.0, **kw)
    assert np.array_equal(got[0], rmat([0, 0, 0, 0, 0, 0, 0, 9, 0, 0, 9, 0, 0, 9, 0, 0]))
    mp = "MULTIPOLYGON (((0 0, 2 0, 2 2, 0 2, 0 0)), ((2 2, 4 2, 4 4, 2 4, 2 2)))"
    _, got, _ = both([mp], burn=9.0, **kw)
    assert np.array_equal(got[0], rmat([0, 0, 9, 9, 0, 0, 9, 9, 9, 9, 0, 0, 9, 9, 0, 0]))
    exp, got, _ = both([sq(0, 0, 2, 4), sq(2, 0, 4, 4)], burn=np.array([10.0, 20.0]), by=["a", "b"], **kw)
    assert_same(exp, got)
    assert np.array_equal(got[0], rmat([10] * 8 + [0] * 8)) and np.array_equal(got[1], rmat([0] * 8 + [20] * 8))


@pytest.mark.parametrize("dtype", oracle.DTYPES)
@pytest.mark.parametrize("fun", oracle.FUNS)
def test_all_dtypes_and_functions_mixed_geometries(dtype, fun):
    seed = oracle.DTYPES.index(dtype) * 7 + oracle.FUNS.index(fun)
    geoms = synth.mixed_geometries(seed, 150, 300, 200, rho=20.0)
    rng = np.random.default_rng(seed)
    dt = np.dtype(dtype)
    # values collide with the background and (for floats) include NaN: exercises the sentinel quirks
    burn = rng.integers(0, 5, len(geoms)).astype(dt)
    bg = 0
    if dt.kind == "f":
        burn[rng.random(len(geoms)) < 0.1] = np.nan
        bg = np.nan if seed % 2 else 2.0
    elif seed % 3 == 0:
        bg = 2
    exp, got, _ = both(geoms, fun, dtype, burn, bg, shape=(200, 300), extent=(0, 0, 300, 200))
    assert_same(exp, got)


def test_integer_wraparound_and_count_restart():
    many = [sq(0, 0, 4, 4)] * 300
    for fun, dtype, bg in [("sum", "uint8", 0), ("sum", "int8", 5), ("count", "uint8", 0), ("count", "uint8", 7),
                           ("sum", "int16", -3)]:
        exp, got, _ = both(many, fun, dtype, np.arange(300) % 120, bg, shape=(4, 4), extent=(0, 0, 4, 4))
        assert_same(exp, got)


def test_by_bands_lexicographic_and_null_fields():
    geoms = synth.mixed_geometries(11, 200, 256, 256)
    n = len(geoms)
    by = [str(i % 12) for i in range(n)]
    valid = (np.arange(n) % 5 != 0).astype(np.uint8)
    for fun in ("first", "last", "min", "max"):
        exp, got, _ = both(geoms, fun, "int32", 1 + (np.arange(n) * 2654435761 % 1000), 0, by=by, field_valid=valid,
                           shape=(256, 256), extent=(0, 0, 256, 256))
        assert exp.shape[0] == 12
        assert_same(exp, got)


def test_non_square_pixels_line_dedup():
    rng = np.random.default_rng(5)
    lines = []
    for _ in range(60):
        p = np.cumsum(rng.normal(0, 15, (rng.integers(2, 30), 2)), 0) + rng.random(2) * 200
        lines.append("LINESTRING (" + ", ".join(f"{a:.3f} {b:.3f}" for a, b in p) + ")")
    lines.append("MULTILINESTRING ((0 0, 200 200), (200 200, 0 0), (0 0, 200 200))")
    for shape in [(97, 311), (200, 200)]:
        for fun in ("sum", "count"):
            exp, got, _ = both(lines, fun, "int32", 1, 0, shape=shape, extent=(0, 0, 200, 200))
            assert_same(exp, got)


def test_geometry_outside_and_on_boundaries():
    geoms = [sq(-50, -50, 500, 500), sq(-10, 3, 2, 5), sq(98, 98, 120, 120), sq(0.5, 0.5, 99.5, 99.5),
             "POLYGON ((10 10, 90 10.0000000000000001, 90 50, 10 50, 10 10))",      # epsilon-horizontal edge
             "POLYGON ((20.5 20.5, 40.5 20.5, 40.5 40.5, 20.5 40.5, 20.5 20.5))",   # vertices on pixel centres
             "LINESTRING (-500 50, 500 50)", "LINESTRING (50 -1e6, 50 1e6)", "LINESTRING (-30 -30, 130 140)",
             "POINT (0 0)", "POINT (100 100)", "POINT (99.999 0.001)", "MULTIPOINT ((5 5), (5 5), (-1 5))",
             "POLYGON ((30 30, 70 70, 70 30, 30 70, 30 30))"]                        # self-intersecting bow-tie
    for fun in ("sum", "count", "first"):
        exp, got, _ = both(geoms, fun, "float32", np.arange(1, len(geoms) + 1, dtype=np.float32), np.nan,
                           shape=(100, 100), extent=(0, 0, 100, 100))
        assert_same(exp, got)


@pytest.mark.parametrize("tile_bytes", [64, 256, 4096, 32768])
def test_tile_width_does_not_change_results(tile_bytes):
    geoms = synth.mixed_geometries(21, 250, 1000, 120, rho=60.0)
    exp, got, st = both(geoms, "sum", "float32", np.arange(len(geoms), dtype=np.float32), np.nan, tile_bytes=tile_bytes,
                        shape=(120, 1000), extent=(0, 0, 1000, 120))
    assert st["tile_width"] == min(1024, tile_bytes // 4)
    assert_same(exp, got)


def test_row_band_shards_equal_full_raster():
    geoms = synth.mixed_geometries(31, 300, 400, 400, rho=40.0)
    burn = np.arange(len(geoms)) % 9
    for fun in ("sum", "last"):
        for r0, r1 in [(0, 100), (100, 101), (101, 399), (399, 400)]:
            exp, got, _ = both(geoms, fun, "int32", burn, 0, rows=(r0, r1), shape=(400, 400), extent=(0, 0, 400, 400))
            assert_same(exp, got)


def test_star_polygons_c1_shape_properties():
    # BASELINE config 1 at full size: parity against the oracle + size-independent properties
    x, y, off = synth.star_polygons(1, 10000, 64, 64, 82.0, 4096, 4096)
    vals = 100 * synth.splitmix_u(1, 10000, 9)
    g = core.Geoms.from_polygons(x, y, off)
    ri = core.raster_info(None, shape=(4096, 4096), extent=(0, 0, 4096, 4096))
    got, st = core.rasterize_dense(g, ri, "sum", "float64", vals, background=np.nan)
    og = oracle.Geoms.from_rings(x, y, off)
    ori = oracle.raster_info(None, shape=(4096, 4096), extent=(0, 0, 4096, 4096))
    exp, _ = oracle.rasterize_dense(og, ori, "sum", "float64", vals, background=np.nan)
    assert_same(exp, got)
    # linearity in the field for count (field-independent) and idempotence of `any`
    cnt, _ = core.rasterize_dense(g, ri, "count", "uint32", 1, background=0)
    anyv, _ = core.rasterize_dense(g, ri, "any", "uint8", 1, background=0)
    assert np.array_equal(anyv[0] == 1, cnt[0] > 0) and np.array_equal(np.isnan(got[0]), cnt[0] == 0)
    ones, _ = core.rasterize_dense(g, ri, "sum", "float64", 1.0, background=np.nan)
    assert np.array_equal(np.nan_to_num(ones[0]).astype(np.uint32), cnt[0])


def test_errors_through_the_abi():
    g = core.Geoms.from_wkt([sq(0, 0, 4, 4)])
    ri = core.raster_info(g, shape=(4, 4))
    with pytest.raises(ValueError, match="Geometry and field lengths must match"):
        core.rasterize_dense(g, ri, field=np.array([1.0, 2.0]))
    with pytest.raises(ValueError, match="Geometry and by lengths must match"):
        core.rasterize_dense(g, ri, band_of_geom=np.array([0, 1], np.int32), n_bands=2)


# ---- all_touched: GDAL-style line walk + 2-pass polygons, burners.rs:94-247 ----------
@pytest.mark.parametrize("fun", oracle.FUNS)
@pytest.mark.parametrize("dtype,bg", [("float64", np.nan), ("uint8", 0), ("int32", 2)])
def test_all_touched_mixed_geometries(fun, dtype, bg):
    seed = 500 + oracle.FUNS.index(fun)
    geoms = synth.mixed_geometries(seed, 120, 300, 200, rho=20.0)
    burn = (np.arange(len(geoms)) % 5).astype(dtype)
    exp, got, _ = both(geoms, fun, dtype, burn, bg, all_touched=True, shape=(200, 300), extent=(0, 0, 300, 200))
    assert_same(exp, got)


def test_all_touched_reference_cases():
    # test_many.py:237-262 geometry at res 0.5, and the custom-shape / extent variants (:283-325, :348-372)
    for kw in [dict(resolution=(0.5, 0.5)), dict(shape=(47, 319)), dict(resolution=(1, 1), extent=(-349, -507, 1, 0)),
               dict(shape=(47, 319), extent=(-349, -507, 1, 0))]:
        exp, got, _ = both(GEOMS_EXPLODED, "sum", "uint8", VALUES_EXPLODED, 0, all_touched=True, **kw)
        assert_same(exp, got)
    # R test-geometry.R:25-30: all_touched burns a superset of the standard cells
    kw = dict(shape=(4, 4), extent=(0, 0, 4, 4))
    _, on, _ = both(["LINESTRING (0 0, 4 4)"], burn=9.0, all_touched=True, **kw)
    _, off, _ = both(["LINESTRING (0 0, 4 4)"], burn=9.0, **kw)
    assert (on > 0).sum() > (off > 0).sum() and ((off > 0) <= (on > 0)).all()


def test_all_touched_shards_tiles_and_edges():
    from oracle.wkt2wkb import wkt_to_wkb

    geoms = synth.mixed_geometries(77, 200, 1500, 150, rho=70.0) + [wkt_to_wkb(s) for s in (
        "LINESTRING (-500 50, 2500 50)", "LINESTRING (750 -1e5, 750.001 1e5)", "LINESTRING (-30 -30, 1600 170)",
        sq(-50, -50, 5000, 5000))]
    burn = np.arange(len(geoms)) % 7
    for fun in ("sum", "last"):
        for rows in [None, (0, 60), (60, 150)]:
            exp, got, _ = both(geoms, fun, "int32", burn, 0, all_touched=True, rows=rows, tile_bytes=1024,
                               shape=(150, 1500), extent=(0, 0, 1500, 150))
            assert_same(exp, got)


# ---- the two engines (crossing records vs tile-binned) must agree with the oracle on polygon jobs ----
def _polygon_job(seed, n, width, height, rho):
    from oracle import wkt2wkb as W

    rng = np.random.default_rng(seed)
    out = []

    def star(cx, cy, r, nv):
        th = 2 * np.pi * (np.arange(nv) + 0.8 * rng.random(nv)) / nv
        rr = r * (0.5 + 0.5 * rng.random(nv))
        p = np.stack([cx + rr * np.cos(th), cy + rr * np.sin(th)], 1)
        if rng.random() < 0.3:
            p = np.round(p * 2) / 2  # vertices on pixel centres / edges
        return np.vstack([p, p[:1]])

    for _ in range(n):
        cx, cy = rng.random() * width * 1.2 - 0.1 * width, rng.random() * height * 1.2 - 0.1 * height
        r = rho * (0.2 + rng.random() ** 3 * 4)
        rings = [star(cx, cy, r, rng.integers(3, 60))]
        if rng.random() < 0.3:
            rings.append(star(cx, cy, r * 0.4, rng.integers(3, 12)))
        if rng.random() < 0.2:
            out.append(W.multipolygon_wkb([rings, [star(cx + r, cy, r * 0.7, 9)]]))
        else:
            out.append(W.polygon_wkb(rings))
    return out


@pytest.mark.parametrize("engine_flag", [8, 16], ids=["records", "tiles"])
@pytest.mark.parametrize("fun,dtype,bg", [("sum", "float32", np.nan), ("sum", "float64", 0.0), ("first", "int32", 0),
                                          ("last", "uint8", 3), ("min", "int16", 2), ("max", "float32", 2.0),
                                          ("count", "uint32", 0), ("any", "uint8", 0), ("sum", "int64", 7)])
def test_engines_polygon_jobs(engine_flag, fun, dtype, bg):
    seed = 900 + oracle.FUNS.index(fun)
    geoms = _polygon_job(seed, 400, 700, 500, 30.0)
    n = len(geoms)
    burn = (np.arange(n) % 6).astype(dtype)
    by = [str(i % 4) for i in range(n)] if seed % 2 else None
    exp, got, st = both(geoms, fun, dtype, burn, bg, by, flags=engine_flag, shape=(500, 700), extent=(0, 0, 700, 500))
    assert st["engine"] == (1 if engine_flag == 16 else 0)
    assert_same(exp, got)


@pytest.mark.parametrize("engine_flag", [8, 16], ids=["records", "tiles"])
def test_engines_shards_odd_rows_and_edges(engine_flag):
    geoms = _polygon_job(77, 300, 900, 400, 50.0)
    from oracle.wkt2wkb import wkt_to_wkb

    geoms += [wkt_to_wkb(s) for s in (
        sq(-50, -50, 5000, 5000), sq(127.5, 63.5, 128.5, 64.5), sq(0, 0, 128, 128), sq(128, 128, 256, 256),
        "POLYGON ((10 10, 890 10.0000000000000001, 890 50, 10 50, 10 10))",       # epsilon-horizontal edge
        "POLYGON ((0.5 0.5, 300 0.5000000000000001, 300 200.5, 0.5 0.5))",         # odd crossing count on row 0
        "POLYGON ((30 30, 770 370, 770 30, 30 370, 30 30))")]                      # bow-tie across many tiles
    burn = np.arange(len(geoms)) % 7 + 1
    for fun, dtype, bg in (("sum", "float64", np.nan), ("count", "int32", 0)):
        for rows in (None, (0, 130), (130, 400)):
            exp, got, st = both(geoms, fun, dtype, burn, bg, rows=rows, flags=engine_flag, shape=(400, 900),
                                extent=(0, 0, 900, 400))
            assert_same(exp, got)


def test_engine_choice_is_automatic():
    small = _polygon_job(5, 200, 2000, 2000, 20.0)
    _, _, st = both(small, shape=(2000, 2000), extent=(0, 0, 2000, 2000))
    assert st["engine"] == 1
    from oracle.wkt2wkb import wkt_to_wkb

    huge = [wkt_to_wkb("POLYGON ((" + ", ".join(f"{1000 + 990 * np.cos(a):.3f} {1000 + 990 * np.sin(a):.3f}" for a in
                                              np.linspace(0, 2 * np.pi, 5000)) + "))")] * 3
    exp, got, st = both(huge, "count", "uint8", 1, 0, shape=(2000, 2000), extent=(0, 0, 2000, 2000))
    assert st["engine"] == 0  # 5000-vertex polygon over 256 tiles: the record pipeline is cheaper
    assert_same(exp, got)


def test_windowed_host_output_and_streamed_upload(monkeypatch):
    """Host rasters come back in row windows through two staging buffers (copy of window k overlaps window k+1),
    and with RZ_FLAG_STREAMED_H2D the polygon pool is pulled from page-locked host memory window by window
    instead of uploaded up front (rz_tiles.cuh: part_bucket / pull_parts).  Small windows exercise both on a
    small job; every variant must equal the oracle and the plain path."""
    from rusterize_b200 import _lib

    x, y, off = synth.star_polygons(12, 6000, 16, 64, 60.0, 2048, 1536)
    vals = (100 * synth.splitmix_u(12, 6000, 9)).astype(np.float32)
    # parts far outside the raster and parts that only touch later windows stay in the pool
    x[:500] += 5000.0
    ri_kw = dict(shape=(1536, 2048), extent=(0, 0, 2048, 1536))
    og = oracle.Geoms.from_rings(x, y, off)
    exp, _ = oracle.rasterize_dense(og, oracle.raster_info(None, **ri_kw), "sum", "float32", vals, background=np.nan)
    ri = core.raster_info(None, **ri_kw)
    plain, st0 = core.rasterize_dense(core.Geoms.from_polygons(x, y, off), ri, "sum", "float32", vals, background=np.nan)
    assert st0["n_windows"] == 1 and np.array_equal(exp, plain, equal_nan=True)
    monkeypatch.setenv("RZ_WINDOW_BYTES", str(2048 * 4 * 100))  # 100-row windows -> 16 windows
    for flags in (0, _lib.FLAG_STREAMED_H2D, _lib.FLAG_STREAMED_H2D | _lib.FLAG_FORCE_TILE_ENGINE):
        g = core.Geoms.from_polygons(x, y, off)  # fresh handle: nothing cached on the device
        got, st = core.rasterize_dense(g, ri, "sum", "float32", vals, background=np.nan, flags=flags)
        assert st["n_windows"] == 16 and st["h2d_bytes"] >= g.n_coords * 16  # x, y (tags are rebuilt on the device)
        assert np.array_equal(exp, got, equal_nan=True), flags
        again, st2 = core.rasterize_dense(g, ri, "sum", "float32", vals, background=np.nan, flags=flags)  # cached copy
        assert st2["h2d_bytes"] < g.n_coords * 16 and np.array_equal(exp, again, equal_nan=True)
        shard, _ = core.rasterize_dense(g, ri, "last", "float32", vals, background=np.nan, rows=(250, 1111),
                                        flags=flags | _lib.FLAG_FORCE_H2D)
        e2, _ = oracle.rasterize_dense(og, oracle.raster_info(None, shape=(861, 2048), extent=(0, 1536 - 1111, 2048, 1536 - 250)),
                                       "last", "float32", vals, background=np.nan)
        assert np.array_equal(e2, shard, equal_nan=True)
    # mixed jobs take the record pipeline: the flag is ignored (and windows still work)
    geoms = synth.mixed_geometries(3, 300, 2048, 1536, rho=60.0)
    e3, g3, _ = both(geoms, "count", "uint16", 1, 0, flags=_lib.FLAG_STREAMED_H2D, **ri_kw)
    assert_same(e3, g3)


def test_all_touched_pixel_cache_box_aliasing():
    """all_touched with sum / count: the reference's FillWriter asks its PixelCache about fill pixels outside the
    cache's bounding box when ring segments lie entirely outside the raster (edges.rs:124-132 drops them from the
    box); the wrapped index (pixel_cache.rs:39-58) aliases onto other cells and the pixel is dropped if that cell
    was walked.  Both engines' dense path must reproduce it, shards and tiles included."""
    big = ["POLYGON ((-50 -50, 300 -50, 300 300, -50 300, -50 -50), (100 100, 150 100, 150 150, 100 150, 100 100))",
           "POLYGON ((-30 -40, 140 -10, 150 120, 60 90, -30 -40))", "POLYGON ((200 -100, 400 128, 200 400, 128 128, 200 -100))",
           "MULTIPOLYGON (((-10 -10, 270 -10, 270 270, -10 270, -10 -10), (20 20, 60 25, 40 70, 20 20)), "
           "((500 500, 600 500, 600 600, 500 500)))", "POLYGON ((10 10, 200 30, 120 220, 10 10))"]
    burn = [1, 2, 4, 8, 16]
    for fun in ("sum", "count"):
        for shape in [(256, 256), (64, 200), (200, 64)]:
            exp, got, _ = both(big, fun, "int32", burn, 0, all_touched=True, shape=shape, extent=(0, 0, 256, 256))
            assert_same(exp, got)
        for rows in [(0, 40), (40, 41), (41, 200)]:
            exp, got, _ = both(big, fun, "float64", burn, np.nan, all_touched=True, rows=rows, tile_bytes=256,
                               shape=(200, 300), extent=(0, 0, 256, 256))
            assert_same(exp, got)
    # the sparse stream of the same job replays to the same raster
    g = core.Geoms.from_wkt(big)
    ri = core.raster_info(g, shape=(200, 300), extent=(0, 0, 256, 256))
    sp = core.rasterize_sparse(g, ri, "sum", "int32", np.array(burn, np.int32), None, None, 1, 0, True)
    dense, _ = core.rasterize_dense(g, ri, "sum", "int32", np.array(burn, np.int32), None, None, 1, 0, True)
    rep = core.sparse_build_array(ri, "sum", 0, sp["counts"], sp["rows"], sp["cols"], sp["data"])
    assert np.array_equal(rep, dense)


def _same_bits(exp, got):
    """Bit-for-bit, signed zeros included; NaNs must coincide (the payload of a GENERATED NaN, inf - inf, is
    platform defined: x86 and the GPU differ there)."""
    assert exp.shape == got.shape and exp.dtype == got.dtype
    if exp.dtype.kind != "f":
        return bool(np.array_equal(exp, got))
    u = {4: np.uint32, 8: np.uint64}[exp.dtype.itemsize]
    nan_e, nan_g = np.isnan(exp), np.isnan(got)
    return bool(np.array_equal(nan_e, nan_g) and np.array_equal(exp.view(u)[~nan_e], got.view(u)[~nan_g]))


def test_tile_engine_apply_modes():
    """tile_apply evaluates `sum` / `count` additively when that is provably the reference's rule (float dtype,
    NaN background, all values finite; integer dtype, background 0) and falls back to the generic rule otherwise.
    Overlapping polygons, values chosen to hit every branch: signed zeros, sums that cancel to 0, NaN and
    infinite values (inf + -inf makes a NaN that the NEXT write must treat as untouched), integer wrap-around."""
    from rusterize_b200 import _lib

    x, y, off = synth.star_polygons(21, 1500, 8, 40, 70.0, 700, 500)
    n = len(off) - 1
    rng = np.random.default_rng(5)
    og = oracle.Geoms.from_rings(x, y, off)
    kw = dict(shape=(500, 700), extent=(0, 0, 700, 500))
    ori, ri = oracle.raster_info(None, **kw), core.raster_info(None, **kw)
    g = core.Geoms.from_polygons(x, y, off)
    finite = rng.choice(np.array([-0.0, 0.0, 1.5, -1.5, 3.25, 1e30, -1e30, 2.0**-140], np.float64), n)
    wild = finite.copy()
    wild[rng.random(n) < 0.08] = np.nan
    wild[rng.random(n) < 0.08] = np.inf
    wild[rng.random(n) < 0.08] = -np.inf
    cases = [("sum", "float32", finite, np.nan), ("sum", "float64", finite, np.nan), ("count", "float32", finite, np.nan),
             ("sum", "float32", wild, np.nan), ("sum", "float64", wild, np.nan), ("count", "float64", wild, np.nan),
             ("sum", "float32", finite, 0.0), ("sum", "float32", finite, 1.5), ("count", "float32", finite, 1.0),
             ("sum", "int32", rng.integers(-3, 4, n), 0), ("sum", "uint8", rng.integers(0, 256, n), 0),
             ("count", "uint8", 1, 0), ("sum", "int16", rng.integers(-3, 4, n), 2), ("count", "int64", 1, 3),
             ("sum", "int64", rng.integers(-2**62, 2**62, n), 0),
             # first / min / max: touched-mask mode when no value can equal the background, generic rule otherwise
             ("min", "int32", rng.integers(1, 50, n), 0), ("max", "int32", rng.integers(-50, 0, n), 0),
             ("first", "int16", rng.integers(1, 50, n), 0), ("min", "uint8", rng.integers(0, 5, n), 0),
             ("max", "int64", rng.integers(-3, 4, n), 2), ("first", "uint32", rng.integers(0, 3, n), 1),
             ("min", "float32", finite, np.nan), ("max", "float64", finite, np.nan), ("first", "float32", finite, np.nan),
             ("min", "float32", wild, np.nan), ("max", "float32", wild, np.nan), ("first", "float64", wild, np.nan),
             ("min", "float64", finite, 1.5), ("max", "float32", finite, 0.0), ("first", "float32", finite, -0.0),
             ("last", "float32", wild, np.nan), ("any", "uint8", 1, 0)]
    for fun, dtype, vals, bg in cases:
        v = np.asarray(vals).astype(dtype) if np.ndim(vals) else vals
        exp, _ = oracle.rasterize_dense(og, ori, fun, dtype, v, background=bg)
        got, st = core.rasterize_dense(g, ri, fun, dtype, v, background=bg, flags=_lib.FLAG_FORCE_TILE_ENGINE)
        assert st["engine"] == 1
        assert _same_bits(exp, got), (fun, dtype, bg)


def test_windowed_host_output_with_bands():
    """Row windows x `by` bands: every window copies one slab per band out of the staging buffer."""
    import os

    x, y, off = synth.star_polygons(13, 3000, 8, 32, 50.0, 1024, 900)
    n = len(off) - 1
    by = [str(i % 5) for i in range(n)]
    band, names = core.group_keys(by)
    vals = (np.arange(n) % 97 + 1).astype(np.int32)
    kw = dict(shape=(900, 1024), extent=(0, 0, 1024, 900))
    og = oracle.Geoms.from_rings(x, y, off)
    exp, _ = oracle.rasterize_dense(og, oracle.raster_info(None, **kw), "max", "int32", vals, None, by, 0, threads=4)
    g = core.Geoms.from_polygons(x, y, off)
    ri = core.raster_info(None, **kw)
    os.environ["RZ_WINDOW_BYTES"] = str(5 * 1024 * 4 * 64)  # windows of at most 64 rows x 5 bands
    try:
        for flags in (0, 8):  # tile engine / record pipeline
            got, st = core.rasterize_dense(g, ri, "max", "int32", vals, None, band, len(names), 0, flags=flags)
            assert st["n_windows"] >= 15, st["n_windows"]
            assert np.array_equal(exp, got), flags
    finally:
        del os.environ["RZ_WINDOW_BYTES"]
