"""GPU parity of the sparse (COO) encoding: triplets must equal the oracle's, element for element
and in the same (burn) order — rust/src/encoding/writers.rs:86-131."""
import numpy as np
import pytest

import oracle
import synth
from cases import GEOMS, VALUES, sq
from rusterize_b200 import core

pytestmark = pytest.mark.gpu


def both(geoms, fun="last", dtype="float64", burn=1, bg=0, by=None, field_valid=None, all_touched=False, **kw):
    og = oracle.Geoms.from_any(geoms)
    ori = oracle.raster_info(og, **kw)
    exp = oracle.rasterize_sparse(og, ori, fun, dtype, burn, field_valid, by, bg, all_touched)
    g = core.Geoms.from_any(geoms)
    ri = core.raster_info(g, **kw)
    band, nb = None, 1
    if by is not None:
        band, bn = core.group_keys(by)
        nb = len(bn)
    got = core.rasterize_sparse(g, ri, fun, dtype, burn, field_valid, band, nb, bg, all_touched)
    return exp, got, ri


def assert_same(exp, got):
    assert np.array_equal(exp["counts"], got["counts"]), (exp["counts"], got["counts"])
    for k in ("rows", "cols"):
        assert np.array_equal(exp[k], got[k]), k
    assert exp["data"].dtype == got["data"].dtype
    assert np.array_equal(exp["data"], got["data"], equal_nan=exp["data"].dtype.kind == "f")


def test_documented_sparse_frame():
    # python/docs/python.md:106-136
    exp, got, ri = both(GEOMS, "sum", "float64", VALUES, np.nan, resolution=(1, 1))
    assert_same(exp, got)
    assert len(got["rows"]) == 29363
    assert list(zip(got["rows"][:3].tolist(), got["cols"][:3].tolist())) == [(6, 40), (6, 41), (6, 42)]
    assert list(zip(got["rows"][-2:].tolist(), got["cols"][-2:].tolist(), got["data"][-2:].tolist())) == \
        [(39, 289, 5.0), (39, 290, 5.0)]


@pytest.mark.parametrize("dtype", ["uint8", "int32", "float32", "float64", "uint64", "int16"])
def test_mixed_geometries(dtype):
    seed = 100 + oracle.DTYPES.index(dtype)
    geoms = synth.mixed_geometries(seed, 200, 300, 300, rho=25.0)
    burn = (np.arange(len(geoms)) % 11).astype(dtype)
    exp, got, _ = both(geoms, "sum", dtype, burn, 0, shape=(300, 300), extent=(0, 0, 300, 300))
    assert_same(exp, got)


def test_bands_and_null_fields():
    geoms = synth.mixed_geometries(7, 250, 256, 256)
    n = len(geoms)
    by = [str(i % 7) for i in range(n)]
    valid = (np.arange(n) % 4 != 1).astype(np.uint8)
    exp, got, _ = both(geoms, "max", "int32", np.arange(n), 0, by=by, field_valid=valid, shape=(256, 256),
                       extent=(0, 0, 256, 256))
    assert len(got["counts"]) == 7
    assert_same(exp, got)


def test_edge_cases():
    geoms = [sq(-50, -50, 500, 500), sq(98, 98, 120, 120), "POLYGON ((30 30, 70 70, 70 30, 30 70, 30 30))",
             "LINESTRING (-500 50, 500 50)", "LINESTRING (0 0, 100 100, 0 0)", "POINT (100 100)", "POINT (5 5)",
             "MULTIPOINT ((5 5), (5 5), (-1 5))", "GEOMETRYCOLLECTION EMPTY", "POLYGON EMPTY"]
    exp, got, _ = both(geoms, "count", "uint16", np.arange(len(geoms)), 0, shape=(100, 100), extent=(0, 0, 100, 100))
    assert_same(exp, got)
    # nothing burned at all
    exp, got, _ = both(["POINT (1000 1000)"], shape=(10, 10), extent=(0, 0, 10, 10))
    assert len(got["rows"]) == 0 and got["counts"].tolist() == [0]


def test_small_parcels_like_config5():
    # BASELINE config 5 in miniature: axis-jittered quads, sparse, f32
    rng = np.random.default_rng(5)
    n, size = 20000, 2048
    cx, cy = rng.random(n) * size, rng.random(n) * size
    w, h = 6 + 8 * rng.random(n), 6 + 8 * rng.random(n)
    j = rng.random((n, 4, 2)) - 0.5
    quad = np.stack([np.stack([cx - w / 2, cy - h / 2], 1), np.stack([cx + w / 2, cy - h / 2], 1),
                     np.stack([cx + w / 2, cy + h / 2], 1), np.stack([cx - w / 2, cy + h / 2], 1)], 1) + j
    ring = np.concatenate([quad, quad[:, :1]], 1).reshape(-1, 2)
    off = np.arange(n + 1, dtype=np.uint64) * 5
    vals = rng.random(n).astype(np.float32)
    g = core.Geoms.from_polygons(ring[:, 0], ring[:, 1], off)
    ri = core.raster_info(None, shape=(size, size), extent=(0, 0, size, size))
    got = core.rasterize_sparse(g, ri, "sum", "float32", vals, background=np.nan)
    og = oracle.Geoms.from_rings(ring[:, 0], ring[:, 1], off)
    ori = oracle.raster_info(None, shape=(size, size), extent=(0, 0, size, size))
    exp = oracle.rasterize_sparse(og, ori, "sum", "float32", vals, None, None, np.nan)
    assert_same(exp, got)
    # size-independent property: replaying the triplets gives the dense raster
    dense, _ = core.rasterize_dense(g, ri, "sum", "float32", vals, background=np.nan)
    assert np.array_equal(oracle.sparse_replay(ori, got | {"data": got["data"]}, "sum", np.nan), dense, equal_nan=True)


def test_non_square_pixels_line_dedup_first_visits_in_burn_order():
    # PixelCache (pixel_cache.rs, writers.rs:15-36): on non-square pixels a line part writes a pixel only
    # the first time it visits it
    rng = np.random.default_rng(9)
    lines = []
    for _ in range(80):
        p = np.cumsum(rng.normal(0, 12, (rng.integers(2, 40), 2)), 0) + rng.random(2) * 200
        lines.append("LINESTRING (" + ", ".join(f"{a:.3f} {b:.3f}" for a, b in p) + ")")
    lines += ["MULTILINESTRING ((0 0, 200 200), (200 200, 0 0), (0 0, 200 200))", "LINESTRING (-500 50, 500 50, -500 50)",
              "POLYGON ((20 20, 120 20, 120 90, 20 20))", "POINT (5 5)", "LINESTRING (10 10, 10 10)"]
    for shape in [(97, 311), (311, 97)]:
        exp, got, _ = both(lines, "sum", "int32", np.arange(len(lines)), 0, shape=shape, extent=(0, 0, 200, 200))
        assert_same(exp, got)
    exp, got, _ = both(lines, "count", "float32", 1, np.nan, by=[str(i % 3) for i in range(len(lines))],
                       shape=(60, 250), extent=(0, 0, 200, 200))
    assert_same(exp, got)


# ---- all_touched (burners.rs:94-247; burn_geometry.rs:89-106, 225-238): a polygon part writes its rings'
# boundary walk, then its fill; with sum / count every part writes a pixel once (PixelCache) -----------------
@pytest.mark.parametrize("fun,dtype,bg", [("sum", "int32", 0), ("count", "float32", np.nan), ("last", "uint8", 0),
                                          ("max", "float64", np.nan), ("first", "int64", 0)])
def test_all_touched_mixed_geometries(fun, dtype, bg):
    geoms = synth.mixed_geometries(31, 220, 300, 200, rho=22.0)
    burn = (np.arange(len(geoms)) % 13 + 1).astype(dtype)
    exp, got, _ = both(geoms, fun, dtype, burn, bg, all_touched=True, shape=(200, 300), extent=(0, 0, 300, 200))
    assert len(exp["rows"]) > 0
    assert_same(exp, got)


def test_all_touched_reference_geometries_bands_and_non_square_pixels():
    for fun in ("sum", "last"):
        exp, got, _ = both(GEOMS, fun, "float64", VALUES, np.nan, all_touched=True, resolution=(1, 1))
        assert_same(exp, got)
        exp, got, _ = both(GEOMS, fun, "int32", VALUES, 0, by=["b", "a", "b", "a", "c"], all_touched=True,
                           shape=(47, 319))
        assert_same(exp, got)
    # A ring segment lying entirely outside the raster is dropped by extract_line (edges.rs:124-132).  With
    # sum / count the reference then asks its PixelCache about fill pixels outside the cache's bounding box,
    # through a wrapped index (pixel_cache.rs:39-58) that aliases onto other cells: reproduced (cache_contains).
    big = ["POLYGON ((-50 -50, 300 -50, 300 300, -50 300, -50 -50), (100 100, 150 100, 150 150, 100 150, 100 100))",
           "MULTILINESTRING ((0 0, 256 256), (256 256, 0 0))", "POLYGON ((10 10, 200 30, 120 220, 10 10))",
           "POLYGON ((-30 -40, 140 -10, 150 120, 60 90, -30 -40))", "POLYGON ((200 -100, 400 128, 200 400, 128 128, 200 -100))"]
    for fun, shape in [("count", (64, 200)), ("sum", (200, 64)), ("sum", (256, 256)), ("last", (200, 64)), ("min", (64, 200))]:
        exp, got, _ = both(big, fun, "float32", [1.5, 2.5, 4.0, 8.0, 16.0], np.nan, all_touched=True, shape=shape,
                           extent=(0, 0, 256, 256))
        assert_same(exp, got)
    geoms = synth.mixed_geometries(5, 150, 256, 256)
    n = len(geoms)
    valid = (np.arange(n) % 5 != 2).astype(np.uint8)
    for fun, shape in [("count", (64, 200)), ("min", (200, 64)), ("sum", (256, 256))]:
        exp, got, _ = both(geoms, fun, "int16", np.arange(n) % 9, 0, by=[str(i % 4) for i in range(n)],
                           field_valid=valid, all_touched=True, shape=shape, extent=(0, 0, 256, 256))
        assert_same(exp, got)


def test_all_touched_sparse_replays_to_the_dense_raster():
    # SparseArray.to_numpy == numpy encoding (python/test/test_many.py:216-224), with all_touched
    geoms = synth.mixed_geometries(77, 120, 128, 128)
    g = core.Geoms.from_any(geoms)
    ri = core.raster_info(g, shape=(128, 128), extent=(0, 0, 128, 128))
    burn = (np.arange(len(geoms)) % 5 + 1).astype(np.float32)
    for fun in ("sum", "count", "min"):
        sp = core.rasterize_sparse(g, ri, fun, "float32", burn, None, None, 1, np.nan, True)
        dense, _ = core.rasterize_dense(g, ri, fun, "float32", burn, None, None, 1, np.nan, True)
        rep = core.sparse_build_array(ri, fun, np.nan, sp["counts"], sp["rows"], sp["cols"], sp["data"])
        rep = rep[0] if isinstance(rep, tuple) else rep
        assert np.array_equal(rep, dense, equal_nan=True), fun
