"""Pins the CPU oracle against every golden vector / known answer the reference's own tests hold
for the burn path.  CPU only."""
import numpy as np
import pytest
from PIL import Image

import oracle
from cases import GEOMS, GEOMS_EXPLODED, R_INFO, VALUES, VALUES_EXPLODED, rmat, sq

GOLDEN = __import__("os").path.join(__import__("os").path.dirname(__file__), "golden")


def test_g1_golden_tif_standard_sum():
    # python/test/test_many.py:228-235
    r = oracle.rusterize(GEOMS, res=(1, 1), burn=VALUES, fun="sum", dtype="uint8")
    g = np.array(Image.open(f"{GOLDEN}/standard_output_sum.tif"))
    assert r.shape == (1, 131, 361)
    assert np.array_equal(r[0], g)
    assert int(r.sum()) == 59204


def test_g2_golden_tif_custom_shape():
    # python/test/test_many.py:329-346 — non-square pixels => line dedup branch
    r = oracle.rusterize(GEOMS_EXPLODED, out_shape=(47, 319), burn=VALUES_EXPLODED, fun="sum", dtype="uint8")
    g = np.array(Image.open(f"{GOLDEN}/standard_output_sum_custom_shape.tif"))
    assert np.array_equal(r[0], g)
    assert int(r.sum()) == 19740


def test_g3_sparse_frame_from_docs():
    # python/docs/python.md:106-136
    sp = oracle.rusterize(GEOMS, res=(1, 1), burn=VALUES, fun="sum", dtype="float64", encoding="sparse")
    ri = sp["raster_info"]
    assert (ri.nrows, ri.ncols) == (131, 361)
    assert (ri.xmin, ri.ymin, ri.xmax, ri.ymax) == (-180.5, -70.5, 180.5, 60.5)
    assert (ri.xres, ri.yres) == (1.0, 1.0)
    assert len(sp["rows"]) == 29363
    head = list(zip(sp["rows"][:5].tolist(), sp["cols"][:5].tolist(), sp["data"][:5].tolist()))
    tail = list(zip(sp["rows"][-5:].tolist(), sp["cols"][-5:].tolist(), sp["data"][-5:].tolist()))
    assert head == [(6, 40, 1.0), (6, 41, 1.0), (6, 42, 1.0), (7, 39, 1.0), (7, 40, 1.0)]
    assert tail == [(39, 286, 5.0), (39, 287, 5.0), (39, 288, 5.0), (39, 289, 5.0), (39, 290, 5.0)]
    # G6: numpy == sparse.to_numpy (test_many.py:216-224)
    dense = oracle.rusterize(GEOMS, res=(1, 1), burn=VALUES, fun="sum", dtype="float64", background=np.nan)
    replay = oracle.sparse_replay(ri, sp, "sum", np.nan)
    assert np.array_equal(dense, replay, equal_nan=True)


# ---- G4: R known-answer matrices (R/rusterize/tests/testthat/test-boundary.R, test-geometry.R) ----
def _r(geoms, fun="last", bg=0.0, burn=1.0, **kw):
    args = dict(R_INFO)
    args.update(kw)
    return oracle.rusterize(geoms, fun=fun, background=bg, burn=burn, dtype="float64", **args)


def test_g4_full_and_partial_cover():
    assert np.array_equal(_r([sq(0, 0, 4, 4)], burn=7.0)[0], np.full((4, 4), 7.0))  # test-boundary.R:13-16
    exp = rmat([0, 0, 9, 9, 0, 0, 9, 9, 0, 0, 0, 0, 0, 0, 0, 0])  # :18-22
    assert np.array_equal(_r([sq(0, 0, 2, 2)], burn=9.0)[0], exp)


def test_g4_field_per_geometry():
    exp = rmat([10] * 8 + [20] * 8)  # test-boundary.R:24-28
    assert np.array_equal(_r([sq(0, 0, 2, 4), sq(2, 0, 4, 4)], burn=np.array([10.0, 20.0]))[0], exp)


@pytest.mark.parametrize("fun,vals,expect", [
    ("sum", [5, 7], 12), ("min", [5, 7], 5), ("max", [5, 7], 7), ("first", [5, 7], 5), ("last", [5, 7], 7),
    ("count", [1, 1], 2), ("any", [5, 7], 1)])
def test_g4_pixel_functions(fun, vals, expect):
    # test-boundary.R:30-42, test-geometry.R:3-7
    two = [sq(0, 0, 4, 4), sq(0, 0, 4, 4)]
    assert np.array_equal(_r(two, fun=fun, burn=np.array(vals, float))[0], np.full((4, 4), float(expect)))


def test_g4_dense_equals_sparse():
    d = _r([sq(0, 0, 4, 4)], burn=3.0)  # test-boundary.R:51-55
    sp = oracle.rusterize([sq(0, 0, 4, 4)], fun="last", background=0.0, burn=3.0, encoding="sparse", **R_INFO)
    assert np.array_equal(d, oracle.sparse_replay(sp["raster_info"], sp, "last", 0.0))


def test_g4_shape_only_vs_res_only():
    a = oracle.rusterize([sq(0, 0, 4, 4)], out_shape=(4, 4), background=0.0)  # test-boundary.R:57-66
    b = oracle.rusterize([sq(0, 0, 4, 4)], res=(1, 1), background=0.0)
    assert a.shape == (1, 4, 4) and b.shape == (1, 5, 5)


def test_g4_by_bands():
    geoms = [sq(0, 0, 2, 4), sq(2, 0, 4, 4)]  # test-boundary.R:84-93
    g = oracle.Geoms.from_any(geoms)
    ri = oracle.raster_info(g, shape=(4, 4), extent=(0, 0, 4, 4))
    arr, names = oracle.rasterize_dense(g, ri, "last", "float64", np.array([10.0, 20.0]), None, ["a", "b"], 0.0)
    assert names == ["a", "b"] and arr.shape == (2, 4, 4)
    assert np.array_equal(arr[0], rmat([10] * 8 + [0] * 8))
    assert np.array_equal(arr[1], rmat([0] * 8 + [20] * 8))


def test_g4_geometry_cases():
    m = _r(["POINT (1.5 1.5)"], burn=9.0)[0]  # test-geometry.R:9-15
    assert m.sum() == 9 and (m > 0).sum() == 1
    m = _r(["LINESTRING (0 0, 4 4)"], burn=9.0)[0]  # :17-23 exact diagonal cells
    assert np.array_equal(m, rmat([0, 0, 0, 0, 0, 0, 0, 9, 0, 0, 9, 0, 0, 9, 0, 0]))
    mp = "MULTIPOLYGON (((0 0, 2 0, 2 2, 0 2, 0 0)), ((2 2, 4 2, 4 4, 2 4, 2 2)))"  # :32-39
    assert np.array_equal(_r([mp], burn=9.0)[0], rmat([0, 0, 9, 9, 0, 0, 9, 9, 9, 9, 0, 0, 9, 9, 0, 0]))
    m = _r([sq(0, 0, 2, 2)], bg=-1.0, burn=9.0)[0]  # :41-47
    assert (m == 9).sum() == 4 and (m[m != 9] == -1).all()
    m = _r([sq(0, 0, 4, 4)], bg=np.nan, burn=3.0)[0]  # :49-52
    assert np.array_equal(m, np.full((4, 4), 3.0))
    on = _r(["LINESTRING (0 0, 4 4)"], burn=9.0, all_touched=True)[0]  # :25-30
    off = _r(["LINESTRING (0 0, 4 4)"], burn=9.0)[0]
    assert (on > 0).sum() > (off > 0).sum()


# ---- G5: Rust unit known answers ----
def _writes(geoms, **kw):
    sp = oracle.rusterize(geoms, out_shape=(10, 10), extent=(0, 0, 10, 10), encoding="sparse", burn=1.0, **kw)
    return list(zip(sp["rows"].tolist(), sp["cols"].tolist()))


def test_g5_point_multipoint_line():
    # rust/src/rasterization/burn_geometry.rs:262-304 (raster 10x10 over (0,0)-(10,10))
    assert _writes(["POINT (2.5 7.5)"]) == [(2, 2)]
    assert _writes(["MULTIPOINT ((1.5 8.5), (5.5 3.5))"]) == [(1, 1), (6, 5)]
    w = _writes(["LINESTRING (1.5 4.5, 8.5 4.5)"])
    assert w and all(r == 5 for r, _ in w)


def test_g5_group_keys():
    # rust/src/rasterize.rs:307-313
    band, names = oracle.group_keys(["b", "a", "b"])
    assert names == ["a", "b"] and band.tolist() == [1, 0, 1]
    # lexicographic, not numeric
    band, names = oracle.group_keys(["10", "9", "2"])
    assert names == ["10", "2", "9"]


def test_length_errors():
    # rust/src/rasterize.rs:208-229
    g = oracle.Geoms.from_any([sq(0, 0, 4, 4)])
    ri = oracle.raster_info(g, shape=(4, 4))
    with pytest.raises(oracle.OracleError, match="Geometry and field lengths must match"):
        oracle.rasterize_dense(g, ri, burn=np.array([1.0, 2.0]))
    with pytest.raises(oracle.OracleError, match="Geometry and by lengths must match"):
        oracle.rasterize_dense(g, ri, by=["a", "b"])


def test_raster_info_errors():
    # rust/src/geo/raster.rs:50-121
    g = oracle.Geoms.from_any([sq(0, 0, 4, 4)])
    for kw, msg in [
        (dict(), "Must set at least one"),
        (dict(shape=(4, 4), resolution=(1, 1)), "mutually exclusive"),
        (dict(shape=(0, 4)), "Shape values must be > 0"),
        (dict(resolution=(0.0, 1.0)), "Resolution values must be > 0"),
        (dict(shape=(4, 4), extent=(0, 0, 0, 0)), "Unspecified extent"),
    ]:
        with pytest.raises(oracle.OracleError, match=msg):
            oracle.raster_info(g, **kw)
    ri = oracle.raster_info(g, resolution=(0.3, 0.3), tap=True)
    assert ri.xmin == 0.0 and ri.xmax == np.ceil(4 / 0.3) * 0.3
