"""Mixed jobs with an order-free pixel function (`any`; `count` / `sum` on an integer dtype with background 0, square
pixels, no all_touched): polygon parts go through the tile engine, line and point pixels are applied to the finished
raster by atomics (csrc/rz_burn.cuh).  Result must equal the oracle bit for bit - wrapping integer adds commute - and
the crossing-record pipeline (forced) must agree.  Jobs that are not eligible keep taking the record pipeline."""
import numpy as np
import pytest

import oracle
import synth
from oracle.wkt2wkb import wkt_to_wkb
from rusterize_b200 import _lib, core

pytestmark = pytest.mark.gpu
INT_DTYPES = ["uint8", "uint16", "uint32", "uint64", "int8", "int16", "int32", "int64"]


def _both(geoms, kw, fun, dtype, vals, bg, by=None, **extra):
    og = oracle.Geoms.from_wkb(geoms)
    exp, names = oracle.rasterize_dense(og, oracle.raster_info(og, **kw), fun, dtype, vals, None, by, bg)
    g = core.Geoms.from_wkb(geoms)
    ri = core.raster_info(g, **kw)
    band, n_b = (None, 1)
    if by is not None:
        band, bn = core.group_keys(by)
        n_b = len(bn)
        assert bn == names
    got, st = core.rasterize_dense(g, ri, fun, dtype, vals, None, band, n_b, bg, **extra)
    rec, st_r = core.rasterize_dense(g, ri, fun, dtype, vals, None, band, n_b, bg, flags=_lib.FLAG_NO_TILE_ENGINE, **extra)
    return exp, got, st, rec, st_r


@pytest.mark.parametrize("dtype", INT_DTYPES)
def test_count_and_sum_on_integer_dtypes(dtype):
    W, H = 517, 389
    geoms = synth.mixed_geometries(41, 600, W, H, rho=45.0)
    n = len(geoms)
    kw = dict(shape=(H, W), extent=(0, 0, W, H))
    rng = np.random.default_rng(41)
    info = np.iinfo(dtype)
    # values that wrap small dtypes (many overlaps): the adds must wrap exactly like the reference's release build
    vals = rng.integers(max(info.min, -100), min(info.max, 100) + 1, n).astype(dtype)
    for fun, v in (("count", 1), ("sum", vals)):
        exp, got, st, rec, st_r = _both(geoms, kw, fun, dtype, v, 0)
        assert st["engine"] == 1 and st_r["engine"] == 0
        assert np.array_equal(exp, got), (fun, dtype)
        assert np.array_equal(exp, rec), (fun, dtype)


@pytest.mark.parametrize("dtype,bg", [("uint8", 0), ("uint8", 7), ("int32", -1), ("float32", np.nan), ("float64", 0.0),
                                      ("int64", 1), ("uint16", 65535)])
def test_any_on_every_kind_of_dtype_and_background(dtype, bg):
    W, H = 300, 260
    geoms = synth.mixed_geometries(42, 400, W, H, rho=30.0)
    by = [str(i % 3) for i in range(len(geoms))]
    exp, got, st, rec, _ = _both(geoms, dict(shape=(H, W), extent=(0, 0, W, H)), "any", dtype, 1, bg, by=by)
    assert st["engine"] == 1
    assert np.array_equal(exp, got, equal_nan=True) and np.array_equal(exp, rec, equal_nan=True)


def test_lines_and_points_only_windows_and_shards(monkeypatch):
    """No polygon part at all (the tile engine only paints the background), small host windows, row shards."""
    W, H = 700, 900
    rng = np.random.default_rng(43)
    wkts = []
    for i in range(300):
        p = np.cumsum(rng.normal(0, 40, (rng.integers(2, 12), 2)), 0) + [rng.random() * W, rng.random() * H]
        wkts.append("LINESTRING (" + ", ".join(f"{a} {b}" for a, b in p) + ")")
    for i in range(200):
        p = rng.random((rng.integers(1, 5), 2)) * [W, H]
        wkts.append("MULTIPOINT (" + ", ".join(f"({a} {b})" for a, b in p) + ")")
    # a self-overlapping line (pixels written several times) and a very long one (warp-cooperative segment)
    wkts.append("LINESTRING (10 10, 200 10, 10 10, 200 10)")
    wkts.append("LINESTRING (-5000 -3000, 9000 4000)")
    geoms = [wkt_to_wkb(w) for w in wkts]
    kw = dict(shape=(H, W), extent=(0, 0, W, H))
    monkeypatch.setenv("RZ_WINDOW_BYTES", str(64 * W * 4))
    exp, got, st, rec, _ = _both(geoms, kw, "count", "uint32", 1, 0)
    assert st["engine"] == 1 and st["n_windows"] > 5
    assert np.array_equal(exp, got) and np.array_equal(exp, rec)
    assert got.max() >= 3  # the overlapping line really was written several times
    g = core.Geoms.from_wkb(geoms)
    ri = core.raster_info(g, **kw)
    for rows in [(0, 1), (123, 457), (899, 900)]:
        shard, st2 = core.rasterize_dense(g, ri, "count", "uint32", 1, background=0, rows=rows)
        assert st2["engine"] == 1 and np.array_equal(shard[0], exp[0, rows[0]:rows[1]])


def test_not_eligible_jobs_keep_the_record_pipeline():
    W, H = 200, 200
    geoms = synth.mixed_geometries(44, 200, W, H, rho=25.0)
    g = core.Geoms.from_wkb(geoms)
    ri = core.raster_info(g, shape=(H, W), extent=(0, 0, W, H))
    for fun, dtype, bg, kw in [("sum", "float32", np.nan, {}), ("count", "uint8", 3, {}), ("last", "int32", 0, {}),
                               ("count", "float64", 0.0, {}), ("count", "uint8", 0, dict(all_touched=True))]:
        _, st = core.rasterize_dense(g, ri, fun, dtype, 1, background=bg, **kw)
        assert st["engine"] == 0, (fun, dtype, bg)
    # non-square pixels: lines are de-duplicated per part (burn_geometry.rs:179)
    ri2 = core.raster_info(g, shape=(H, 2 * W), extent=(0, 0, W, H))
    _, st = core.rasterize_dense(g, ri2, "count", "uint8", 1, background=0)
    assert st["engine"] == 0
    # a segment beyond the supported domain is reported, not skipped silently
    far = core.Geoms.from_wkb([wkt_to_wkb("LINESTRING (0 0, 3e9 10)"), wkt_to_wkb("POLYGON ((1 1, 5 1, 5 5, 1 1))")])
    with pytest.raises(RuntimeError, match="2\\^29"):
        core.rasterize_dense(far, ri, "count", "uint32", 1, background=0)
