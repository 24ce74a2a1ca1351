"""Multi-device calls of the C ABI (rz_rasterize_dense_multi / rz_rasterize_sparse_multi): one call, row bands
(dense) or contiguous geometry ranges (sparse) per device, results landing in ONE caller-owned array / stream.
Everything is compared with the CPU oracle, bit for bit (rust/src/rasterize.rs:77-157 is what one call
returns).  On a single-GPU box the dense sharding is exercised with a repeated device (RZ_ALLOW_REPEATED_DEVICES);
the tests that need two real devices skip."""
import numpy as np
import pytest

import oracle
import synth
from rusterize_b200 import _lib, core

pytestmark = pytest.mark.gpu


def _n_dev():
    return _lib.lib().rz_device_count()


@pytest.fixture()
def repeated(monkeypatch):
    monkeypatch.setenv("RZ_ALLOW_REPEATED_DEVICES", "1")


def test_dense_multi_row_bands_mixed_geometries_all_functions(repeated):
    """Mixed polygons / lines / points / collections with `by` bands: 3 row bands (odd row count, so the bands differ
    in height) reproduce the oracle for every pixel function, with and without all_touched."""
    W, H = 517, 389
    geoms = synth.mixed_geometries(31, 400, W, H, rho=40.0)
    n = len(geoms)
    kw = dict(shape=(H, W), extent=(0, 0, W, H))
    by = [str(i % 3) for i in range(n)]
    band, names = core.group_keys(by)
    g = core.Geoms.from_wkb(geoms)
    og = oracle.Geoms.from_wkb(geoms)
    ri, ori = core.raster_info(g, **kw), oracle.raster_info(og, **kw)
    rng = np.random.default_rng(31)
    for fun, dtype, bg in [("sum", "float32", np.nan), ("count", "uint32", 0), ("first", "int32", 0), ("last", "float64", -1.0),
                           ("min", "int16", 0), ("max", "uint8", 0), ("any", "uint8", 0)]:
        vals = rng.integers(1, 50, n).astype(dtype)
        for touched in (False, True):
            exp, onames = oracle.rasterize_dense(og, ori, fun, dtype, vals, None, by, bg, all_touched=touched)
            got, st = core.rasterize_dense(g, ri, fun, dtype, vals, None, band, len(names), bg, all_touched=touched,
                                           devices=[0, 0, 0])
            assert onames == names
            assert np.array_equal(exp, got, equal_nan=True), (fun, dtype, touched)
            assert len(st["per_device"]) == 3 and st["out_bytes"] == exp.nbytes
            # each shard was given fewer parts than the whole set
            assert all(p["n_parts"] <= g.n_parts for p in st["per_device"])


def test_dense_multi_tile_engine_polygons_and_row_window(repeated):
    """Polygon-only job (tile engine on every shard), 4 shards of a row window of the grid; the shards' part subsets
    are cached in the handle and reused by the second call."""
    W, H = 1500, 1100
    x, y, off = synth.star_polygons(32, 4000, 8, 40, 40.0, W, H)
    vals = (10 * synth.splitmix_u(32, 4000, 9)).astype(np.float32)
    kw = dict(shape=(H, W), extent=(0, 0, W, H))
    g = core.Geoms.from_polygons(x, y, off)
    ri = core.raster_info(None, **kw)
    og = oracle.Geoms.from_rings(x, y, off)
    exp = oracle.rasterize_dense(og, oracle.raster_info(None, **kw), "sum", "float32", vals, None, None, np.nan)[0]
    for rows in (None, (130, 977)):
        for it in range(2):
            got, st = core.rasterize_dense(g, ri, "sum", "float32", vals, background=np.nan, rows=rows, devices=[0] * 4)
            want = exp if rows is None else exp[:, rows[0]:rows[1]]
            assert np.array_equal(want, got, equal_nan=True), (rows, it)
            assert st["engine"] == 1
            assert sum(p["n_parts"] for p in st["per_device"]) < 2 * g.n_parts  # shards overlap only at band borders


def test_dense_multi_more_devices_than_rows_and_empty_shards(repeated):
    x, y, off = synth.star_polygons(33, 50, 5, 9, 3.0, 40, 3)
    kw = dict(shape=(3, 40), extent=(0, 0, 40, 3))
    g = core.Geoms.from_polygons(x, y, off)
    og = oracle.Geoms.from_rings(x, y, off)
    exp = oracle.rasterize_dense(og, oracle.raster_info(None, **kw), "count", "int32", 1, None, None, 0)[0]
    got, st = core.rasterize_dense(g, core.raster_info(None, **kw), "count", "int32", 1, background=0, devices=[0] * 5)
    assert np.array_equal(exp, got)
    # a shard with no part at all still writes its background rows
    x2, y2, off2 = synth.star_polygons(34, 20, 5, 9, 4.0, 64, 16)
    y2 = y2 + 240.0  # all parts in the top rows of a 256-row grid
    kw2 = dict(shape=(256, 64), extent=(0, 0, 64, 256))
    g2 = core.Geoms.from_polygons(x2, y2, off2)
    exp2 = oracle.rasterize_dense(oracle.Geoms.from_rings(x2, y2, off2), oracle.raster_info(None, **kw2), "sum", "float64", 2.5,
                                  None, None, np.nan)[0]
    got2, st2 = core.rasterize_dense(g2, core.raster_info(None, **kw2), "sum", "float64", 2.5, background=np.nan,
                                     devices=[0, 0, 0, 0])
    assert np.array_equal(exp2, got2, equal_nan=True)
    assert st2["per_device"][3]["n_parts"] == 0


def test_multi_rejects_bad_device_lists():
    x, y, off = synth.star_polygons(35, 10, 5, 9, 3.0, 40, 30)
    g = core.Geoms.from_polygons(x, y, off)
    ri = core.raster_info(None, shape=(30, 40), extent=(0, 0, 40, 30))
    with pytest.raises(ValueError, match="twice"):
        core.rasterize_dense(g, ri, "sum", "float32", 1, background=0, devices=[0, 0])
    with pytest.raises(ValueError, match="twice"):
        core.rasterize_sparse(g, ri, "sum", "float32", 1, background=0, devices=[0, 0])
    with pytest.raises(RuntimeError, match="Invalid CUDA device"):
        core.rasterize_dense(g, ri, "sum", "float32", 1, background=0, devices=[_n_dev() + 3])


def test_sparse_multi_single_device_equals_plain_call():
    geoms = synth.mixed_geometries(36, 300, 300, 200, rho=25.0)
    n = len(geoms)
    g = core.Geoms.from_wkb(geoms)
    ri = core.raster_info(None, shape=(200, 300), extent=(0, 0, 300, 200))
    vals = np.arange(n, dtype=np.float32)
    by = [str(i % 2) for i in range(n)]
    band, names = core.group_keys(by)
    a = core.rasterize_sparse(g, ri, "sum", "float32", vals, None, band, 2, np.nan)
    b = core.rasterize_sparse(g, ri, "sum", "float32", vals, None, band, 2, np.nan, devices=[0])
    for k in ("rows", "cols", "data", "counts"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.skipif("_n_dev() < 2")
def test_two_devices_dense_and_sparse_equal_the_oracle():
    """Two real devices: dense row bands and sparse geometry ranges, `by` bands included."""
    nd = min(_n_dev(), 4)
    devs = list(range(nd))
    W, H = 900, 700
    geoms = synth.mixed_geometries(37, 1500, W, H, rho=30.0)
    n = len(geoms)
    kw = dict(shape=(H, W), extent=(0, 0, W, H))
    by = [str(i % 5) for i in range(n)]
    band, names = core.group_keys(by)
    g = core.Geoms.from_wkb(geoms)
    og = oracle.Geoms.from_wkb(geoms)
    ri, ori = core.raster_info(g, **kw), oracle.raster_info(og, **kw)
    vals = (np.arange(n) % 17 + 1).astype(np.float32)
    for touched in (False, True):
        exp, _ = oracle.rasterize_dense(og, ori, "sum", "float32", vals, None, by, np.nan, all_touched=touched)
        got, st = core.rasterize_dense(g, ri, "sum", "float32", vals, None, band, len(names), np.nan, all_touched=touched,
                                       devices=devs)
        assert np.array_equal(exp, got, equal_nan=True), touched
        osp = oracle.rasterize_sparse(og, ori, "sum", "float32", vals, None, by, np.nan, all_touched=touched)
        sp = core.rasterize_sparse(g, ri, "sum", "float32", vals, None, band, len(names), np.nan, all_touched=touched,
                                   devices=devs)
        for k in ("counts", "rows", "cols", "data"):
            assert np.array_equal(np.asarray(osp[k]), np.asarray(sp[k])), (k, touched)
        assert len(sp["stats"]["per_device"]) == nd
    # config-5 shaped: parcels, one band, geometry ranges balanced by estimated work
    px, py, poff = synth.parcels(5, 60000, 4000, 4000)
    pv = synth.splitmix_u(5, 60000, 9).astype(np.float32)
    gp = core.Geoms.from_polygons(px, py, poff)
    rip = core.raster_info(None, shape=(4000, 4000), extent=(0, 0, 4000, 4000))
    one = core.rasterize_sparse(gp, rip, "sum", "float32", pv, background=np.nan)
    many = core.rasterize_sparse(gp, rip, "sum", "float32", pv, background=np.nan, devices=devs)
    for k in ("counts", "rows", "cols", "data"):
        assert np.array_equal(one[k], many[k]), k
    trip = [p["out_bytes"] for p in many["stats"]["per_device"]]
    assert max(trip) < 1.25 * (sum(trip) / nd)  # balanced split


@pytest.mark.parametrize("runs", [None, "3"])
def test_one_shot_dense_soa_equals_the_oracle(repeated, monkeypatch, runs):
    """rz_rasterize_dense_soa: flatten + upload + burn + copy-back in ONE call; with several devices each flattens only
    the parts of its row band straight out of the caller's arrays, and a device's rows go through in runs (the next
    run is flattened while the previous one burns and copies back; forced to 3 runs here, large rasters choose it
    themselves).  Mixed geometries with `by` bands, a row window, all_touched, one device and four shards."""
    if runs:
        monkeypatch.setenv("RZ_ONE_SHOT_RUNS", runs)
    W, H = 517, 389
    geoms = synth.mixed_geometries(38, 500, W, H, rho=40.0)
    n = len(geoms)
    soa = synth.wkb_to_soa(geoms)
    kw = dict(shape=(H, W), extent=(0, 0, W, H))
    by = [str(i % 3) for i in range(n)]
    band, names = core.group_keys(by)
    og = oracle.Geoms.from_wkb(geoms)
    ri, ori = core.raster_info(None, **kw), oracle.raster_info(None, **kw)
    vals = (np.arange(n) % 13 + 1).astype(np.float32)
    for touched in (False, True):
        exp, _ = oracle.rasterize_dense(og, ori, "sum", "float32", vals, None, by, np.nan, all_touched=touched)
        for devs in ([0], [0, 0, 0, 0]):
            got, st = core.rasterize_dense_soa(soa, ri, "sum", "float32", vals, None, band, len(names), np.nan,
                                               all_touched=touched, devices=devs)
            assert np.array_equal(exp, got, equal_nan=True), (touched, devs)
            assert len(st["per_device"]) == len(devs)
        win, _ = core.rasterize_dense_soa(soa, ri, "sum", "float32", vals, None, band, len(names), np.nan,
                                          all_touched=touched, devices=[0, 0, 0], rows=(100, 333))
        assert np.array_equal(exp[:, 100:333], win, equal_nan=True)
    if _n_dev() >= 2:
        devs = list(range(min(_n_dev(), 4)))
        exp, _ = oracle.rasterize_dense(og, ori, "count", "uint16", 1, None, by, 0)
        got, st = core.rasterize_dense_soa(soa, ri, "count", "uint16", 1, None, band, len(names), 0, devices=devs)
        assert np.array_equal(exp, got)
    with pytest.raises(ValueError, match="Geometry and field lengths must match"):
        core.rasterize_dense_soa(soa, ri, "sum", "float32", vals[:-1], None, band, len(names), np.nan, devices=[0])
