"""The reference-facing Python surface (`rusterize()` -> `_rusterize()` -> C ABI) on the GPU: cases
ported from /root/reference/python/test/test_many.py to list-of-WKT / list-of-WKB inputs."""
import numpy as np
import pytest
from PIL import Image

import oracle
import synth
from cases import GEOMS, VALUES
from oracle.wkt2wkb import wkt_to_wkb
from rusterize_b200 import SparseArray, _lib, core, rusterize

pytestmark = pytest.mark.gpu
GOLDEN = __import__("os").path.join(__import__("os").path.dirname(__file__), "golden")


def test_coherence_with_golden_tif():
    # TestCoherence.test_standard (test_many.py:228-235)
    r = rusterize(GEOMS, res=(1, 1), dtype="uint8", burn=VALUES, fun="sum", encoding="numpy").squeeze()
    assert np.array_equal(r, np.array(Image.open(f"{GOLDEN}/standard_output_sum.tif")))


def test_input_formats_agree():
    # TestFormats.test_inputs (test_many.py:167-201), the inputs that exist without geopandas/polars
    a = rusterize(GEOMS, res=(1, 1), dtype="uint8", fun="sum", encoding="numpy")
    b = rusterize(np.asarray(GEOMS), res=(1, 1), dtype="uint8", fun="sum", encoding="numpy")
    wkb = [wkt_to_wkb(s) for s in GEOMS]
    c = rusterize(wkb, res=(1, 1), dtype="uint8", fun="sum", encoding="numpy")
    d = rusterize(np.asarray(wkb, dtype=object), res=(1, 1), dtype="uint8", fun="sum", encoding="numpy")
    assert a.shape == (1, 131, 361)
    assert np.array_equal(a, b) and np.array_equal(a, c) and np.array_equal(a, d)


def test_arguments():
    # TestArguments (test_many.py:115-163)
    r = rusterize(GEOMS, res=(1, 1), burn=99, encoding="numpy").squeeze()
    assert np.nanmax(r) == 99 and np.nanmin(r[r > 0]) == 99 and np.isnan(r[0, 0])
    r = rusterize(GEOMS, res=(1, 1), burn=1, background=-1, encoding="numpy").squeeze()
    assert r[0, 0] == -1
    # background that does not fit the dtype silently becomes 0 (python/src/rusterize.rs:50-53)
    r = rusterize(GEOMS, res=(1, 1), burn=1, dtype="uint8", encoding="numpy").squeeze()
    assert r[0, 0] == 0
    r = rusterize(GEOMS, res=(1, 1), burn=1, dtype="uint8", background=-1, encoding="numpy").squeeze()
    assert r[0, 0] == 0
    with pytest.raises(TypeError):
        rusterize(GEOMS, res=(1, 1), burn=1.5, dtype="uint8", encoding="numpy")


def test_outputs_numpy_equals_sparse():
    # TestFormats.test_outputs (test_many.py:216-224): numpy == sparse.to_numpy
    r_numpy = rusterize(GEOMS, res=(1, 1), dtype="uint8", burn=VALUES, encoding="numpy")
    sp = rusterize(GEOMS, res=(1, 1), dtype="uint8", burn=VALUES, encoding="sparse")
    assert isinstance(sp, SparseArray)
    assert np.array_equal(r_numpy, sp.to_numpy())
    assert sp.shape() == (1, 131, 361) and sp.extent() == (-180.5, -70.5, 180.5, 60.5) and sp.resolution() == (1.0, 1.0)
    assert repr(sp) == ("SparseArray:\n- Shape: (1, 131, 361)\n- Extent: (-180.5, -70.5, 180.5, 60.5)\n"
                        "- Resolution: (1.0, 1.0)\n- EPSG: None\n- Estimated size: 47.29 KB")
    sp64 = rusterize(GEOMS, res=(1, 1), burn=VALUES.astype(float), fun="sum", encoding="sparse")
    assert "378.33 KB" in repr(sp64) and len(sp64.rows) == 29363  # python/docs/python.md:106-136


@pytest.mark.parametrize("fun", oracle.FUNS)
def test_sparse_replay_matches_dense(fun):
    geoms = synth.mixed_geometries(41, 300, 400, 300, rho=40.0)
    n = len(geoms)
    burn = (np.arange(n) % 5).astype(np.float32)
    burn[::17] = np.nan
    g = core.Geoms.from_wkb(geoms)
    ri = core.raster_info(g, shape=(300, 400), extent=(0, 0, 400, 300))
    band, names = core.group_keys([str(i % 3) for i in range(n)])
    dense, _ = core.rasterize_dense(g, ri, fun, "float32", burn, None, band, 3, 2.0)
    sp = core.rasterize_sparse(g, ri, fun, "float32", burn, None, band, 3, 2.0)
    replay = core.sparse_build_array(ri, fun, 2.0, sp["counts"], sp["rows"], sp["cols"], sp["data"])
    assert np.array_equal(dense, replay, equal_nan=True)
    og = oracle.Geoms.from_wkb(geoms)
    ori = oracle.raster_info(og, shape=(300, 400), extent=(0, 0, 400, 300))
    assert np.array_equal(oracle.sparse_replay(ori, sp, fun, 2.0), replay, equal_nan=True)


def test_replay_duplicate_pixels_in_order():
    # many writes to few pixels: order-dependent functions must replay strictly in triplet order
    rng = np.random.default_rng(3)
    n = 5000
    rows, cols = rng.integers(0, 3, n).astype(np.uint64), rng.integers(0, 4, n).astype(np.uint64)
    data = rng.integers(0, 4, n).astype(np.int32)
    ri = core.raster_info(None, shape=(3, 4), extent=(0, 0, 4, 3))
    ori = oracle.raster_info(None, shape=(3, 4), extent=(0, 0, 4, 3))
    counts = np.array([3000, 2000], np.uint64)
    for fun in oracle.FUNS:
        got = core.sparse_build_array(ri, fun, 2, counts, rows, cols, data)
        exp = oracle.sparse_replay(ori, dict(counts=counts, rows=rows, cols=cols, data=data), fun, 2)
        assert np.array_equal(exp, got), fun


def test_inputs_on_device_flag_equals_host_inputs():
    """RZ_FLAG_INPUTS_ON_DEVICE: field values, validity bytes and band ids read from device memory give the same
    raster as the host arrays (nothing is copied per call), dense through both engines."""
    import torch

    import synth

    n = 1500
    x, y, off = synth.star_polygons(51, n, 6, 30, 30.0, 640, 480)
    g = core.Geoms.from_polygons(x, y, off)
    ri = core.raster_info(None, shape=(480, 640), extent=(0, 0, 640, 480))
    rng = np.random.default_rng(51)
    vals = rng.integers(1, 99, n).astype(np.int32)
    valid = (rng.random(n) < 0.9).astype(np.uint8)
    band, names = core.group_keys([str(i % 4) for i in range(n)])
    d_vals, d_valid, d_band = (torch.from_numpy(a).cuda() for a in (vals, valid, band))
    torch.cuda.synchronize()
    dev = dict(field=d_vals.data_ptr(), valid=d_valid.data_ptr(), band=d_band.data_ptr())
    for flags in (0, _lib.FLAG_NO_TILE_ENGINE):
        exp, _ = core.rasterize_dense(g, ri, "max", "int32", vals, valid, band, 4, 0, flags=flags)
        got, st = core.rasterize_dense(g, ri, "max", "int32", 0, None, None, 4, 0, flags=flags, inputs_dev=dev)
        assert np.array_equal(exp, got) and exp.shape == (4, 480, 640)
        assert st["h2d_bytes"] == 0  # geometry cached, inputs resident
    one = torch.tensor([2.5], dtype=torch.float32, device="cuda")
    exp, _ = core.rasterize_dense(g, ri, "sum", "float32", 2.5, background=np.nan)
    got, _ = core.rasterize_dense(g, ri, "sum", "float32", 0, background=np.nan, inputs_dev=dict(field=one.data_ptr(), scalar=True))
    assert np.array_equal(exp, got, equal_nan=True)
    with pytest.raises(ValueError, match="host inputs"):
        core.rasterize_dense(g, ri, "sum", "float32", 0, background=np.nan, inputs_dev=dict(field=one.data_ptr(), scalar=True),
                             devices=[0])


def test_r_array_layout_flag(monkeypatch):
    """RZ_FLAG_OUT_ROW_COL_BAND: the raster in R's (row, col, band) column-major layout, i.e. C-order [band][col][row]
    (R/rusterize/src/rust/src/encoding/rarrays.rs:9-17 permutes [band][row][col] with axes [0, 2, 1]) - host output in
    several row windows, device output, row shards, several devices writing one array; i32 and f64 are the R dtypes."""
    import torch

    W, H = 333, 257
    geoms = synth.mixed_geometries(52, 300, W, H, rho=30.0)
    n = len(geoms)
    by = [str(i % 3) for i in range(n)]
    band, names = core.group_keys(by)
    g = core.Geoms.from_wkb(geoms)
    ri = core.raster_info(g, shape=(H, W), extent=(0, 0, W, H))
    og = oracle.Geoms.from_wkb(geoms)
    ori = oracle.raster_info(og, shape=(H, W), extent=(0, 0, W, H))
    RCB = _lib.FLAG_OUT_ROW_COL_BAND
    monkeypatch.setenv("RZ_WINDOW_BYTES", str(40 * W * 3))  # 40 rows of the 1-byte dtype, 5 of the 8-byte one
    monkeypatch.setenv("RZ_ALLOW_REPEATED_DEVICES", "1")
    for dtype, bg, fun in (("int32", -2147483648, "sum"), ("float64", np.nan, "max"), ("uint8", 0, "count")):
        vals = (np.arange(n) % 9 + 1).astype(dtype)
        exp, _ = oracle.rasterize_dense(og, ori, fun, dtype, vals, None, by, bg)
        want = np.ascontiguousarray(exp.transpose(0, 2, 1))
        got, st = core.rasterize_dense(g, ri, fun, dtype, vals, None, band, 3, bg, flags=RCB)
        assert got.shape == (3, W, H) and st["n_windows"] > 2
        assert np.array_equal(want, got, equal_nan=True), dtype
        d_out = torch.zeros((3, W, H), dtype=getattr(torch, dtype), device="cuda")
        core.rasterize_dense(g, ri, fun, dtype, vals, None, band, 3, bg, flags=RCB, out=d_out.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(want, d_out.cpu().numpy(), equal_nan=True), dtype
        shard, _ = core.rasterize_dense(g, ri, fun, dtype, vals, None, band, 3, bg, flags=RCB, rows=(31, 200))
        assert np.array_equal(want[:, :, 31:200], shard, equal_nan=True), dtype
        multi, _ = core.rasterize_dense(g, ri, fun, dtype, vals, None, band, 3, bg, flags=RCB, devices=[0, 0, 0])
        assert np.array_equal(want, multi, equal_nan=True), dtype


def test_flatten_and_upload_at_once_equals_the_two_step_route():
    """rz_geoms_from_soa_to: pools copied to the device while they are flattened; the first call pays no upload and
    the raster equals the two-step route's (several threads, closing vertices, all three pools)."""
    geoms = synth.mixed_geometries(53, 3000, 900, 700, rho=35.0)
    soa = synth.wkb_to_soa(geoms)
    ri = core.raster_info(None, shape=(700, 900), extent=(0, 0, 900, 700))
    vals = (np.arange(len(geoms)) % 11 + 1).astype(np.float32)
    a = core.Geoms.from_soa(*soa)
    b = core.Geoms.from_soa(*soa, device=0)
    exp, st_a = core.rasterize_dense(a, ri, "sum", "float32", vals, background=np.nan)
    got, st_b = core.rasterize_dense(b, ri, "sum", "float32", vals, background=np.nan)
    assert np.array_equal(exp, got, equal_nan=True)
    assert st_a["h2d_bytes"] > 16 * a.n_coords and st_b["h2d_bytes"] == vals.nbytes  # b was already resident
    x, y, off = synth.star_polygons(53, 200000, 10, 40, 8.0, 3000, 3000)  # > 4 MB per pool: page-locked blocks, many ranges
    big = core.Geoms.from_polygons(x, y, off)
    idx = np.arange(len(off), dtype=np.uint64)
    big_dev = core.Geoms.from_soa(idx, np.zeros(len(off) - 1, np.uint8), idx, off, x, y, device=0)
    ri2 = core.raster_info(None, shape=(3000, 3000), extent=(0, 0, 3000, 3000))
    e2, _ = core.rasterize_dense(big, ri2, "count", "uint16", 1, background=0)
    g2, s2 = core.rasterize_dense(big_dev, ri2, "count", "uint16", 1, background=0)
    assert np.array_equal(e2, g2) and s2["h2d_bytes"] <= 8


def test_host_empty_outputs_are_recycled_page_locked_blocks(monkeypatch):
    """rz_host_alloc / rz_host_free through core.host_empty: an array over library memory is filled by a call exactly
    like a numpy array, large outputs are allocated there by default, and a freed block is handed out again."""
    W, H = 700, 500
    x, y, off = synth.star_polygons(41, 300, 8, 30, 40.0, W, H)
    vals = (synth.splitmix_u(41, 300, 9) * 100).astype(np.int32)
    g = core.Geoms.from_polygons(x, y, off)
    ri = core.raster_info(None, shape=(H, W), extent=(0, 0, W, H))
    ref, _ = core.rasterize_dense(g, ri, "max", "int32", vals, background=-1, out=np.empty((1, H, W), np.int32))
    a = core.host_empty((1, H, W), "int32")
    assert a.flags.writeable and a.flags.c_contiguous and a.shape == (1, H, W)
    got, _ = core.rasterize_dense(g, ri, "max", "int32", vals, background=-1, out=a)
    assert got is a and np.array_equal(ref, a)
    addr = a.ctypes.data
    del a, got
    import gc

    gc.collect()
    b = core.host_empty((1, H, W), "int32")
    assert b.ctypes.data == addr  # the block went back to the pool and came out again
    del b
    monkeypatch.setattr(core, "_HOST_EMPTY_MIN", 0)
    auto, _ = core.rasterize_dense(g, ri, "max", "int32", vals, background=-1)
    assert np.array_equal(ref, auto)
    keep = auto[0, 10:20].copy()
    view = auto[0, 10:20]
    del auto
    gc.collect()
    assert np.array_equal(view, keep)  # a view keeps the block alive
    soa = (np.arange(301, dtype=np.uint64), np.zeros(300, np.uint8), np.arange(301, dtype=np.uint64), off, x, y)
    one, _ = core.rasterize_dense_soa(soa, ri, "max", "int32", vals, background=-1, devices=[0])
    assert np.array_equal(ref, one)
    assert core.host_empty((0, 5), "float32").shape == (0, 5)


@pytest.mark.parametrize("piece", ["4096", "100000", str(128 << 20)])
def test_pageable_outputs_through_bounce_blocks(monkeypatch, piece):
    """A pageable destination (what a binding's own Array3 / numpy array is) is filled through two page-locked bounce
    blocks and host copy threads once it is large; forced here for small rasters, with pieces smaller than a row, a
    few rows and the whole window: bands, a row window, row-band shards and the one-shot call all land bit-exact."""
    monkeypatch.setenv("RZ_BOUNCE_MIN_BYTES", "0")
    monkeypatch.setenv("RZ_BOUNCE_BYTES", piece)
    monkeypatch.setenv("RZ_ALLOW_REPEATED_DEVICES", "1")
    W, H = 517, 389
    geoms = synth.mixed_geometries(43, 300, W, H, rho=40.0)
    n = len(geoms)
    by = [str(i % 3) for i in range(n)]
    band, names = core.group_keys(by)
    g = core.Geoms.from_wkb(geoms)
    og = oracle.Geoms.from_wkb(geoms)
    kw = dict(shape=(H, W), extent=(0, 0, W, H))
    ri, ori = core.raster_info(g, **kw), oracle.raster_info(og, **kw)
    vals = (np.arange(n) % 11 + 1).astype(np.float64)
    exp, _ = oracle.rasterize_dense(og, ori, "sum", "float64", vals, None, by, np.nan)
    out = np.full((3, H, W), 7.0)  # pageable
    got, st = core.rasterize_dense(g, ri, "sum", "float64", vals, None, band, 3, np.nan, out=out)
    assert np.array_equal(exp, got, equal_nan=True)
    assert st["host_syncs"] > 1  # one wait per piece
    win, _ = core.rasterize_dense(g, ri, "sum", "float64", vals, None, band, 3, np.nan, rows=(100, 333))
    assert np.array_equal(exp[:, 100:333], win, equal_nan=True)
    many, _ = core.rasterize_dense(g, ri, "sum", "float64", vals, None, band, 3, np.nan, devices=[0, 0, 0])
    assert np.array_equal(exp, many, equal_nan=True)
    one, _ = core.rasterize_dense_soa(synth.wkb_to_soa(geoms), ri, "sum", "float64", vals, None, band, 3, np.nan,
                                      devices=[0, 0], out=np.empty((3, H, W)))
    assert np.array_equal(exp, one, equal_nan=True)
    monkeypatch.setenv("RZ_BOUNCE", "0")
    plain, st0 = core.rasterize_dense(g, ri, "sum", "float64", vals, None, band, 3, np.nan, out=np.empty((3, H, W)))
    assert np.array_equal(exp, plain, equal_nan=True)
