"""Tile-binned engine, second generation: compact (row x word ranged) mask blocks in part order, device-side
counts with buffer bounds cached per (geometry set, grid), no host synchronisation in a steady-state call.
Everything is compared with the CPU oracle (bit-exact)."""
import numpy as np
import pytest

import oracle
import synth
from oracle.wkt2wkb import wkt_to_wkb
from rusterize_b200 import _lib, core

pytestmark = pytest.mark.gpu
TILES = _lib.FLAG_FORCE_TILE_ENGINE


def _oracle(x, y, off, fun, dtype, vals, bg, valid=None, by=None, **kw):
    og = oracle.Geoms.from_rings(x, y, off)
    return oracle.rasterize_dense(og, oracle.raster_info(None, **kw), fun, dtype, vals, valid, by, bg)[0]


def test_plan_cache_and_steady_state_without_host_sync():
    import torch

    x, y, off = synth.star_polygons(21, 3000, 8, 40, 40.0, 1500, 1100)
    vals = (10 * synth.splitmix_u(21, 3000, 9)).astype(np.float32)
    kw = dict(shape=(1100, 1500), extent=(0, 0, 1500, 1100))
    exp = _oracle(x, y, off, "sum", "float32", vals, np.nan, **kw)
    g = core.Geoms.from_polygons(x, y, off)
    ri = core.raster_info(None, **kw)
    d_out = torch.empty((1, 1100, 1500), dtype=torch.float32, device="cuda")
    stream = torch.cuda.Stream()
    syncs = []
    for i in range(3):
        d_out.fill_(7.0)
        torch.cuda.synchronize()
        _, st = core.rasterize_dense(g, ri, "sum", "float32", vals, background=np.nan, out=d_out.data_ptr(),
                                     stream=stream.cuda_stream)
        stream.synchronize()
        assert st["engine"] == 1 and st["plan_cached"] == (1 if i else 0)
        syncs.append(st["host_syncs"])
        assert np.array_equal(exp[0], d_out[0].cpu().numpy(), equal_nan=True), i
    assert syncs[0] >= 1 and syncs[1] == 0 and syncs[2] == 0  # only the first call reads counts back
    # a different grid on the same handle plans again; the first grid's plan is still there
    kw2 = dict(shape=(550, 750), extent=(0, 0, 1500, 1100))
    got2, st2 = core.rasterize_dense(g, core.raster_info(None, **kw2), "count", "int32", 1, background=0, flags=TILES)
    assert st2["plan_cached"] == 0
    assert np.array_equal(_oracle(x, y, off, "count", "int32", 1, 0, **kw2), got2)
    _, st3 = core.rasterize_dense(g, ri, "sum", "float32", vals, background=np.nan, out=d_out.data_ptr(),
                                  stream=stream.cuda_stream)
    assert st3["plan_cached"] == 1 and st3["host_syncs"] == 0


def test_cached_bounds_cover_skipped_parts_and_changing_values():
    """The cached bounds count every polygon part; calls that skip parts (null fields) leave filler records that
    sort behind every tile, and the value-dependent fast paths of tile_apply are chosen on the device per call."""
    x, y, off = synth.star_polygons(22, 2500, 5, 30, 35.0, 900, 700)
    n = 2500
    kw = dict(shape=(700, 900), extent=(0, 0, 900, 700))
    g = core.Geoms.from_polygons(x, y, off)
    ri = core.raster_info(None, **kw)
    rng = np.random.default_rng(22)
    for it, (fun, dtype, bg) in enumerate([("sum", "float32", np.nan), ("sum", "float32", np.nan), ("min", "int32", 0),
                                           ("min", "int32", 0), ("first", "float64", np.nan), ("count", "float32", np.nan),
                                           ("max", "uint8", 0), ("last", "int16", -1)]):
        vals = rng.integers(1, 9, n).astype(dtype)
        if it % 2:  # values that break the fast path: NaN for floats, the background for integers
            vals[rng.integers(0, n, 40)] = np.nan if np.dtype(dtype).kind == "f" else bg
        valid = (rng.random(n) < (0.5 if it % 3 == 0 else 0.97)).astype(np.uint8)
        exp = _oracle(x, y, off, fun, dtype, vals, bg, valid=valid, **kw)
        got, st = core.rasterize_dense(g, ri, fun, dtype, vals, valid, background=bg, flags=TILES)
        assert st["engine"] == 1
        assert np.array_equal(exp, got, equal_nan=True), (fun, dtype, it)


@pytest.mark.parametrize("dtype", ["float32", "float64", "uint8"])
def test_wide_and_tall_parts_chunks_units_and_word_ranges(dtype):
    """Parts wider than one 512-column chunk (carry-in parity between chunks), word ranges that start anywhere in a
    tile, parts taller than a warp's toggle mask (several units), 1-pixel and sub-pixel parts, parts hanging over
    every raster edge; both tile heights (64 rows, 32 for 8-byte dtypes)."""
    W, H = 2300, 700
    polys = [
        "POLYGON ((3 3, 2290 10, 2200 690, 40 650, 3 3))",                                  # 5 chunks, 11 tile rows
        "POLYGON ((97 20, 1500 300, 97 600, 700 300, 97 20))",                              # concave, starts mid-tile
        "POLYGON ((515 5, 1030 5, 1030 695, 515 695, 515 5), (600 100, 900 100, 900 500, 600 500, 600 100))",
        "POLYGON ((-500 -300, 3000 -300, 3000 1000, -500 1000, -500 -300))",                # covers everything
        "POLYGON ((129.2 65.1, 129.9 65.1, 129.9 65.8, 129.2 65.8, 129.2 65.1))",           # sub-pixel: no centre inside
        "POLYGON ((127.4 63.4, 128.6 63.4, 128.6 64.6, 127.4 64.6, 127.4 63.4))",           # 4 tiles' corner
        "POLYGON ((2299.2 0, 2300 0, 2300 700, 2299.2 700, 2299.2 0))",                      # last column only
        "POLYGON ((1000 350, 1600 351, 2299 349, 1600 352, 1000 350))",                      # sliver across chunks
        "MULTIPOLYGON (((10 10, 60 10, 60 60, 10 60, 10 10)), ((2000 600, 2250 600, 2250 690, 2000 690, 2000 600)))",
    ]
    rng = np.random.default_rng(5)
    for _ in range(60):  # random boxes with arbitrary word alignment
        x0, y0 = rng.uniform(-50, W), rng.uniform(-50, H)
        w, h = rng.uniform(1, 900), rng.uniform(1, 300)
        polys.append(f"POLYGON (({x0} {y0}, {x0 + w} {y0 + 0.3 * h}, {x0 + 0.9 * w} {y0 + h}, {x0 + 0.1 * w} {y0 + 0.8 * h}, {x0} {y0}))")
    geoms = [wkt_to_wkb(p) for p in polys]
    vals = (np.arange(len(geoms)) % 5 + 1).astype(dtype)
    kw = dict(shape=(H, W), extent=(0, 0, W, H))
    og = oracle.Geoms.from_any(geoms)
    g = core.Geoms.from_any(geoms)
    ri = core.raster_info(g, **kw)
    for fun, bg in (("sum", np.nan if dtype != "uint8" else 0), ("first", 0), ("count", 0)):
        exp, _ = oracle.rasterize_dense(og, oracle.raster_info(og, **kw), fun, dtype, vals, None, None, bg)
        got, st = core.rasterize_dense(g, ri, fun, dtype, vals, background=bg, flags=TILES)
        assert st["engine"] == 1
        assert np.array_equal(exp, got, equal_nan=True), (fun, int(np.sum(exp != got)))
    # row shards and host windows over the same handle (each window has its own cached bounds)
    exp, _ = oracle.rasterize_dense(og, oracle.raster_info(og, **kw), "sum", dtype, vals, None, None, 0)
    for rows in ((0, 100), (100, 333), (333, 700)):
        got, _ = core.rasterize_dense(g, ri, "sum", dtype, vals, background=0, rows=rows, flags=TILES)
        assert np.array_equal(exp[:, rows[0]:rows[1]], got, equal_nan=True), rows


def test_bands_and_non_power_of_two_resolution():
    x, y, off = synth.star_polygons(23, 1500, 6, 24, 30.0, 1000, 800)
    n = 1500
    by = [str(i % 5) for i in range(n)]
    vals = (np.arange(n) % 11 - 5).astype(np.int32)
    kw = dict(shape=(611, 777), extent=(0, 0, 1000, 800))  # xres, yres are not powers of two
    og = oracle.Geoms.from_rings(x, y, off)
    exp, names = oracle.rasterize_dense(og, oracle.raster_info(None, **kw), "max", "int32", vals, None, by, -7)
    band, bn = core.group_keys(by)
    assert bn == names
    g = core.Geoms.from_polygons(x, y, off)
    got, st = core.rasterize_dense(g, core.raster_info(None, **kw), "max", "int32", vals, None, band, len(bn), -7,
                                   flags=TILES)
    assert st["engine"] == 1 and np.array_equal(exp, got)


def test_non_finite_coordinates_take_the_record_pipeline():
    """A ring with a NaN vertex has odd crossing counts on arbitrary rows; only the record pipeline drops the largest
    crossing of any odd row like chunks_exact(2) (burners.rs:305), so such geometry sets never take the tile engine
    and the result does not depend on a cost model."""
    polys = [np.array([[10, 10], [90, 10], [90, 90], [10, 90], [10, 10]], float) for _ in range(4)]
    polys[1] = polys[1] + 100.0
    polys[2] = np.array([[150, 20], [250, 30], [np.nan, 60], [240, 120], [160, 110], [150, 20]], float)
    from oracle.wkt2wkb import polygon_wkb

    geoms = [polygon_wkb([p]) for p in polys]
    g = core.Geoms.from_any(geoms)
    ri = core.raster_info(None, shape=(256, 300), extent=(0, 0, 300, 256))
    got, st = core.rasterize_dense(g, ri, "sum", "float32", 1.0, background=0, flags=TILES)
    assert st["engine"] == 0
    og = oracle.Geoms.from_any(geoms)
    exp, _ = oracle.rasterize_dense(og, oracle.raster_info(None, shape=(256, 300), extent=(0, 0, 300, 256)), "sum",
                                    "float32", 1.0, None, None, 0)
    # the NaN ring itself is a documented divergence (its crossings land on column 0); the rows it does not reach
    # (raster rows below 130) hold the finite polygons and must be exact
    assert np.array_equal(exp[:, :130], got[:, :130])


def test_two_streams_share_the_device_scratch_safely():
    """Back-to-back device-output calls on different streams: the second call waits for the first (an event), so
    the shared per-device scratch is never used by two streams at once."""
    import torch

    x, y, off = synth.star_polygons(31, 20000, 16, 64, 30.0, 3000, 3000)
    kw = dict(shape=(3000, 3000), extent=(0, 0, 3000, 3000))
    g = core.Geoms.from_polygons(x, y, off)
    ri = core.raster_info(None, **kw)
    va = np.ones(20000, np.float32)
    vb = np.full(20000, 2.0, np.float32)
    ea = _oracle(x, y, off, "sum", "float32", va, np.nan, **kw)
    eb = _oracle(x, y, off, "sum", "float32", vb, np.nan, **kw)
    oa = torch.empty((1, 3000, 3000), dtype=torch.float32, device="cuda")
    ob = torch.empty_like(oa)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    core.rasterize_dense(g, ri, "sum", "float32", va, background=np.nan, out=oa.data_ptr(), stream=s1.cuda_stream)  # plan
    for _ in range(3):
        core.rasterize_dense(g, ri, "sum", "float32", va, background=np.nan, out=oa.data_ptr(), stream=s1.cuda_stream)
        core.rasterize_dense(g, ri, "sum", "float32", vb, background=np.nan, out=ob.data_ptr(), stream=s2.cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(ea[0], oa[0].cpu().numpy(), equal_nan=True)
        assert np.array_equal(eb[0], ob[0].cpu().numpy(), equal_nan=True)


@pytest.mark.parametrize("shape", [(256, 512), (64, 128), (192, 384)])
def test_filler_records_behind_the_last_tile(shape):
    """The tile sort covers exactly the tile bits: filler records (parts skipped by a null field leave the record
    array short of its cached bound) carry an all-ones tile field, equal to the last tile's id when the tile count is
    a power of two.  They must stay behind that tile's real records (stable sort, fillers last in the input)."""
    H, W = shape
    n = 600
    x, y, off = synth.star_polygons(61, n, 5, 12, 40.0, W, H)
    # make sure the last tile (bottom-right) is busy
    x[: int(off[60])] = x[: int(off[60])] * 0.1 + (W - 60)
    y[: int(off[60])] = y[: int(off[60])] * 0.1 + 5
    kw = dict(shape=(H, W), extent=(0, 0, W, H))
    g = core.Geoms.from_polygons(x, y, off)
    ri = core.raster_info(None, **kw)
    rng = np.random.default_rng(61)
    vals = rng.integers(1, 50, n).astype(np.float32)
    for keep in (1.0, 0.5, 0.05):
        valid = (rng.random(n) < keep).astype(np.uint8)
        exp = _oracle(x, y, off, "sum", "float32", vals, np.nan, valid=valid, **kw)
        got, st = core.rasterize_dense(g, ri, "sum", "float32", vals, valid, background=np.nan, flags=TILES)
        assert st["engine"] == 1 and np.array_equal(exp, got, equal_nan=True), (shape, keep)
