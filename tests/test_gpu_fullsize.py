"""BASELINE.json configs 2 and 3 at their full sizes on the GPU.  The oracle is run on a row band of
the same grid (bit-exact comparison of those rows); the rest is checked through size-independent
properties (count/any consistency, min <= max, first/last are burned values, shards == full)."""
import numpy as np
import pytest

import oracle
import synth
from rusterize_b200 import _lib, core

pytestmark = pytest.mark.gpu


def test_config2_mixed_count_and_any():
    """BASELINE config 2 exactly as bench.py runs it (SplitMix64 seed 2, synth.config2_soa): 100k mixed
    geometries - 60 % star polygons, 25 % line strings, 15 % points / multipoints - on 16384 x 16384."""
    size = 16384
    soa = synth.config2_soa(2, 100_000, size)
    g = core.Geoms.from_soa(*soa)
    ri = core.raster_info(None, shape=(size, size), extent=(0, 0, size, size))
    cnt, st = core.rasterize_dense(g, ri, "count", "uint32", 1, background=0)
    anyv, _ = core.rasterize_dense(g, ri, "any", "uint8", 1, background=0)
    assert st["engine"] == 1  # count / any are order-free: polygons through the tile engine, lines and points by atomics
    assert np.array_equal(anyv[0] == 1, cnt[0] > 0)
    assert cnt.sum() > 0
    # oracle on a row band of the same grid (bit-exact path): the geometries that can touch it, in input order
    r0, r1 = 4096, 4096 + 1024
    sco, y = soa[3].astype(np.int64), soa[5]
    ymin, ymax = np.minimum.reduceat(y, sco[:-1]), np.maximum.reduceat(y, sco[:-1])
    keep = (size - ymax <= r1 + 3) & (size - ymin >= r0 - 3)
    og = oracle.Geoms.from_wkb(synth.soa_to_wkb(synth.soa_select(soa, keep)))
    ori = oracle.raster_info(None, shape=(r1 - r0, size), extent=(0, size - r1, size, size - r0))
    exp_c, _ = oracle.rasterize_dense(og, ori, "count", "uint32", 1, background=0)
    exp_a, _ = oracle.rasterize_dense(og, ori, "any", "uint8", 1, background=0)
    assert np.array_equal(exp_c[0], cnt[0, r0:r1]) and np.array_equal(exp_a[0], anyv[0, r0:r1])
    # a shard computed on its own equals the same rows of the full raster; so does the forced record pipeline
    shard, _ = core.rasterize_dense(g, ri, "count", "uint32", 1, background=0, rows=(r0, r1))
    assert np.array_equal(shard[0], cnt[0, r0:r1])
    rec, st_r = core.rasterize_dense(g, ri, "count", "uint32", 1, background=0, rows=(r0, r1), flags=_lib.FLAG_NO_TILE_ENGINE)
    assert st_r["engine"] == 0 and np.array_equal(rec[0], cnt[0, r0:r1])


def test_config3_32_layers_first_last_min_max():
    size, n = 8192, 100_000
    x, y, off = synth.star_polygons(3, n, 64, 64, 256.0, size, size)
    idx = np.arange(n, dtype=np.int64)
    field = (1 + (idx * 2654435761 % 10**6)).astype(np.int32)
    by = [str(i % 32) for i in range(n)]
    band, names = core.group_keys(by)
    assert names[:4] == ["0", "1", "10", "11"] and len(names) == 32  # lexicographic band order
    g = core.Geoms.from_polygons(x, y, off)
    ri = core.raster_info(None, shape=(size, size), extent=(0, 0, size, size))
    res = {}
    rows = (2048, 2048 + 256)  # keep host memory modest: 32 bands x 256 rows x 8192 cols x 4 B = 268 MB per function
    for fun in ("first", "last", "min", "max"):
        res[fun], st = core.rasterize_dense(g, ri, fun, "int32", field, None, band, 32, 0, rows=rows)
    og = oracle.Geoms.from_rings(x, y, off)
    ori = oracle.raster_info(None, shape=(rows[1] - rows[0], size), extent=(0, size - rows[1], size, size - rows[0]))
    for fun in ("first", "last", "min", "max"):
        exp, onames = oracle.rasterize_dense(og, ori, fun, "int32", field, None, by, 0, threads=8)
        assert onames == names
        assert np.array_equal(exp, res[fun]), fun
    burned = res["last"] != 0
    assert (res["min"][burned] <= res["max"][burned]).all()
    assert np.isin(res["first"][burned][:100000], field).all()


def test_config5_sparse_parcels_and_geometry_range_shards():
    """BASELINE config 5 (sparse encoding of small parcels) at 1/25 of its size: 400k jittered quads over
    26214 x 26214 (same density as 10M over 131072 x 131072).  The triplet stream must equal the oracle's, and
    shards made of contiguous geometry ranges (sparse output is ordered band -> geometry -> burn
    order, so ranges concatenate) must reproduce it exactly."""
    n, size = 400_000, 26214
    x, y, off = synth.parcels(5, n, size, size)
    vals = synth.splitmix_u(5, n, 20).astype(np.float32)
    g = core.Geoms.from_polygons(x, y, off)
    ri = core.raster_info(None, shape=(size, size), extent=(0, 0, size, size))
    got = core.rasterize_sparse(g, ri, "sum", "float32", vals, background=np.nan)
    og = oracle.Geoms.from_rings(x, y, off)
    ori = oracle.raster_info(None, shape=(size, size), extent=(0, 0, size, size))
    exp = oracle.rasterize_sparse(og, ori, "sum", "float32", vals, None, None, np.nan)
    assert len(exp["rows"]) > 30 * n
    for k in ("counts", "rows", "cols", "data"):
        assert np.array_equal(exp[k], got[k]), k
    # 4 shards of contiguous geometry ranges
    parts = []
    for d in range(4):
        a, b = n * d // 4, n * (d + 1) // 4
        o = off[a:b + 1]
        gs = core.Geoms.from_polygons(x[int(o[0]):int(o[-1])], y[int(o[0]):int(o[-1])], o - o[0])
        parts.append(core.rasterize_sparse(gs, ri, "sum", "float32", vals[a:b], background=np.nan))
    for k in ("rows", "cols", "data"):
        assert np.array_equal(np.concatenate([p[k] for p in parts]), got[k]), k
    assert sum(int(p["counts"][0]) for p in parts) == int(got["counts"][0])
    # replaying the stream gives the dense raster of the same job (python/test/test_many.py:216-224)
    rows = (4096, 4096 + 512)
    dense, _ = core.rasterize_dense(g, ri, "sum", "float32", vals, background=np.nan, rows=rows)
    sel = (got["rows"] >= rows[0]) & (got["rows"] < rows[1])
    ri_band = core.raster_info(None, shape=(rows[1] - rows[0], size), extent=(0, size - rows[1], size, size - rows[0]))
    rep = core.sparse_build_array(ri_band, "sum", np.nan, np.array([int(sel.sum())], np.uint64),
                                  got["rows"][sel] - np.uint64(rows[0]), got["cols"][sel], got["data"][sel])
    assert np.array_equal(rep, dense, equal_nan=True)
