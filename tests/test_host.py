"""CPU-only checks of the product's host side: the C ABI loads and exports every declared symbol,
the WKB / WKT flatteners, grid math and band grouping agree with the oracle, and compute entry
points fail loudly (no CPU fallback) without a GPU."""
import re
from pathlib import Path

import os

import numpy as np
import pytest

import oracle
from cases import GEOMS, GEOMS_EXPLODED, sq
from oracle.wkt2wkb import wkt_to_wkb
from rusterize_b200 import _lib, core

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    hdr = (ROOT / "include" / "rz_b200.h").read_text()
    declared = set(re.findall(r"\b(rz_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"librz_b200.so does not export {name}"
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    assert b"sm_100a" in L.rz_version()


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    g = core.Geoms.from_wkt([sq(0, 0, 4, 4)])
    ri = core.raster_info(g, shape=(4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        core.rasterize_dense(g, ri)
    # the one-shot call, the sparse call and the page-locked allocator fail the same way
    soa = (np.arange(2, dtype=np.uint64), np.zeros(1, np.uint8), np.arange(2, dtype=np.uint64),
           np.array([0, 5], np.uint64), np.array([0.0, 4, 4, 0, 0]), np.array([0.0, 0, 4, 4, 0]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        core.rasterize_dense_soa(soa, ri, devices=[0])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        core.rasterize_sparse(g, ri)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        core.host_empty((1 << 20,), "float32")
    assert core.host_trim(0) == 0  # (nothing pooled, nothing to give back: works without a device)


def test_flatten_parts_and_pooling():
    # burn_geometry.rs:24-210 — one part per Polygon/MultiPolygon/LineString/MultiLineString/
    # Point/MultiPoint; collections contribute their members' parts in order.
    g = core.Geoms.from_wkt(GEOMS)
    kind, geom = g.parts()
    assert len(g) == 5
    assert kind.tolist() == [0, 0, 0, 1, 2, 0, 1, 0]
    assert geom.tolist() == [0, 1, 2, 3, 4, 4, 4, 4]
    x, y, tag = g.pool(0)
    # first polygon: exterior 5 + hole 4 vertices, both rings of part 0, ring ends flagged
    assert (tag[:9] & 0x3FFFFFFF).tolist() == [0] * 9
    assert [i for i in range(9) if tag[i] & 0x80000000] == [4, 8]
    lx, ly, ltag = g.pool(1)
    assert len(lx) == 18 + 2 and sum(1 for t in ltag if t & 0x80000000) == 10
    assert not any(t & 0x40000000 for t in ltag)  # no closed line string
    assert g.bounds() == (-180.0, -70.0, 180.0, 60.0)


def test_wkt_and_wkb_readers_agree():
    for geoms in (GEOMS, GEOMS_EXPLODED, ["MULTIPOINT ((1 2), (3 4))", "MULTIPOINT (1 2, 3 4)",
                                          "POLYGON Z ((0 0 1, 4 0 1, 4 4 1, 0 0 1))", "LINESTRING (0 0, 1 1, 0 0)",
                                          "POINT (1e-3 -2.5E2)", "GEOMETRYCOLLECTION EMPTY", "POLYGON EMPTY"]):
        a = core.Geoms.from_wkt(geoms)
        b = core.Geoms.from_wkb([wkt_to_wkb(s) for s in geoms])
        assert len(a) == len(b)
        assert all(np.array_equal(p, q) for p, q in zip(a.parts(), b.parts()))
        for k in range(3):
            assert all(np.array_equal(p, q) for p, q in zip(a.pool(k), b.pool(k)))


def test_polygon_rings_are_closed_and_line_closed_flag():
    g = core.Geoms.from_wkt(["POLYGON ((0 0, 4 0, 4 4))", "LINESTRING (0 0, 1 1, 0 0)", "LINESTRING (0 0, 1 1)"])
    x, y, tag = g.pool(0)
    assert list(zip(x, y)) == [(0, 0), (4, 0), (4, 4), (0, 0)]  # geo_types::Polygon::new closes rings
    lx, ly, lt = g.pool(1)
    assert [bool(t & 0x40000000) for t in lt] == [True, True, True, False, False]


def test_big_endian_and_ewkb():
    import struct

    be = struct.pack(">BI", 0, 2) + struct.pack(">I", 2) + struct.pack(">dddd", 1.0, 2.0, 3.0, 4.0)
    ewkb = struct.pack("<BI", 1, 0x20000001) + struct.pack("<I", 4326) + struct.pack("<dd", 5.0, 6.0)
    g = core.Geoms.from_wkb([be, ewkb])
    assert g.pool(1)[0].tolist() == [1.0, 3.0] and g.pool(2)[0].tolist() == [5.0]


def test_dropped_and_invalid_inputs():
    # python/src/geo/parse_geometry.rs: empty points are dropped, all-dropped input is a ValueError
    g = core.Geoms.from_wkb([wkt_to_wkb("POINT EMPTY"), wkt_to_wkb("POINT (1 1)")])
    assert len(g) == 1
    with pytest.raises(ValueError, match="Could not parse geometry"):
        core.Geoms.from_wkb([wkt_to_wkb("POINT EMPTY")])
    with pytest.raises(RuntimeError, match="Cannot parse geometry"):
        core.Geoms.from_wkb([b"\x01\x03\x00\x00\x00\x05"])
    with pytest.raises(RuntimeError, match="Cannot parse geometry"):
        core.Geoms.from_wkt(["POLYGON ((0 0, 1 1"])
    with pytest.raises(ValueError, match="No geometries found"):
        core.Geoms.from_any([])


@pytest.mark.parametrize("kw", [
    dict(resolution=(1, 1)), dict(resolution=(0.5, 2.0)), dict(shape=(47, 319)), dict(resolution=(7, 3), tap=True),
    dict(shape=(10, 20), extent=(-349, -507, 1, 0)), dict(resolution=(1, 1), extent=(-349, -507, 1, 0)),
    dict(resolution=(0.3, 0.7), extent=(-10.05, -3.3, 17.77, 9.1), tap=True)])
def test_raster_info_matches_oracle(kw):
    g = core.Geoms.from_wkt(GEOMS)
    og = oracle.Geoms.from_any(GEOMS)
    a = core.raster_info(g, **kw)
    b = oracle.raster_info(og, **kw)
    assert (a.nrows, a.ncols, a.xmin, a.ymin, a.xmax, a.ymax, a.xres, a.yres) == b.as_tuple()


def test_raster_info_errors():
    g = core.Geoms.from_wkt([sq(0, 0, 4, 4)])
    cases = [
        (dict(), ValueError, "Must set at least one of `shape` or `resolution`"),
        (dict(shape=(4, 4), resolution=(1, 1)), ValueError, "Shape and resolution are mutually exclusive"),
        (dict(shape=(0, 4)), ValueError, "Shape values must be > 0."),
        (dict(resolution=(0.0, 1.0)), ValueError, "Resolution values must be > 0."),
        (dict(shape=(4, 4), extent=(0, 0, 0, 0)), ValueError, "Unspecified extent"),
    ]
    for kw, exc, msg in cases:
        with pytest.raises(exc, match=msg):
            core.raster_info(g, **kw)
    empty = core.Geoms.from_wkt(["GEOMETRYCOLLECTION EMPTY"])
    with pytest.raises(RuntimeError, match="Cannot infer bounding box from geometry."):
        core.raster_info(empty, shape=(4, 4))


def test_group_keys_matches_oracle():
    keys = [str(i % 32) for i in range(200)] + ["é", "a", "B", ""]
    band, names = core.group_keys(keys)
    oband, onames = oracle.group_keys(keys)
    assert names == onames and np.array_equal(band, oband)
    assert names[:4] == ["", "0", "1", "10"]  # byte-lexicographic (rasterize.rs:199-205)


def test_threaded_wkb_ingestion_equals_serial(monkeypatch):
    """rz_geoms_from_wkb parses large inputs with several threads (contiguous ranges, appended in order): the
    flattened form - pools, tags, parts table, bounds - must be identical to the serial walk, dropped geometries
    (POINT EMPTY) and collections included."""
    import synth
    from oracle import wkt2wkb as W

    geoms = synth.mixed_geometries(17, 900, 512, 512)
    geoms[5] = W.wkt_to_wkb("POINT EMPTY")
    geoms[450] = W.wkt_to_wkb("GEOMETRYCOLLECTION (POINT (1 2), LINESTRING (0 0, 3 3), POLYGON ((0 0, 4 0, 4 4, 0 0)))")
    geoms[899] = W.wkt_to_wkb("POINT EMPTY")
    forms = []
    for threads in ("1", "2", "7"):
        monkeypatch.setenv("RZ_PARSE_THREADS", threads)
        g = core.Geoms.from_wkb(geoms)
        kind, pg = g.parts()
        forms.append((len(g), g.bounds(), kind, pg, [g.pool(k) for k in range(3)]))
    for f in forms[1:]:
        assert f[0] == forms[0][0] == 898 and f[1] == forms[0][1]
        assert np.array_equal(f[2], forms[0][2]) and np.array_equal(f[3], forms[0][3])
        for a, b in zip(f[4], forms[0][4]):
            assert all(np.array_equal(u, v) for u, v in zip(a, b))
    monkeypatch.setenv("RZ_PARSE_THREADS", "3")
    with pytest.raises(RuntimeError, match="Cannot parse geometry"):
        core.Geoms.from_wkb(geoms[:600] + [b"\x01\x03\x00\x00\x00\x01"] + geoms[600:])


def test_parallel_soa_ingestion_equals_wkb_flattening(monkeypatch):
    """rz_geoms_from_soa writes the pools with several threads at exact offsets (counting pass + copy/extents
    sweep): pools, lazily rebuilt tags, parts table and bounds must equal what the streaming Flattener makes of
    the same geometries (WKB route), for any thread count - unclosed rings, closed line strings, collections, empty
    sequences and non-finite coordinates included."""
    import synth
    from oracle import wkt2wkb as W

    geoms = synth.mixed_geometries(23, 700, 512, 512)
    geoms[3] = W.wkt_to_wkb("POLYGON ((0 0, 4 0, 4 4))")  # unclosed ring: closed by the flattener
    geoms[4] = W.wkt_to_wkb("POLYGON ((0 0, 4 0, 4 4, 0 0), (1 1, 2 1, 2 2))")
    geoms[5] = W.wkt_to_wkb("LINESTRING (0 0, 1 1, 0 0)")
    geoms[600] = W.wkt_to_wkb("GEOMETRYCOLLECTION (POINT (1 2), LINESTRING (0 0, 3 3), POLYGON ((0 0, 4 0, 4 4, 0 0)))")
    ref = core.Geoms.from_wkb(geoms)
    soa = synth.wkb_to_soa(geoms)
    rk, rg = ref.parts()
    for threads in ("1", "2", "5", "16"):
        monkeypatch.setenv("RZ_PARSE_THREADS", threads)
        g = core.Geoms.from_soa(*soa)
        k, pg = g.parts()
        assert len(g) == len(ref) and g.bounds() == ref.bounds()
        assert np.array_equal(k, rk) and np.array_equal(pg, rg)
        for kind in range(3):
            for a, b in zip(g.pool(kind), ref.pool(kind)):
                assert np.array_equal(a, b, equal_nan=True)
    # an empty sequence, an empty part and a NaN coordinate
    gpo = np.array([0, 1, 2, 3], np.uint64)
    kinds = np.array([0, 1, 0], np.uint8)
    pso = np.array([0, 2, 2, 3], np.uint64)
    sco = np.array([0, 0, 4, 7], np.uint64)
    x = np.array([0, 4, 4, 0, 1, np.nan, 3.0])
    y = np.array([0, 0, 4, 0, 1, 2, 3.0])
    g = core.Geoms.from_soa(gpo, kinds, pso, sco, x, y)
    px, py, pt = g.pool(0)
    assert len(g) == 3 and g.n_parts == 3 and len(px) == 4 + 4  # second ring gets its closing vertex (NaN != NaN)
    assert [i for i in range(8) if pt[i] & 0x80000000] == [3, 7] and (pt & 0x3FFFFFFF).tolist() == [0] * 4 + [2] * 4
    with pytest.raises(RuntimeError, match="Invalid part kind"):
        core.Geoms.from_soa(gpo, np.array([0, 9, 0], np.uint8), pso, sco, x, y)


def test_row_shard_keeps_order_indices_and_vertices():
    """rz_geoms_row_shard (the library-side cull of a row-band sharded job): the shard holds exactly the parts whose
    y-extent comes within a few rows of the band, in the original order, with their original geometry indices and
    vertices; parts of every kind are tested by their own extent."""
    import synth

    W, H = 400, 1000
    geoms = synth.mixed_geometries(29, 500, W, H, rho=20.0)
    g = core.Geoms.from_wkb(geoms)
    ri = core.raster_info(None, shape=(H, W), extent=(0, 0, W, H))
    kind, geom = g.parts()
    pools = [g.pool(k) for k in range(3)]
    # vertex range of every part inside its pool, from the tags (part id in the low 30 bits)
    ymin, ymax = np.full(len(kind), np.inf), np.full(len(kind), -np.inf)
    for k in range(3):
        _, py, pt = pools[k]
        ids = (pt & 0x3FFFFFFF).astype(np.int64)
        np.minimum.at(ymin, ids, py)
        np.maximum.at(ymax, ids, py)
    for r0, r1 in [(0, 250), (250, 500), (777, 1000), (0, 1000)]:
        sh = g.row_shard(ri, r0, r1)
        sk, sg = sh.parts()
        top, bot = H - ymax, H - ymin  # pixel rows of each part's extent (res 1)
        must = (bot >= r0) & (top <= r1)          # certainly needed
        may = (bot >= r0 - 3) & (top <= r1 + 3)   # allowed slack
        assert len(sh) == len(g)  # geometry indices are unchanged
        # the kept parts form a subsequence of the original parts table
        j = 0
        kept = np.zeros(len(kind), bool)
        for i in range(len(kind)):
            if j < len(sk) and sk[j] == kind[i] and sg[j] == geom[i] and may[i]:
                kept[i] = True
                j += 1
        assert j == len(sk)
        assert (kept | ~must).all()
        for k in range(3):
            px, py, pt = pools[k]
            sel = kept[(pt & 0x3FFFFFFF).astype(np.int64)]
            sx, sy, _ = sh.pool(k)
            assert np.array_equal(sx, px[sel]) and np.array_equal(sy, py[sel])
    with pytest.raises(ValueError, match="Invalid row shard"):
        g.row_shard(ri, 10, 10)


def test_band_shard_straight_from_soa_equals_shard_of_the_flattened_set():
    """rz_geoms_from_soa_rows (extents pass + flatten with a keep mask: what every device of a one-shot multi-GPU call
    does) gives exactly the geometry set rz_geoms_from_soa + rz_geoms_row_shard gives: same parts in the same order,
    same geometry indices, same pools - for any thread count."""
    import synth

    W, H = 400, 1000
    geoms = synth.mixed_geometries(33, 600, W, H, rho=20.0)
    soa = synth.wkb_to_soa(geoms)
    full = core.Geoms.from_soa(*soa)
    ri = core.raster_info(None, shape=(H, W), extent=(0, 0, W, H))
    for threads in ("1", "3", "16"):
        os.environ["RZ_PARSE_THREADS"] = threads
        try:
            for r0, r1, touched in [(0, 250, False), (250, 500, True), (777, 1000, False), (0, 1000, False), (499, 500, False)]:
                a = full.row_shard(ri, r0, r1, touched)
                b = core.Geoms.from_soa_rows(soa, ri, r0, r1, touched)
                assert len(a) == len(b) == len(full) and a.n_parts == b.n_parts
                for u, v in zip(a.parts(), b.parts()):
                    assert np.array_equal(u, v)
                for kind in range(3):
                    for u, v in zip(a.pool(kind), b.pool(kind)):
                        assert np.array_equal(u, v, equal_nan=True)
        finally:
            del os.environ["RZ_PARSE_THREADS"]


def test_band_shard_from_soa_keeps_parts_with_non_finite_ordinates_everywhere():
    """A part with a NaN ordinate has no usable y-extent: the straight-from-arrays shard keeps it in every band (a
    superset never changes a band's pixels - the part burns only where its finite edges are), the two-step shard cuts
    it by the extent of its finite vertices.  Both keep every finite part a band needs, in order."""
    x = np.array([1, 3, 3, 1, 1, 5, 7, 7, 5, 5.0])
    y = np.array([1, 1, 3, 3, 1, 90, 90, np.nan, 95, 90.0])
    soa = (np.arange(3, dtype=np.uint64), np.zeros(2, np.uint8), np.arange(3, dtype=np.uint64),
           np.array([0, 5, 10], np.uint64), x, y)
    ri = core.raster_info(None, shape=(100, 10), extent=(0, 0, 10, 100))
    full = core.Geoms.from_soa(*soa)
    for r0, r1, two_step, straight in [(0, 10, [1], [1]), (50, 60, [], [1]), (90, 100, [0], [0, 1])]:
        assert list(full.row_shard(ri, r0, r1).parts()[1]) == two_step
        assert list(core.Geoms.from_soa_rows(soa, ri, r0, r1).parts()[1]) == straight
