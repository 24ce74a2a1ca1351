"""Output wrapping of the binding layer: `build_xarray` (python/src/encoding/xarray.rs:10-58 with the
pixel-centre coordinates of python/src/geo/raster.rs:45-62) and `SparseArray.to_frame`
(rust/src/encoding/arrays.rs:184-207).  xarray / rioxarray / polars are not installed in this image: minimal stand-in
modules record exactly what the wrappers hand to `xarray.DataArray.from_dict`, `.rio.write_crs` and
`polars.DataFrame`, so that the code has run and its arguments are checked."""
import importlib
import sys
import types

import numpy as np
import pytest

R = importlib.import_module("rusterize_b200._rusterize")  # (the package re-exports a function of the same name)
from rusterize_b200._lib import RasterInfo


@pytest.fixture()
def standins(monkeypatch):
    calls = {}

    class Rio:
        def __init__(self, owner):
            self.owner = owner

        def write_crs(self, epsg):
            self.owner.crs = epsg
            return self.owner

    class DataArray:
        crs = None

        def __init__(self, d):
            self.d = d
            self.rio = Rio(self)

        @classmethod
        def from_dict(cls, d):
            calls["from_dict"] = d
            return cls(d)

    class DataFrame:
        def __init__(self, cols):
            self.columns = list(cols)
            self.data = dict(cols)

    xr = types.ModuleType("xarray")
    xr.DataArray = DataArray
    pl = types.ModuleType("polars")
    pl.DataFrame = DataFrame
    monkeypatch.setitem(sys.modules, "xarray", xr)
    monkeypatch.setitem(sys.modules, "rioxarray", types.ModuleType("rioxarray"))
    monkeypatch.setitem(sys.modules, "polars", pl)
    return calls


def _ri(nrows, ncols, xmin, ymin, xres, yres, epsg):
    ri = RasterInfo()
    ri.nrows, ri.ncols = nrows, ncols
    ri.xmin, ri.ymin, ri.xres, ri.yres = xmin, ymin, xres, yres
    ri.xmax, ri.ymax = xmin + ncols * xres, ymin + nrows * yres
    ri.epsg = epsg
    return ri


def test_build_xarray_dims_pixel_centre_coords_and_crs(standins):
    ri = _ri(3, 4, 10.0, -6.0, 0.5, 2.0, 32632)
    data = np.arange(2 * 3 * 4, dtype=np.float32).reshape(2, 3, 4)
    out = R.build_xarray(ri, data, ["a", "b"])
    d = standins["from_dict"]
    assert d["dims"] == ["bands", "y", "x"] and d["data"] is data
    # ndarray::Array::range(start, end, step): ymax - yres/2 downwards, xmin + xres/2 upwards (pixel centres)
    assert np.array_equal(d["coords"]["y"]["data"], [-1.0, -3.0, -5.0]) and d["coords"]["y"]["dims"] == "y"
    assert np.array_equal(d["coords"]["x"]["data"], [10.25, 10.75, 11.25, 11.75]) and d["coords"]["x"]["dims"] == "x"
    assert d["coords"]["bands"] == {"dims": "bands", "data": ["a", "b"]}
    assert out.crs == 32632
    # no EPSG: write_crs is not called
    assert R.build_xarray(_ri(3, 4, 10.0, -6.0, 0.5, 2.0, -1), data, ["a", "b"]).crs is None
    # Array::range's element count is ceil((end - start) / step): a 1 x 1 grid has one centre per axis
    R.build_xarray(_ri(1, 1, 0.0, 0.0, 3.0, 3.0, -1), data[:1, :1, :1], ["band_1"])
    d = standins["from_dict"]
    assert np.array_equal(d["coords"]["y"]["data"], [1.5]) and np.array_equal(d["coords"]["x"]["data"], [1.5])


def test_sparse_to_frame_columns_and_one_based_bands(standins):
    ri = _ri(4, 4, 0.0, 0.0, 1.0, 1.0, -1)
    rows = np.array([0, 1, 1, 3, 2], np.uint64)
    cols = np.array([0, 1, 2, 3, 2], np.uint64)
    data = np.array([5, 5, 5, 7, 7], np.int32)
    sp = R.SparseArray(ri, ["x", "y"], rows, cols, data, np.array([3, 2], np.uint64), "sum", 0)
    f = sp.to_frame()
    assert f.columns == ["band", "row", "col", "values"]  # arrays.rs:189-204
    assert np.array_equal(f.data["band"], [1, 1, 1, 2, 2]) and f.data["band"].dtype == np.uint64
    assert f.data["row"] is rows and f.data["col"] is cols and f.data["values"] is data
    one = R.SparseArray(ri, ["band_1"], rows, cols, data, np.array([5], np.uint64), "sum", 0).to_frame()
    assert one.columns == ["row", "col", "values"]  # no band column for a single band
    assert sp.shape() == (2, 4, 4) and sp.extent() == (0.0, 0.0, 4.0, 4.0) and sp.epsg() is None
    assert "Estimated size: 128 bytes" in repr(sp)


@pytest.mark.gpu
def test_rusterize_front_end_xarray_and_sparse_frame_on_the_fixture(standins):
    """The reference's own fixture through `_rusterize` with encoding='xarray' and 'sparse' (to_xarray / to_frame on
    the result): python/test/test_many.py:228-235 and the frame of python/docs/python.md:106-136."""
    import cases
    from oracle.wkt2wkb import wkt_to_wkb

    wkbs = [wkt_to_wkb(w) for w in cases.GEOMS]
    raw = dict(shape=None, extent=None, resolution=(1.0, 1.0), tap=False, epsg=4326)
    burn = np.arange(1, 6, dtype=np.uint8)
    xa = R._rusterize(wkbs, raw, "sum", pyburn=burn, pybackground=0, pyencoding="xarray", pydtype="uint8")
    d = standins["from_dict"]
    assert d["data"].shape == (1, 131, 361) and int(d["data"].sum()) == 59204 and xa.crs == 4326
    assert d["coords"]["x"]["data"][0] == -180.0 and d["coords"]["y"]["data"][0] == 60.0  # centres of the buffered grid
    sp = R._rusterize(wkbs, raw, "sum", pyburn=burn, pybackground=0, pyencoding="sparse", pydtype="uint8")
    f = sp.to_frame()
    assert f.columns == ["row", "col", "values"] and len(f.data["row"]) == 29363
    assert list(zip(f.data["row"][:5].tolist(), f.data["col"][:5].tolist())) == [(6, 40), (6, 41), (6, 42), (7, 39), (7, 40)]
    sp.to_xarray()
    assert np.array_equal(standins["from_dict"]["data"], d["data"])  # replayed triplets == the dense raster
