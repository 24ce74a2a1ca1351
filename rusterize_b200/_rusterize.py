"""Host-side mirror of the reference's binding crate (python/src/*.rs) on top of the C ABI.

`_rusterize(...)` keeps the exact signature, argument meaning and error behaviour of the PyO3
function of the same name (python/src/rusterize.rs:127-184) so that the reference's Python
front-end (python/python/rusterize/__init__.py) routes through it unchanged; the compute itself is
librz_b200.so (CUDA, sm_100a).  Nothing here falls back to a CPU implementation.
"""
from __future__ import annotations

import numpy as np

from . import core
from ._lib import DTYPES, FUNS


# ------------------------------------------------------------------------------------------------
# optional dependencies (mirrors python/python/rusterize/_dependencies.py)
# ------------------------------------------------------------------------------------------------
def _has_module(name: str) -> bool:
    import importlib.util

    try:
        return importlib.util.find_spec(name) is not None
    except (ModuleNotFoundError, ValueError):
        return False


def _xarray_available() -> bool:
    return _has_module("xarray") and _has_module("rioxarray")


def _polars_available() -> bool:
    return _has_module("polars")


def _mro_mentions(obj, module: str) -> bool:
    try:
        return any(f"{module}." in str(c) for c in type(obj).mro())
    except TypeError:
        return False


def _check_for_geopandas(obj) -> bool:
    return _has_module("geopandas") and _mro_mentions(obj, "geopandas")


def _check_for_polars_st(obj) -> bool:
    return _has_module("polars_st") and _mro_mentions(obj, "polars")


# ------------------------------------------------------------------------------------------------
# value extraction rules of the binding (python/src/rusterize.rs:50-53, 77-89)
# ------------------------------------------------------------------------------------------------
def _extract_scalar(obj, dt: np.dtype):
    """PyO3 `obj.extract::<N>()`: returns (ok, value)."""
    if dt.kind in "iu":
        if isinstance(obj, (int, np.integer)):  # __index__; bool is an int
            v = int(obj)
            info = np.iinfo(dt)
            if info.min <= v <= info.max:
                return True, dt.type(v)
        return False, None
    if isinstance(obj, (int, float, np.integer, np.floating)) and not isinstance(obj, (str, bytes)):
        with np.errstate(over="ignore"):
            return True, dt.type(float(obj))
    return False, None


def _background(pybackground, dt: np.dtype):
    """`pybackground.and_then(|b| b.extract().ok()).unwrap_or_default()` — a value that does not
    extract into the dtype (np.nan or -1 for uint8, ...) silently becomes 0."""
    if pybackground is None:
        return dt.type(0)
    ok, v = _extract_scalar(pybackground, dt)
    return v if ok else dt.type(0)


def _burn(pyburn, dt: np.dtype):
    if pyburn is None:
        return dt.type(1)  # FieldSource::Scalar(N::one())
    ok, v = _extract_scalar(pyburn, dt)
    if ok:
        return v
    if isinstance(pyburn, np.ndarray) and pyburn.ndim == 1 and pyburn.dtype == dt:
        return np.ascontiguousarray(pyburn)
    raise TypeError(f"`burn` cannot be converted to a scalar or a 1-D numpy array of dtype {dt.name}")


# ------------------------------------------------------------------------------------------------
# geometry ingestion (python/src/geo/parse_geometry.rs:36-150)
# ------------------------------------------------------------------------------------------------
def _shapely_to_wkb(obj):
    import shapely

    if not shapely.__version__.startswith("2"):
        raise ValueError("Shapely version 2 required")
    return shapely.to_wkb(obj, output_dimension=2, include_srid=False, flavor="iso")


def parse_geometry(obj) -> core.Geoms:
    if hasattr(obj, "geom_type"):  # geopandas GeoDataFrame / GeoSeries
        return core.Geoms.from_wkb(list(_shapely_to_wkb(obj)))
    if isinstance(obj, (list, np.ndarray)):
        if len(obj) == 0:
            raise ValueError("No geometries found.")
        first = obj[0]
        if isinstance(first, (bytes, np.bytes_, bytearray, memoryview)):
            return core.Geoms.from_wkb(list(obj))
        if isinstance(first, str):
            return core.Geoms.from_wkt([str(s) for s in obj])
        if hasattr(first, "geom_type"):  # list of shapely geometries
            return core.Geoms.from_wkb(list(_shapely_to_wkb(obj)))
        raise ValueError("Sequence must contain geometries as shapely Geometry, bytes (WKB), or string (WKT).")
    if _polars_available():
        import polars as pl

        if isinstance(obj, pl.Series):
            if obj.dtype == pl.Binary:
                return core.Geoms.from_wkb([b for b in obj.to_list() if b is not None])
            if obj.dtype == pl.String:
                return core.Geoms.from_wkt([s for s in obj.to_list() if s is not None])
            raise TypeError("Unsupported dtype for geometry column")
    raise TypeError("Unsupported geometry input type.")


# ------------------------------------------------------------------------------------------------
# outputs (python/src/encoding/{pyarray,xarray}.rs, python/src/geo/raster.rs:45-62)
# ------------------------------------------------------------------------------------------------
def _coordinates(ri):
    # ndarray::Array::range(start, end, step): ceil((end-start)/step) elements
    def arange(start, end, step):
        n = max(int(np.ceil((end - start) / step)), 0)
        return start + step * np.arange(n, dtype=np.float64)

    y = arange(ri.ymax - ri.yres / 2.0, ri.ymax - ri.nrows * ri.yres, -ri.yres)
    x = arange(ri.xmin + ri.xres / 2.0, ri.xmin + ri.ncols * ri.xres, ri.xres)
    return y, x


def build_xarray(ri, data: np.ndarray, band_names):
    import rioxarray  # noqa: F401
    import xarray as xr

    y, x = _coordinates(ri)
    out = xr.DataArray.from_dict({
        "data": data, "dims": ["bands", "y", "x"],
        "coords": {"x": {"dims": "x", "data": x}, "y": {"dims": "y", "data": y},
                   "bands": {"dims": "bands", "data": list(band_names)}}})
    if ri.epsg >= 0:
        out = out.rio.write_crs(ri.epsg)
    return out


class SparseArray:
    """COO triplets of every burned pixel write, per band, in burn order
    (rust/src/encoding/arrays.rs:63-95; python/src/encoding/pyarray.rs:108-140)."""

    def __init__(self, ri, band_names, rows, cols, data, counts, fun, background):
        self._ri, self._band_names = ri, list(band_names)
        self.rows, self.cols, self.data, self.counts = rows, cols, data, counts
        self._fun, self._bg = fun, background

    # -- accessors mirrored from SparseArray<N> --------------------------------------------------
    def shape(self):
        return (len(self._band_names), int(self._ri.nrows), int(self._ri.ncols))

    def extent(self):
        return (self._ri.xmin, self._ri.ymin, self._ri.xmax, self._ri.ymax)

    def resolution(self):
        return (self._ri.xres, self._ri.yres)

    def epsg(self):
        return None if self._ri.epsg < 0 else int(self._ri.epsg)

    def band_names(self):
        return list(self._band_names)

    def _size_hint(self) -> str:
        b, r, c = self.shape()
        nbytes = self.data.dtype.itemsize * b * r * c
        if nbytes < 1000:
            return f"{nbytes} bytes"
        if nbytes < 1_000_000:
            return f"{np.float32(nbytes) / np.float32(1000.0):.2f} KB"
        if nbytes < 1_000_000_000:
            return f"{np.float32(nbytes) / np.float32(1_000_000.0):.2f} MB"
        return f"{np.float32(nbytes) / np.float32(1_000_000_000.0):.2f} GB"

    def __repr__(self) -> str:
        def f(v):  # Rust `{:?}` of f64 always shows a decimal point
            return repr(float(v))

        ext = "(" + ", ".join(f(v) for v in self.extent()) + ")"
        res = "(" + ", ".join(f(v) for v in self.resolution()) + ")"
        return (f"SparseArray:\n- Shape: {self.shape()}\n- Extent: {ext}\n- Resolution: {res}\n"
                f"- EPSG: {self.epsg()}\n- Estimated size: {self._size_hint()}")

    def to_numpy(self) -> np.ndarray:
        """SparseArray::build_array (arrays.rs:103-143): replay the triplets on the GPU."""
        return core.sparse_build_array(self._ri, self._fun, self._bg, self.counts, self.rows, self.cols, self.data)

    def to_xarray(self):
        return build_xarray(self._ri, self.to_numpy(), self._band_names)

    def to_frame(self):
        """arrays.rs:184-207: columns [band (1-based, only if >1 band)], row, col, values."""
        import polars as pl

        cols = {}
        if len(self.counts) > 1:
            cols["band"] = np.repeat(np.arange(1, len(self.counts) + 1, dtype=np.uint64), self.counts.astype(np.int64))
        cols["row"], cols["col"], cols["values"] = self.rows, self.cols, self.data
        return pl.DataFrame(cols)


# ------------------------------------------------------------------------------------------------
# the binding entry point
# ------------------------------------------------------------------------------------------------
def _rusterize(geometry, raw_raster_info, pypixel_fn, pydf=None, pyfield=None, pyby=None, pyburn=None,
               pybackground=None, pytouched=False, pyencoding="xarray", pydtype="float64"):
    geoms = geometry if isinstance(geometry, core.Geoms) else parse_geometry(geometry)
    try:  # python/src/rusterize.rs:146-148: grid errors surface as RuntimeError
        ri = core.raster_info(geoms, shape=raw_raster_info.get("shape"), extent=raw_raster_info.get("extent"),
                              resolution=raw_raster_info.get("resolution"), tap=raw_raster_info.get("tap", False),
                              epsg=raw_raster_info.get("epsg"))
    except ValueError as e:
        raise RuntimeError(str(e)) from None
    if pypixel_fn not in FUNS:
        raise ValueError("Unknown pixel function")
    if pydtype not in DTYPES or pyencoding not in ("xarray", "numpy", "sparse"):
        raise NotImplementedError("Invalid dtype or encoding provided.")  # `unimplemented!` in the reference
    dt = np.dtype(pydtype)
    background = _background(pybackground, dt)

    field, field_valid, by = None, None, None
    if pydf is not None and (pyfield or pyby):
        import polars as pl

        if pyfield:
            pl_dtype = {"uint8": pl.UInt8, "uint16": pl.UInt16, "uint32": pl.UInt32, "uint64": pl.UInt64,
                        "int8": pl.Int8, "int16": pl.Int16, "int32": pl.Int32, "int64": pl.Int64,
                        "float32": pl.Float32, "float64": pl.Float64}[pydtype]
            col = pydf.get_column(pyfield).cast(pl_dtype)
            if col.null_count() > 0:  # FieldSource::Column with nulls: those geometries are skipped
                field_valid = (~col.is_null()).to_numpy().astype(np.uint8)
                col = col.fill_null(0)
            field = np.ascontiguousarray(col.to_numpy(), dtype=dt)
        if pyby:
            bycol = pydf.get_column(pyby).cast(pl.String)
            if bycol.null_count() > 0:
                raise RuntimeError("Found nulls in `by` column. Consider droppping them.")
            by = bycol.to_list()
    if field is None:
        field = _burn(pyburn, dt)

    band, names = (None, ["band_1"])
    if by is not None:
        band, names = core.group_keys(by)
    # every visible GPU takes a share of the job (row bands / geometry ranges); RZ_DEVICES narrows the list
    kw = dict(field=field, field_valid=field_valid, band_of_geom=band, n_bands=len(names), background=background,
              all_touched=bool(pytouched), devices=core.default_devices())
    try:  # python/src/rusterize.rs:121-123
        if pyencoding == "sparse":
            sp = core.rasterize_sparse(geoms, ri, pypixel_fn, pydtype, **kw)
            return SparseArray(ri, names, sp["rows"], sp["cols"], sp["data"], sp["counts"], pypixel_fn, background)
        arr, _ = core.rasterize_dense(geoms, ri, pypixel_fn, pydtype, **kw)
    except ValueError as e:
        raise RuntimeError(str(e)) from None
    if pyencoding == "xarray":
        return build_xarray(ri, arr, names)
    return arr
