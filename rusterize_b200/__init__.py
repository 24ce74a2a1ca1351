"""rusterize_b200 — B200-native (CUDA, sm_100a) implementation of rusterize's polygon / line /
point burn path behind the reference's own `rusterize()` surface
(/root/reference/python/python/rusterize/__init__.py:83-356).  See DESIGN.md / INTEGRATION.md.

    from rusterize_b200 import rusterize
    arr = rusterize(list_of_wkt_or_wkb, res=(1, 1), burn=values, fun="sum", encoding="numpy", dtype="uint8")
"""
from __future__ import annotations

from types import NoneType

import numpy as np

from . import core  # noqa: F401
from ._rusterize import (  # noqa: F401
    SparseArray,
    _check_for_geopandas,
    _check_for_polars_st,
    _polars_available,
    _rusterize,
    _xarray_available,
)
from .core import Geoms, group_keys, raster_info, rasterize_dense, rasterize_sparse  # noqa: F401

__version__ = "0.1.0"

_DTYPE_MSG = ("`dtype` must be a one of 'uint8', 'uint16', 'uint32', 'uint64', 'int8', 'int16', 'int32', 'int64', "
              "'float32', 'float64'")


def _type_checks(res, out_shape, extent, field, by, burn, fun, background, encoding, all_touched, tap, dtype):
    seq = (tuple, list, NoneType)
    table = [
        (res, seq, "`resolution` must be a tuple or list of (xres, yres)."),
        (out_shape, seq, "`out_shape` must be a tuple or list of (nrows, ncols)."),
        (extent, seq, "`extent` must be a tuple or list of (xmin, ymin, xmax, ymax)."),
        (field, (str, NoneType), "`field` must be a string column name."),
        (by, (str, NoneType), "`by` must be a string column name."),
        (burn, (int, float, np.ndarray, NoneType), "`burn` must be an integer, float, or a numpy.ndarray."),
        (fun, str, "`pixel_fn` must be one of sum, first, last, min, max, count, or any."),
        (background, (int, float, NoneType), "`background` must be integer, float, or None."),
        (encoding, str, "`encoding` must be one of 'xarray', 'numpy', or 'sparse'."),
        (all_touched, bool, "`all_touched` must be a boolean."),
        (tap, bool, "`tap` must be a boolean."),
        (dtype, str, _DTYPE_MSG),
    ]
    for value, types, msg in table:
        if not isinstance(value, types):
            raise TypeError(msg)


def rusterize(data, like=None, res=None, out_shape=None, extent=None, field=None, by=None, burn=None, fun="last",
              background=np.nan, encoding="xarray", all_touched=False, tap=False, dtype="float64"):
    """Same parameters, defaults, validation and return types as the reference's `rusterize()`;
    see its docstring (python/python/rusterize/__init__.py:99-148).  `data` may be a
    geopandas.GeoDataFrame / GeoSeries, a polars.DataFrame with a "geometry" column, or a list /
    numpy array of WKT strings, WKB bytes or shapely geometries."""
    if isinstance(data, (list, np.ndarray)):
        kind = "raw"
    elif _check_for_geopandas(data) and type(data).__name__ == "GeoSeries":
        kind = "geoseries"
    elif _check_for_geopandas(data) and type(data).__name__ == "GeoDataFrame":
        kind = "geopandas"
    elif _check_for_polars_st(data) and type(data).__name__ in ("DataFrame", "GeoDataFrame"):
        kind = "polars"
    else:
        raise TypeError("`data` must be either geopandas.GeoDataFrame, geopandas.GeoSeries, polars.DataFrame, list, "
                        "or numpy.ndarray")
    if kind in ("geoseries", "geopandas") and data.empty:
        raise ValueError("Input data is empty.")
    if kind == "polars" and data.is_empty():
        raise ValueError("Input data is empty.")

    _type_checks(res, out_shape, extent, field, by, burn, fun, background, encoding, all_touched, tap, dtype)

    if encoding not in ("xarray", "numpy", "sparse"):
        raise ValueError("`encoding` must be one of `xarray`, 'numpy', or `sparse`.")
    if encoding == "xarray" and not _xarray_available():
        raise ModuleNotFoundError("`xarray` and `rioxarray` must be installed if encoding is `xarray`. Install with "
                                  "`pip install xarray rioxarray`.")
    if field and burn is not None:
        raise ValueError("Only one of `field` or `burn` can be specified.")
    if isinstance(burn, np.ndarray) and burn.size != len(data):
        raise ValueError("If `burn` is a `numpy.ndarray`, it must have the same length as `data`.")

    bounds = shape = resolution = None
    if like is not None:
        ok = False
        if _xarray_available():
            import xarray as xr

            ok = isinstance(like, (xr.DataArray, xr.Dataset))
        if not ok:
            raise TypeError("`like` must be a xarray.DataArray or xarray.Dataset")
        if any((res, out_shape, extent)):
            raise ValueError("`like` is mutually exclusive with `res`, `out_shape`, and `extent`.")
        if not hasattr(like, "rio"):
            raise AttributeError("The `like` object must have a 'rio' accessor.")
        try:
            shape = like.squeeze().shape
            bounds = like.rio.bounds()
        except Exception as e:
            raise AttributeError("No spatial dimension found for like object") from e
    else:
        if not res and not out_shape and not extent:
            raise ValueError("One of `res`, `out_shape`, or `extent` must be provided.")
        if res and out_shape:
            raise ValueError("`res` and `out_shape` are mutually exclusive; provide only one.")
        if extent:
            if not res and not out_shape:
                raise ValueError("Must also specify `res` or `out_shape` with extent.")
            if len(extent) != 4 or all(e == 0 for e in extent):
                raise ValueError("`extent` must be a tuple or list of (xmin, ymin, xmax, ymax).")
            bounds = extent
        if res:
            if len(res) != 2 or any(r <= 0 for r in res) or any(not isinstance(r, (int, float)) for r in res):
                raise ValueError("`res` must be 2 positive numbers.")
            resolution = res
        if out_shape:
            if len(out_shape) != 2 or any(s <= 0 for s in out_shape) or any(not isinstance(s, int) for s in out_shape):
                raise ValueError("`out_shape` must be 2 positive integers.")
            shape = out_shape

    wanted = sorted({c for c in (field, by) if c and c != "geometry"})
    df = None
    epsg = None
    if kind == "geopandas":
        epsg = data.crs.to_epsg() if data.crs else None
        if wanted:
            if field and not by:  # single column: hand it over as the burn array
                try:
                    burn = data[field].to_numpy()
                    field = None
                except KeyError as e:
                    raise KeyError("Column not found in GeoDataFrame.") from e
            else:
                if not _polars_available():
                    raise ModuleNotFoundError("polars must be installed when data is geopandas.GeoDataFrame.")
                import polars as pl

                try:
                    df = pl.from_pandas(data[wanted])
                except KeyError as e:
                    raise KeyError("Column not found in GeoDataFrame.") from e
        geometries = data.geometry
    elif kind == "polars":
        import polars as pl

        try:
            srid = data.select(pl.col("geometry").first().st.srid()).item()
        except pl.exceptions.ColumnNotFoundError as e:
            raise ValueError("If `polars.DataFrame`, a 'geometry' column is expected.") from e
        epsg = None if srid == 0 else srid
        if wanted:
            try:
                df = data.select(pl.col([*wanted, "geometry"]))
            except pl.exceptions.ColumnNotFoundError as e:
                raise KeyError("Column not found in polars DataFrame.") from e
        geometries = data.select(pl.col("geometry")).to_series()
    elif kind == "geoseries":
        geometries = data.geometry
        burn = burn if burn is not None else data.index.to_numpy()
        try:
            epsg = data.crs.to_epsg()
        except AttributeError:
            pass
    else:
        geometries = data

    if isinstance(burn, np.ndarray) and burn.dtype != dtype:
        burn = np.ascontiguousarray(burn, dtype=dtype)

    raw = {"shape": shape, "extent": bounds, "resolution": resolution, "tap": tap, "epsg": epsg}
    return _rusterize(geometries, raw, fun, df, field, by, burn, background, all_touched, encoding, dtype)
