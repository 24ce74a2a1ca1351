"""ctypes front-end of the CPU oracle (oracle/rz_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this package; the product (rusterize_b200) never does.  The oracle restates the reference's
CPU algorithm (see the header of rz_oracle.cpp for the file:line map) and is pinned against the
reference's golden vectors by tests/test_oracle_golden.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from .wkt2wkb import wkt_to_wkb

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "librz_oracle.so"

DTYPES = ["uint8", "uint16", "uint32", "uint64", "int8", "int16", "int32", "int64", "float32", "float64"]
FUNS = ["sum", "first", "last", "min", "max", "count", "any"]


def build(force: bool = False) -> Path:
    src = _HERE / "rz_oracle.cpp"
    if force or not _SO.exists() or (src.exists() and _SO.stat().st_mtime < src.stat().st_mtime):
        subprocess.run(["make", "-C", str(_HERE), "-B", "_build/librz_oracle.so"], check=True, capture_output=True)
    return _SO


class RasterInfo(C.Structure):
    _fields_ = [
        ("nrows", C.c_uint64), ("ncols", C.c_uint64),
        ("xmin", C.c_double), ("ymin", C.c_double), ("xmax", C.c_double), ("ymax", C.c_double),
        ("xres", C.c_double), ("yres", C.c_double),
    ]

    def as_tuple(self):
        return (self.nrows, self.ncols, self.xmin, self.ymin, self.xmax, self.ymax, self.xres, self.yres)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_SO))
        L.rzo_geoms_from_wkb.restype = C.c_void_p
        L.rzo_geoms_from_wkb.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_uint64), C.c_uint64, C.c_char_p, C.c_uint64]
        L.rzo_geoms_from_rings.restype = C.c_void_p
        L.rzo_geoms_from_rings.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.rzo_geoms_len.restype = C.c_uint64
        L.rzo_geoms_len.argtypes = [C.c_void_p]
        L.rzo_geoms_free.argtypes = [C.c_void_p]
        L.rzo_geoms_bounds.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.rzo_raster_info_build.argtypes = [
            C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_double), C.c_int, C.c_double, C.c_double,
            C.c_int, C.c_void_p, C.POINTER(RasterInfo), C.c_char_p, C.c_uint64]
        L.rzo_group_keys.restype = C.c_int64
        L.rzo_group_keys.argtypes = [C.POINTER(C.c_char_p), C.c_uint64, C.c_void_p, C.c_void_p]
        common = [C.c_void_p, C.POINTER(RasterInfo), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint64,
                  C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_int]
        L.rzo_rasterize_dense.argtypes = common + [C.c_void_p, C.c_char_p, C.c_uint64]
        L.rzo_rasterize_sparse.argtypes = common + [C.POINTER(C.c_void_p), C.c_char_p, C.c_uint64]
        for name in ("rzo_sparse_len", "rzo_sparse_bands"):
            getattr(L, name).restype = C.c_uint64
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ("rzo_sparse_rows", "rzo_sparse_cols", "rzo_sparse_counts", "rzo_sparse_data"):
            getattr(L, name).restype = C.c_void_p
            getattr(L, name).argtypes = [C.c_void_p]
        L.rzo_sparse_free.argtypes = [C.c_void_p]
        L.rzo_sparse_replay.argtypes = [C.POINTER(RasterInfo), C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code  # 1 ValueError, 2 RuntimeError


class Geoms:
    """Decoded geometry list (geo_types::Geometry equivalents)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def from_wkb(cls, wkbs):
        L = lib()
        n = len(wkbs)
        arr = (C.c_char_p * n)(*[bytes(b) for b in wkbs])
        lens = (C.c_uint64 * n)(*[len(b) for b in wkbs])
        err = C.create_string_buffer(256)
        h = L.rzo_geoms_from_wkb(arr, lens, n, err, 256)
        if not h:
            raise OracleError(2, err.value.decode())
        return cls(h)

    @classmethod
    def from_any(cls, geoms):
        """list of WKT str / WKB bytes"""
        return cls.from_wkb([wkt_to_wkb(g) if isinstance(g, str) else bytes(g) for g in geoms])

    @classmethod
    def from_rings(cls, x, y, off):
        x = np.ascontiguousarray(x, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        off = np.ascontiguousarray(off, np.uint64)
        return cls(lib().rzo_geoms_from_rings(x.ctypes.data, y.ctypes.data, off.ctypes.data, len(off) - 1))

    def __len__(self):
        return lib().rzo_geoms_len(self._h)

    def bounds(self):
        b = (C.c_double * 4)()
        if lib().rzo_geoms_bounds(self._h, b) != 0:
            return None
        return tuple(b)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().rzo_geoms_free(self._h)
            self._h = None


def raster_info(geoms: Geoms | None, shape=None, extent=None, resolution=None, tap=False) -> RasterInfo:
    ri = RasterInfo()
    err = C.create_string_buffer(256)
    ext = (C.c_double * 4)(*extent) if extent is not None else None
    rc = lib().rzo_raster_info_build(
        int(shape is not None), *(shape if shape is not None else (0, 0)),
        int(extent is not None), ext,
        int(resolution is not None), *(map(float, resolution) if resolution is not None else (0.0, 0.0)),
        int(tap), geoms._h if geoms is not None else None, C.byref(ri), err, 256)
    if rc:
        raise OracleError(rc, err.value.decode())
    return ri


def group_keys(keys):
    n = len(keys)
    arr = (C.c_char_p * n)(*[k.encode() for k in keys])
    band = np.empty(n, np.int32)
    first = np.empty(max(n, 1), np.uint64)
    nb = lib().rzo_group_keys(arr, n, band.ctypes.data, first.ctypes.data)
    return band, [keys[int(i)] for i in first[:nb]]


def _prep(geoms, ri, fun, dtype, burn, field_valid, by, background):
    dt = np.dtype(dtype)
    if burn is None:
        burn = 1
    if np.ndim(burn) == 0:
        field = np.array([burn]).astype(dt)
        scalar, flen = 1, 0
    else:
        field = np.ascontiguousarray(burn).astype(dt)
        scalar, flen = 0, len(field)
    bg = np.array([background]).astype(dt) if background is not None else np.zeros(1, dt)
    fv = np.ascontiguousarray(field_valid, np.uint8) if field_valid is not None else None
    if by is not None:
        band, names = group_keys([str(k) for k in by])
        bylen = len(band)
    else:
        band, names, bylen = None, ["band_1"], 0
    args = [geoms._h, C.byref(ri), DTYPES.index(dt.name), FUNS.index(fun)]
    tail = [field.ctypes.data, scalar, flen, fv.ctypes.data if fv is not None else None,
            band.ctypes.data if band is not None else None, bylen, len(names), bg.ctypes.data]
    keep = (field, bg, fv, band)
    return dt, args, tail, names, keep


def rasterize_dense(geoms: Geoms, ri: RasterInfo, fun="last", dtype="float64", burn=None, field_valid=None, by=None,
                    background=np.nan, all_touched=False, threads=1):
    """-> (array [B,R,C], band_names).  `background` must already be representable in dtype."""
    with np.errstate(invalid="ignore"):
        dt, args, tail, names, keep = _prep(geoms, ri, fun, dtype, burn, field_valid, by, background)
    out = np.empty((len(names), ri.nrows, ri.ncols), dt)
    err = C.create_string_buffer(256)
    rc = lib().rzo_rasterize_dense(*args, int(all_touched), *tail, int(threads), out.ctypes.data, err, 256)
    if rc:
        raise OracleError(rc, err.value.decode())
    return out, names


def rasterize_sparse(geoms: Geoms, ri: RasterInfo, fun="last", dtype="float64", burn=None, field_valid=None, by=None,
                     background=np.nan, all_touched=False, threads=1):
    """-> dict(rows, cols, data, counts, band_names)"""
    with np.errstate(invalid="ignore"):
        dt, args, tail, names, keep = _prep(geoms, ri, fun, dtype, burn, field_valid, by, background)
    h = C.c_void_p()
    err = C.create_string_buffer(256)
    L = lib()
    rc = L.rzo_rasterize_sparse(*args, int(all_touched), *tail, int(threads), C.byref(h), err, 256)
    if rc:
        raise OracleError(rc, err.value.decode())
    n, nb = L.rzo_sparse_len(h), L.rzo_sparse_bands(h)

    def view(ptr, count, t):
        if count == 0:
            return np.empty(0, t)
        return np.frombuffer((C.c_char * (count * np.dtype(t).itemsize)).from_address(ptr), dtype=t).copy()

    res = dict(
        rows=view(L.rzo_sparse_rows(h), n, np.uint64), cols=view(L.rzo_sparse_cols(h), n, np.uint64),
        data=view(L.rzo_sparse_data(h), n, dt), counts=view(L.rzo_sparse_counts(h), nb, np.uint64), band_names=names)
    L.rzo_sparse_free(h)
    return res


def sparse_replay(ri: RasterInfo, sp: dict, fun, background):
    dt = sp["data"].dtype
    with np.errstate(invalid="ignore"):
        bg = np.array([background]).astype(dt) if background is not None else np.zeros(1, dt)
    nb = len(sp["counts"])
    out = np.empty((nb, ri.nrows, ri.ncols), dt)
    lib().rzo_sparse_replay(C.byref(ri), DTYPES.index(dt.name), FUNS.index(fun), bg.ctypes.data, nb,
                            sp["counts"].ctypes.data, sp["rows"].ctypes.data, sp["cols"].ctypes.data,
                            sp["data"].ctypes.data, out.ctypes.data)
    return out


def rusterize(geoms, res=None, out_shape=None, extent=None, burn=None, by=None, fun="last", background=np.nan,
              encoding="numpy", all_touched=False, tap=False, dtype="float64", field_valid=None, threads=1):
    """Oracle counterpart of rusterize(list_of_wkt_or_wkb, ..., encoding='numpy'|'sparse')."""
    g = geoms if isinstance(geoms, Geoms) else Geoms.from_any(geoms)
    ri = raster_info(g, shape=out_shape, extent=extent, resolution=res, tap=tap)
    dt = np.dtype(dtype)
    bg = background
    if dt.kind in "iu":  # python/src/rusterize.rs:50-53 — a value that does not extract becomes 0
        ok = isinstance(bg, (int, np.integer)) and not isinstance(bg, bool) and np.iinfo(dt).min <= bg <= np.iinfo(dt).max
        bg = bg if ok else 0
    elif bg is None:
        bg = 0.0
    if encoding == "sparse":
        sp = rasterize_sparse(g, ri, fun, dtype, burn, field_valid, by, bg, all_touched, threads)
        sp["raster_info"] = ri
        sp["background"] = bg
        return sp
    arr, _ = rasterize_dense(g, ri, fun, dtype, burn, field_valid, by, bg, all_touched, threads)
    return arr
