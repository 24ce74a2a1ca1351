"""Typed Python layer directly above the C ABI (include/rz_b200.h): geometry handles, grid math,
band grouping and the dense / sparse burn calls.  Mirrors the reference's core crate surface
(`Rasterize::rasterize::<DenseArray<N> | SparseArray<N>>(RasterizeContext<N>)`,
rust/src/rasterize.rs:54-62; `RasterInfoBuilder`, rust/src/geo/raster.rs:36-181)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import DTYPES, FUNS, Context, GeomSoA, RasterInfo, RawRasterInfo, Stats, errbuf, lib, ptr, raise_for


class Geoms:
    """A parsed + flattened geometry set (the `&[geo::Geometry<f64>]` argument of the reference)."""

    def __init__(self, handle: int):
        if not handle:
            raise RuntimeError("null rz_geoms handle")
        self._h = handle

    # -- constructors --------------------------------------------------------------------------
    @classmethod
    def from_wkb(cls, wkbs) -> "Geoms":
        L = lib()
        n = len(wkbs)
        bufs = [bytes(b) for b in wkbs]
        arr = (C.c_char_p * n)(*bufs)
        lens = (C.c_uint64 * n)(*[len(b) for b in bufs])
        err = errbuf()
        h = L.rz_geoms_from_wkb(arr, lens, n, err, len(err))
        if not h:
            msg = err.value.decode()
            # parse failures are panics in the reference; an all-dropped input is a ValueError
            raise (ValueError if msg.startswith("Could not parse") else RuntimeError)(msg)
        return cls(h)

    @classmethod
    def from_wkt(cls, wkts) -> "Geoms":
        L = lib()
        n = len(wkts)
        arr = (C.c_char_p * n)(*[s.encode() for s in wkts])
        err = errbuf()
        h = L.rz_geoms_from_wkt(arr, n, err, len(err))
        if not h:
            msg = err.value.decode()
            raise (ValueError if msg.startswith("Could not parse") else RuntimeError)(msg)
        return cls(h)

    @classmethod
    def from_any(cls, geoms) -> "Geoms":
        """list / ndarray of WKT str or WKB bytes (python/src/geo/parse_geometry.rs:46-63)."""
        if len(geoms) == 0:
            raise ValueError("No geometries found.")
        first = geoms[0]
        if isinstance(first, (bytes, bytearray, memoryview, np.bytes_)):
            return cls.from_wkb(geoms)
        if isinstance(first, str):
            return cls.from_wkt([str(s) for s in geoms])
        raise ValueError("Sequence must contain geometries as shapely Geometry, bytes (WKB), or string (WKT).")

    @classmethod
    def from_soa(cls, geom_part_off, part_kind, part_seq_off, seq_coord_off, x, y, device=None) -> "Geoms":
        a = [np.ascontiguousarray(geom_part_off, np.uint64), np.ascontiguousarray(part_kind, np.uint8),
             np.ascontiguousarray(part_seq_off, np.uint64), np.ascontiguousarray(seq_coord_off, np.uint64),
             np.ascontiguousarray(x, np.float64), np.ascontiguousarray(y, np.float64)]
        soa = GeomSoA(len(a[0]) - 1, len(a[1]), len(a[3]) - 1, len(a[4]), *[v.ctypes.data for v in a])
        err = errbuf()
        if device is None:
            h = lib().rz_geoms_from_soa(C.byref(soa), err, len(err))
        else:  # flatten and upload at the same time: the set is resident on `device` when this returns
            h = lib().rz_geoms_from_soa_to(C.byref(soa), int(device), err, len(err))
        if not h:
            raise RuntimeError(err.value.decode())
        return cls(h)

    @classmethod
    def from_soa_rows(cls, soa, ri, row_begin: int, row_end: int, all_touched: bool = False) -> "Geoms":
        """The row-band shard of `from_soa(*soa).row_shard(...)` without flattening the whole set first."""
        a = [np.ascontiguousarray(soa[0], np.uint64), np.ascontiguousarray(soa[1], np.uint8),
             np.ascontiguousarray(soa[2], np.uint64), np.ascontiguousarray(soa[3], np.uint64),
             np.ascontiguousarray(soa[4], np.float64), np.ascontiguousarray(soa[5], np.float64)]
        s = GeomSoA(len(a[0]) - 1, len(a[1]), len(a[3]) - 1, len(a[4]), *[v.ctypes.data for v in a])
        err = errbuf()
        h = lib().rz_geoms_from_soa_rows(C.byref(s), C.byref(ri), int(row_begin), int(row_end), int(bool(all_touched)), err, len(err))
        if not h:
            raise ValueError(err.value.decode())
        return cls(h)

    @classmethod
    def from_polygons(cls, x, y, ring_off) -> "Geoms":
        """G single-ring polygons: polygon i = coords[ring_off[i]:ring_off[i+1]] (closed if needed)."""
        ring_off = np.ascontiguousarray(ring_off, np.uint64)
        g = len(ring_off) - 1
        idx = np.arange(g + 1, dtype=np.uint64)
        return cls.from_soa(idx, np.zeros(g, np.uint8), idx, ring_off, x, y)

    # -- accessors -----------------------------------------------------------------------------
    def __len__(self) -> int:
        return lib().rz_geoms_len(self._h)

    @property
    def n_parts(self) -> int:
        return lib().rz_geoms_n_parts(self._h)

    @property
    def n_coords(self) -> int:
        return lib().rz_geoms_n_coords(self._h)

    def bounds(self):
        b = (C.c_double * 4)()
        if lib().rz_geoms_bounds(self._h, b) != 0:
            return None
        return tuple(b)

    def upload(self, device: int = 0) -> None:
        err = errbuf()
        raise_for(lib().rz_geoms_upload(self._h, device, err, len(err)), err)

    def evict(self) -> None:
        lib().rz_geoms_evict(self._h)

    def row_shard(self, ri, row_begin: int, row_end: int, all_touched: bool = False) -> "Geoms":
        """The parts that can write raster rows [row_begin, row_end) of grid `ri`, as their own geometry set (same
        order and geometry indices): what one GPU of a row-band sharded job is given."""
        err = errbuf()
        h = lib().rz_geoms_row_shard(self._h, C.byref(ri), int(row_begin), int(row_end), int(bool(all_touched)), err, len(err))
        if not h:
            raise ValueError(err.value.decode())
        return Geoms(h)

    def parts(self):
        """(part_kind[u8], part_geom[u64]) of the flattened form."""
        L, n = lib(), self.n_parts
        if n == 0:
            return np.empty(0, np.uint8), np.empty(0, np.uint64)
        kind = np.ctypeslib.as_array(C.cast(L.rz_geoms_part_kind(self._h), C.POINTER(C.c_uint8)), (n,)).copy()
        geom = np.ctypeslib.as_array(C.cast(L.rz_geoms_part_geom(self._h), C.POINTER(C.c_uint64)), (n,)).copy()
        return kind, geom

    def pool(self, kind: int):
        """(x, y, tag) of one vertex pool (0 polygon rings, 1 line strings, 2 points)."""
        L = lib()
        n = L.rz_geoms_pool_len(self._h, kind)
        if n == 0:
            return np.empty(0), np.empty(0), np.empty(0, np.uint32)
        x = np.ctypeslib.as_array(C.cast(L.rz_geoms_pool_x(self._h, kind), C.POINTER(C.c_double)), (n,)).copy()
        y = np.ctypeslib.as_array(C.cast(L.rz_geoms_pool_y(self._h, kind), C.POINTER(C.c_double)), (n,)).copy()
        t = np.ctypeslib.as_array(C.cast(L.rz_geoms_pool_tag(self._h, kind), C.POINTER(C.c_uint32)), (n,)).copy()
        return x, y, t

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None and _lib._lib is not None:  # (module globals are gone at interpreter exit)
            _lib._lib.rz_geoms_free(h)


def raster_info(geoms: Geoms | None, shape=None, extent=None, resolution=None, tap=False, epsg=None) -> RasterInfo:
    """RasterInfoBuilder::{build, build_with} (rust/src/geo/raster.rs:50-156)."""
    raw = RawRasterInfo()
    raw.has_shape = int(shape is not None)
    raw.has_extent = int(extent is not None)
    raw.has_resolution = int(resolution is not None)
    raw.tap = int(bool(tap))
    if shape is not None:
        raw.nrows, raw.ncols = int(shape[0]), int(shape[1])
    if extent is not None:
        raw.extent = (C.c_double * 4)(*[float(v) for v in extent])
    if resolution is not None:
        raw.xres, raw.yres = float(resolution[0]), float(resolution[1])
    raw.epsg = -1 if epsg is None else int(epsg)
    out = RasterInfo()
    err = errbuf()
    rc = lib().rz_raster_info_build(C.byref(raw), geoms._h if geoms is not None else None, C.byref(out), err, len(err))
    raise_for(rc, err)
    return out


def group_keys(keys):
    """group_keys (rust/src/rasterize.rs:199-205) -> (band_of_geom int32[n], sorted band names)."""
    n = len(keys)
    enc = [str(k).encode() for k in keys]
    arr = (C.c_char_p * n)(*enc)
    band = np.empty(n, np.int32)
    first = np.empty(max(n, 1), np.uint64)
    nb = lib().rz_group_keys(arr, n, band.ctypes.data, first.ctypes.data)
    return band, [str(keys[int(i)]) for i in first[:nb]]


def _context(geoms, ri, fun, dtype, field, field_valid, band_of_geom, n_bands, background, all_touched, device, rows,
             stream, flags, tile_bytes):
    dt = np.dtype(dtype)
    if dt.name not in DTYPES:
        raise ValueError("Unsupported dtype")
    if fun not in FUNS:
        raise ValueError("Unknown pixel function")  # python/src/rusterize.rs:149-151
    with np.errstate(invalid="ignore", over="ignore"):
        if np.ndim(field) == 0:
            f = np.array([field]).astype(dt)
            scalar, flen = 1, 0
        else:
            f = np.ascontiguousarray(np.asarray(field).astype(dt, copy=False))
            scalar, flen = 0, len(f)
        bg = np.array([background]).astype(dt)
    fv = None if field_valid is None else np.ascontiguousarray(field_valid, np.uint8)
    band = None if band_of_geom is None else np.ascontiguousarray(band_of_geom, np.int32)
    ctx = Context()
    ctx.raster_info = ri
    ctx.dtype = DTYPES.index(dt.name)
    ctx.pixel_fn = FUNS.index(fun)
    ctx.field = f.ctypes.data
    ctx.field_is_scalar = scalar
    ctx.all_touched = int(bool(all_touched))
    ctx.field_len = flen
    ctx.field_valid = ptr(fv)
    ctx.band_of_geom = ptr(band)
    ctx.by_len = 0 if band is None else len(band)
    ctx.n_bands = int(n_bands)
    ctx.device = int(device)
    ctx.background = bg.ctypes.data
    ctx.row_begin, ctx.row_end = (0, 0) if rows is None else (int(rows[0]), int(rows[1]))
    ctx.stream = stream
    ctx.flags = int(flags)
    ctx.tile_bytes = int(tile_bytes)
    return ctx, dt, (f, bg, fv, band)


def default_devices():
    """Devices a call uses when none are named: every visible CUDA device, or the ordinals listed in RZ_DEVICES
    (comma separated; e.g. RZ_DEVICES=0 keeps calls on one GPU)."""
    import os

    env = os.environ.get("RZ_DEVICES")
    if env:
        return [int(v) for v in env.split(",") if v.strip() != ""]
    return list(range(max(1, lib().rz_device_count())))


class _HostBlock:
    """Owner of one rz_host_alloc block: handed back to the library's pool when the last array over it goes away."""

    def __init__(self, p):
        self.p = p

    def __del__(self):
        try:
            lib().rz_host_free(self.p)
        except Exception:  # interpreter shutdown
            pass


def host_empty(shape, dtype):
    """np.empty in the library's page-locked host memory (rz_host_alloc): huge pages already faulted in, recycled
    between calls, registered with CUDA (device -> host copies land at the PCIe rate without staging) and interleaved
    over the machine's memory nodes so that GPUs on either socket write it equally fast."""
    dt = np.dtype(dtype)
    shape = tuple(int(v) for v in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    n = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
    if n == 0:
        return np.empty(shape, dt)
    err = errbuf()
    p = lib().rz_host_alloc(n, err, len(err))
    if not p:
        raise RuntimeError(err.value.decode(errors="replace"))
    buf = (C.c_uint8 * n).from_address(p)
    buf._rz_owner = _HostBlock(p)  # the array's base chain (memoryview -> buf) keeps the block alive
    return np.frombuffer(buf, dtype=dt).reshape(shape)


def host_trim(keep_bytes=0):
    """Return the library's free page-locked blocks to the system until at most keep_bytes stay pooled."""
    return int(lib().rz_host_trim(int(keep_bytes)))


_HOST_EMPTY_MIN = 64 << 20  # outputs from this size on are allocated page-locked (smaller ones: plain numpy)


def _new_output(shape, dt):
    if int(np.prod(shape, dtype=np.int64)) * np.dtype(dt).itemsize >= _HOST_EMPTY_MIN:
        return host_empty(shape, dt)
    return np.empty(shape, dt)


def _device_array(devices):
    devs = [int(d) for d in devices]
    return (C.c_int32 * len(devs))(*devs), len(devs)


def _use_device_inputs(ctx, inputs_dev, n_geoms):
    """RZ_FLAG_INPUTS_ON_DEVICE: `inputs_dev` = dict(field=<device pointer to n_geoms values of the dtype, or to one
    value with scalar=True>, valid=<device pointer or None>, band=<device pointer or None>)."""
    ctx.field = int(inputs_dev["field"])
    ctx.field_is_scalar = int(bool(inputs_dev.get("scalar", False)))
    ctx.field_len = 0 if ctx.field_is_scalar else n_geoms
    ctx.field_valid = inputs_dev.get("valid") or None
    band = inputs_dev.get("band") or None
    ctx.band_of_geom = band
    ctx.by_len = n_geoms if band else 0
    ctx.flags |= _lib.FLAG_INPUTS_ON_DEVICE


def rasterize_dense(geoms: Geoms, ri: RasterInfo, fun="last", dtype="float64", field=1, field_valid=None,
                    band_of_geom=None, n_bands=1, background=0, all_touched=False, out=None, device=0, rows=None,
                    stream=None, flags=0, tile_bytes=0, devices=None, inputs_dev=None):
    """DenseArray::build (rust/src/rasterize.rs:71-116) on the GPU.

    `out`: None (a new numpy array is returned), a C-contiguous numpy array to fill, or an int
    device pointer (then RZ_FLAG_OUT_ON_DEVICE is implied and nothing is copied back).
    `devices`: a list of CUDA ordinals -> one call over several GPUs (row bands, rz_rasterize_dense_multi); the
    stats dict then aggregates and carries the per-device stats under "per_device".
    Returns (array_or_None, stats dict).  Shape [n_bands, rows, ncols]."""
    ctx, dt, keep = _context(geoms, ri, fun, dtype, field, field_valid, band_of_geom, n_bands, background,
                             all_touched, device, rows, stream, flags, tile_bytes)
    nb = n_bands if (band_of_geom is not None or (inputs_dev and inputs_dev.get("band"))) else 1
    if inputs_dev:
        _use_device_inputs(ctx, inputs_dev, len(geoms))
    nrows = ri.nrows if rows is None else rows[1] - rows[0]
    arr = None
    if isinstance(out, (int, np.integer)):
        ctx.flags |= _lib.FLAG_OUT_ON_DEVICE
        out_ptr = int(out)
    else:
        # RZ_FLAG_OUT_ROW_COL_BAND: C-order [band][col][row] == R's (row, col, band) column-major array
        shape = (nb, ri.ncols, nrows) if (int(flags) & _lib.FLAG_OUT_ROW_COL_BAND) else (nb, nrows, ri.ncols)
        arr = _new_output(shape, dt) if out is None else out
        if arr.dtype != dt or not arr.flags.c_contiguous or arr.size != nb * nrows * ri.ncols:
            raise ValueError("`out` must be a C-contiguous array of the output dtype and shape")
        out_ptr = arr.ctypes.data
    st = Stats()
    err = errbuf()
    if devices is not None and len(devices) > 0:
        darr, nd = _device_array(devices)
        per = (Stats * nd)()
        rc = lib().rz_rasterize_dense_multi(geoms._h, C.byref(ctx), darr, nd, out_ptr, C.byref(st), per, err, len(err))
        raise_for(rc, err)
        d = st.as_dict()
        d["per_device"] = [p.as_dict() for p in per]
        return arr, d
    rc = lib().rz_rasterize_dense(geoms._h, C.byref(ctx), out_ptr, C.byref(st), err, len(err))
    raise_for(rc, err)
    return arr, st.as_dict()


def rasterize_dense_soa(soa, ri: RasterInfo, fun="last", dtype="float64", field=1, field_valid=None, band_of_geom=None,
                        n_bands=1, background=0, all_touched=False, out=None, devices=None, rows=None, flags=0):
    """DenseArray::build in ONE library call (rz_rasterize_dense_soa): the six rz_geom_soa arrays, the context, the
    devices, the host array.  Flattening, upload, burn and copy-back all happen inside; with several devices each one
    flattens only its row band's parts straight out of `soa`.  Returns (array, stats dict)."""
    a = [np.ascontiguousarray(soa[0], np.uint64), np.ascontiguousarray(soa[1], np.uint8),
         np.ascontiguousarray(soa[2], np.uint64), np.ascontiguousarray(soa[3], np.uint64),
         np.ascontiguousarray(soa[4], np.float64), np.ascontiguousarray(soa[5], np.float64)]
    s = GeomSoA(len(a[0]) - 1, len(a[1]), len(a[3]) - 1, len(a[4]), *[v.ctypes.data for v in a])
    ctx, dt, keep = _context(None, ri, fun, dtype, field, field_valid, band_of_geom, n_bands, background, all_touched,
                             0, rows, None, flags, 0)
    nb = n_bands if band_of_geom is not None else 1
    nrows = ri.nrows if rows is None else rows[1] - rows[0]
    shape = (nb, ri.ncols, nrows) if (int(flags) & _lib.FLAG_OUT_ROW_COL_BAND) else (nb, nrows, ri.ncols)
    arr = _new_output(shape, dt) if out is None else out
    if arr.dtype != dt or not arr.flags.c_contiguous or arr.size != nb * nrows * ri.ncols:
        raise ValueError("`out` must be a C-contiguous array of the output dtype and shape")
    darr, nd = _device_array(default_devices() if devices is None else devices)
    per = (Stats * nd)()
    st = Stats()
    err = errbuf()
    rc = lib().rz_rasterize_dense_soa(C.byref(s), C.byref(ctx), darr, nd, arr.ctypes.data, C.byref(st), per, err, len(err))
    raise_for(rc, err)
    d = st.as_dict()
    d["per_device"] = [p.as_dict() for p in per]
    return arr, d


def rasterize_sparse(geoms: Geoms, ri: RasterInfo, fun="last", dtype="float64", field=1, field_valid=None,
                     band_of_geom=None, n_bands=1, background=0, all_touched=False, device=0, stream=None, flags=0,
                     devices=None):
    """SparseArray::build (rust/src/rasterize.rs:118-157) on the GPU -> dict(rows, cols, data, counts, stats).
    `devices`: a list of CUDA ordinals -> contiguous geometry ranges per GPU, streams concatenated by offset
    (rz_rasterize_sparse_multi)."""
    ctx, dt, keep = _context(geoms, ri, fun, dtype, field, field_valid, band_of_geom, n_bands, background,
                             all_touched, device, None, stream, flags, 0)
    L = lib()
    h = C.c_void_p()
    st = Stats()
    err = errbuf()
    per = None
    if devices is not None and len(devices) > 0:
        darr, nd = _device_array(devices)
        per = (Stats * nd)()
        rc = L.rz_rasterize_sparse_multi(geoms._h, C.byref(ctx), darr, nd, C.byref(h), C.byref(st), per, err, len(err))
    else:
        rc = L.rz_rasterize_sparse(geoms._h, C.byref(ctx), C.byref(h), C.byref(st), err, len(err))
    raise_for(rc, err)
    owner = _SparseOwner(h)  # the arrays below are views of the library's buffers: freed with the last of them
    n, nb = L.rz_sparse_len(h), L.rz_sparse_n_bands(h)

    def view(p, count, t):
        if count == 0:
            return np.empty(0, t)
        buf = (C.c_char * (count * np.dtype(t).itemsize)).from_address(p)
        buf._owner = owner
        return np.frombuffer(buf, dtype=t)

    stats = st.as_dict()
    if per is not None:
        stats["per_device"] = [p.as_dict() for p in per]
    return dict(rows=view(L.rz_sparse_rows(h), n, np.uint64), cols=view(L.rz_sparse_cols(h), n, np.uint64),
                data=view(L.rz_sparse_data(h), n, dt),
                counts=view(L.rz_sparse_counts(h), nb, np.uint64).copy(), stats=stats)


class _SparseOwner:
    """Keeps an rz_sparse handle alive for the numpy views of its buffers (zero-copy hand-over, like the
    reference's into_pyarray, python/src/encoding/pyarray.rs:27-28)."""

    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None and _lib._lib is not None:
            _lib._lib.rz_sparse_free(h)


def sparse_build_array(ri: RasterInfo, fun, background, counts, rows, cols, data, device=0):
    """SparseArray::build_array (rust/src/encoding/arrays.rs:103-143): replay the triplets through the
    pixel function on the GPU -> array [n_bands, nrows, ncols]."""
    dt = data.dtype
    ctx = Context()
    ctx.raster_info = ri
    ctx.dtype = DTYPES.index(dt.name)
    ctx.pixel_fn = FUNS.index(fun)
    with np.errstate(invalid="ignore", over="ignore"):
        bg = np.array([background]).astype(dt)
    ctx.background = bg.ctypes.data
    ctx.device = int(device)
    counts = np.ascontiguousarray(counts, np.uint64)
    rows = np.ascontiguousarray(rows, np.uint64)
    cols = np.ascontiguousarray(cols, np.uint64)
    data = np.ascontiguousarray(data)
    out = np.empty((len(counts), ri.nrows, ri.ncols), dt)
    st = Stats()
    err = errbuf()
    rc = lib().rz_sparse_build_array(C.byref(ctx), len(counts), counts.ctypes.data, rows.ctypes.data, cols.ctypes.data,
                                     data.ctypes.data, out.ctypes.data, C.byref(st), err, len(err))
    raise_for(rc, err)
    return out
