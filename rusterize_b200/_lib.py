"""ctypes binding of librz_b200.so (include/rz_b200.h).  Fails loudly when the library is missing:
there is no CPU fallback for the burn path."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
SO_PATH = Path(__import__("os").environ.get("RZ_B200_SO", _HERE / "librz_b200.so"))

DTYPES = ["uint8", "uint16", "uint32", "uint64", "int8", "int16", "int32", "int64", "float32", "float64"]
FUNS = ["sum", "first", "last", "min", "max", "count", "any"]

RZ_OK, RZ_VALUE_ERROR, RZ_RUNTIME_ERROR = 0, 1, 2
FLAG_OUT_ON_DEVICE, FLAG_FORCE_H2D, FLAG_SYNC_STAGES = 1, 2, 4
FLAG_NO_TILE_ENGINE, FLAG_FORCE_TILE_ENGINE, FLAG_STREAMED_H2D, FLAG_INPUTS_ON_DEVICE = 8, 16, 64, 128
FLAG_OUT_ROW_COL_BAND = 256


class RasterInfo(C.Structure):
    _fields_ = [
        ("nrows", C.c_uint64), ("ncols", C.c_uint64),
        ("xmin", C.c_double), ("ymin", C.c_double), ("xmax", C.c_double), ("ymax", C.c_double),
        ("xres", C.c_double), ("yres", C.c_double),
        ("epsg", C.c_int32), ("_pad", C.c_int32),
    ]


class RawRasterInfo(C.Structure):
    _fields_ = [
        ("has_shape", C.c_int32), ("has_extent", C.c_int32), ("has_resolution", C.c_int32), ("tap", C.c_int32),
        ("nrows", C.c_uint64), ("ncols", C.c_uint64),
        ("extent", C.c_double * 4),
        ("xres", C.c_double), ("yres", C.c_double),
        ("epsg", C.c_int32), ("_pad", C.c_int32),
    ]


class GeomSoA(C.Structure):
    _fields_ = [
        ("n_geoms", C.c_uint64), ("n_parts", C.c_uint64), ("n_seqs", C.c_uint64), ("n_coords", C.c_uint64),
        ("geom_part_off", C.c_void_p), ("part_kind", C.c_void_p), ("part_seq_off", C.c_void_p),
        ("seq_coord_off", C.c_void_p), ("x", C.c_void_p), ("y", C.c_void_p),
    ]


class Context(C.Structure):
    _fields_ = [
        ("raster_info", RasterInfo),
        ("dtype", C.c_int32), ("pixel_fn", C.c_int32),
        ("field", C.c_void_p),
        ("field_is_scalar", C.c_int32), ("all_touched", C.c_int32),
        ("field_len", C.c_uint64),
        ("field_valid", C.c_void_p),
        ("band_of_geom", C.c_void_p),
        ("by_len", C.c_uint64),
        ("n_bands", C.c_int32), ("device", C.c_int32),
        ("background", C.c_void_p),
        ("row_begin", C.c_uint64), ("row_end", C.c_uint64),
        ("stream", C.c_void_p),
        ("flags", C.c_uint32), ("tile_bytes", C.c_uint32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("n_parts", C.c_uint64), ("n_poly_vertices", C.c_uint64), ("n_line_vertices", C.c_uint64),
        ("n_points", C.c_uint64), ("n_records", C.c_uint64), ("n_crossings", C.c_uint64), ("n_tasks", C.c_uint64),
        ("key_bits", C.c_uint32), ("sort_passes", C.c_uint32), ("tile_width", C.c_uint32), ("n_windows", C.c_uint32),
        ("h2d_ms", C.c_float), ("count_ms", C.c_float), ("emit_ms", C.c_float), ("sort_ms", C.c_float),
        ("index_ms", C.c_float), ("fill_ms", C.c_float), ("d2h_ms", C.c_float), ("total_ms", C.c_float),
        ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("out_bytes", C.c_uint64),
        ("kernel_launches", C.c_uint32), ("engine", C.c_uint32),
        ("n_mask_words", C.c_uint64), ("host_syncs", C.c_uint32), ("plan_cached", C.c_uint32),
        ("wall_ms", C.c_float), ("shard_ms", C.c_float),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("_")}


# every symbol include/rz_b200.h declares: name -> (restype, argtypes)
_ERR = [C.c_char_p, C.c_size_t]
SYMBOLS = {
    "rz_geoms_from_wkb": (C.c_void_p, [C.POINTER(C.c_char_p), C.POINTER(C.c_uint64), C.c_uint64] + _ERR),
    "rz_geoms_from_wkt": (C.c_void_p, [C.POINTER(C.c_char_p), C.c_uint64] + _ERR),
    "rz_geoms_from_soa": (C.c_void_p, [C.POINTER(GeomSoA)] + _ERR),
    "rz_geoms_from_soa_to": (C.c_void_p, [C.POINTER(GeomSoA), C.c_int] + _ERR),
    "rz_geoms_len": (C.c_uint64, [C.c_void_p]),
    "rz_geoms_n_parts": (C.c_uint64, [C.c_void_p]),
    "rz_geoms_n_coords": (C.c_uint64, [C.c_void_p]),
    "rz_geoms_bounds": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "rz_geoms_upload": (C.c_int, [C.c_void_p, C.c_int] + _ERR),
    "rz_geoms_evict": (None, [C.c_void_p]),
    "rz_geoms_free": (None, [C.c_void_p]),
    "rz_geoms_row_shard": (C.c_void_p, [C.c_void_p, C.POINTER(RasterInfo), C.c_uint64, C.c_uint64, C.c_int] + _ERR),
    "rz_geoms_from_soa_rows": (C.c_void_p, [C.POINTER(GeomSoA), C.POINTER(RasterInfo), C.c_uint64, C.c_uint64, C.c_int] + _ERR),
    "rz_geoms_part_kind": (C.c_void_p, [C.c_void_p]),
    "rz_geoms_part_geom": (C.c_void_p, [C.c_void_p]),
    "rz_geoms_pool_len": (C.c_uint64, [C.c_void_p, C.c_int]),
    "rz_geoms_pool_x": (C.c_void_p, [C.c_void_p, C.c_int]),
    "rz_geoms_pool_y": (C.c_void_p, [C.c_void_p, C.c_int]),
    "rz_geoms_pool_tag": (C.c_void_p, [C.c_void_p, C.c_int]),
    "rz_raster_info_build": (C.c_int, [C.POINTER(RawRasterInfo), C.c_void_p, C.POINTER(RasterInfo)] + _ERR),
    "rz_group_keys": (C.c_int64, [C.POINTER(C.c_char_p), C.c_uint64, C.c_void_p, C.c_void_p]),
    "rz_rasterize_dense": (C.c_int, [C.c_void_p, C.POINTER(Context), C.c_void_p, C.POINTER(Stats)] + _ERR),
    "rz_rasterize_sparse": (C.c_int, [C.c_void_p, C.POINTER(Context), C.POINTER(C.c_void_p), C.POINTER(Stats)] + _ERR),
    "rz_rasterize_dense_multi": (C.c_int, [C.c_void_p, C.POINTER(Context), C.POINTER(C.c_int32), C.c_int32, C.c_void_p,
                                           C.POINTER(Stats), C.POINTER(Stats)] + _ERR),
    "rz_rasterize_sparse_multi": (C.c_int, [C.c_void_p, C.POINTER(Context), C.POINTER(C.c_int32), C.c_int32,
                                            C.POINTER(C.c_void_p), C.POINTER(Stats), C.POINTER(Stats)] + _ERR),
    "rz_rasterize_dense_soa": (C.c_int, [C.POINTER(GeomSoA), C.POINTER(Context), C.POINTER(C.c_int32), C.c_int32, C.c_void_p,
                                         C.POINTER(Stats), C.POINTER(Stats)] + _ERR),
    "rz_sparse_len": (C.c_uint64, [C.c_void_p]),
    "rz_sparse_n_bands": (C.c_uint64, [C.c_void_p]),
    "rz_sparse_rows": (C.c_void_p, [C.c_void_p]),
    "rz_sparse_cols": (C.c_void_p, [C.c_void_p]),
    "rz_sparse_data": (C.c_void_p, [C.c_void_p]),
    "rz_sparse_counts": (C.c_void_p, [C.c_void_p]),
    "rz_sparse_free": (None, [C.c_void_p]),
    "rz_sparse_build_array": (C.c_int, [C.POINTER(Context), C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.POINTER(Stats)] + _ERR),
    "rz_host_alloc": (C.c_void_p, [C.c_size_t, C.c_char_p, C.c_size_t]),
    "rz_host_free": (None, [C.c_void_p]),
    "rz_host_trim": (C.c_uint64, [C.c_uint64]),
    "rz_device_count": (C.c_int, []),
    "rz_version": (C.c_char_p, []),
    "rz_abi_layout": (C.c_int, [C.POINTER(C.c_uint64), C.c_int]),
}

_lib = None


def abi_layout_of_bindings():
    """The 16 numbers rz_abi_layout() reports, computed from the ctypes mirrors above."""
    return [C.sizeof(RasterInfo), C.sizeof(RawRasterInfo), C.sizeof(GeomSoA), C.sizeof(Context), C.sizeof(Stats),
            Context.field.offset, Context.band_of_geom.offset, Context.background.offset, Context.row_begin.offset,
            Context.stream.offset, Context.flags.offset, Stats.h2d_ms.offset, Stats.h2d_bytes.offset,
            Stats.kernel_launches.offset, Stats.n_mask_words.offset, Stats.wall_ms.offset]


def lib() -> C.CDLL:
    """Load librz_b200.so.  Raises if it has not been built: the burn path has no CPU fallback."""
    global _lib
    if _lib is None:
        if not SO_PATH.exists():
            raise ImportError(
                f"{SO_PATH} is missing. Build it with `python -m rusterize_b200.build` "
                "(nvcc, sm_100a). rusterize_b200 has no CPU fallback.")
        L = C.CDLL(str(SO_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        theirs = (C.c_uint64 * 16)()
        if L.rz_abi_layout(theirs, 16) != 16 or list(theirs) != abi_layout_of_bindings():
            raise ImportError(f"{SO_PATH}: struct layouts differ from rusterize_b200/_lib.py "
                              f"(library {list(theirs)}, bindings {abi_layout_of_bindings()}); rebuild the library")
        _lib = L
    return _lib


class RzError(Exception):
    pass


def raise_for(code: int, err: C.Array) -> None:
    """Map the C ABI's return codes onto the exceptions the reference's Python binding raises
    (python/src/rusterize.rs:121-123, 148-151)."""
    if code == RZ_OK:
        return
    msg = err.value.decode(errors="replace")
    if code == RZ_VALUE_ERROR:
        raise ValueError(msg)
    raise RuntimeError(msg)


def errbuf():
    return C.create_string_buffer(512)


def ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data
