"""WKT -> ISO WKB (2-D, little endian) in pure Python.  TEST INFRASTRUCTURE ONLY.

The oracle consumes WKB; the reference's tests are written as WKT strings
(/root/reference/python/test/test_many.py:19-25), so tests convert with this helper.  It is
deliberately independent of the product's C++ WKT reader (rusterize_b200/csrc/rz_wkt.cpp) so the two
can be checked against each other.
"""
from __future__ import annotations

import re
import struct

_TYPES = {
    "POINT": 1,
    "LINESTRING": 2,
    "POLYGON": 3,
    "MULTIPOINT": 4,
    "MULTILINESTRING": 5,
    "MULTIPOLYGON": 6,
    "GEOMETRYCOLLECTION": 7,
}
_TOKEN = re.compile(r"\s*([A-Za-z]+|\(|\)|,|[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|nan|inf))", re.I)


class _Lexer:
    def __init__(self, s: str):
        self.toks = _TOKEN.findall(s)
        if "".join(self.toks).replace(" ", "") != re.sub(r"\s+", "", s):
            raise ValueError(f"cannot tokenise WKT: {s[:60]!r}")
        self.i = 0

    def peek(self):
        return self.toks[self.i] if self.i < len(self.toks) else None

    def next(self):
        t = self.peek()
        self.i += 1
        return t

    def expect(self, t):
        got = self.next()
        if got != t:
            raise ValueError(f"WKT: expected {t!r}, got {got!r}")


def _hdr(t: int) -> bytes:
    return struct.pack("<BI", 1, t)


def _coord(lx: _Lexer, ndim: int):
    vals = [float(lx.next()) for _ in range(ndim)]
    return vals[0], vals[1]


def _coords(lx: _Lexer, ndim: int):
    lx.expect("(")
    out = [_coord(lx, ndim)]
    while lx.peek() == ",":
        lx.next()
        out.append(_coord(lx, ndim))
    lx.expect(")")
    return out


def _line_body(pts) -> bytes:
    return struct.pack("<I", len(pts)) + b"".join(struct.pack("<dd", *p) for p in pts)


def _geom(lx: _Lexer) -> bytes:
    name = lx.next().upper()
    t = _TYPES[name]
    ndim = 2
    while lx.peek() is not None and lx.peek().upper() in ("Z", "M", "ZM"):
        ndim = 2 + len(lx.next())
    if lx.peek() is not None and lx.peek().upper() == "EMPTY":
        lx.next()
        if t == 1:
            return _hdr(1) + struct.pack("<dd", float("nan"), float("nan"))
        return _hdr(t) + struct.pack("<I", 0)
    if t == 1:
        (p,) = _coords(lx, ndim)
        return _hdr(1) + struct.pack("<dd", *p)
    if t == 2:
        return _hdr(2) + _line_body(_coords(lx, ndim))
    if t == 3:
        return _hdr(3) + _poly_body(lx, ndim)
    lx.expect("(")
    parts = []
    while True:
        if t == 4:  # MULTIPOINT ((1 2), (3 4)) or MULTIPOINT (1 2, 3 4)
            if lx.peek() == "(":
                (p,) = _coords(lx, ndim)
            else:
                p = _coord(lx, ndim)
            parts.append(_hdr(1) + struct.pack("<dd", *p))
        elif t == 5:
            parts.append(_hdr(2) + _line_body(_coords(lx, ndim)))
        elif t == 6:
            parts.append(_hdr(3) + _poly_body(lx, ndim))
        else:
            parts.append(_geom(lx))
        if lx.peek() == ",":
            lx.next()
            continue
        break
    lx.expect(")")
    return _hdr(t) + struct.pack("<I", len(parts)) + b"".join(parts)


def _poly_body(lx: _Lexer, ndim: int) -> bytes:
    lx.expect("(")
    rings = [_coords(lx, ndim)]
    while lx.peek() == ",":
        lx.next()
        rings.append(_coords(lx, ndim))
    lx.expect(")")
    return struct.pack("<I", len(rings)) + b"".join(_line_body(r) for r in rings)


def wkt_to_wkb(s: str) -> bytes:
    lx = _Lexer(s)
    out = _geom(lx)
    if lx.peek() is not None:
        raise ValueError("trailing tokens in WKT")
    return out


# ------------------------------------------------------------------------------------------------
# Builders from numeric data (used by the synthetic workloads)
# ------------------------------------------------------------------------------------------------
def polygon_wkb(rings) -> bytes:
    """rings: list of (n,2) float arrays / lists."""
    return _hdr(3) + struct.pack("<I", len(rings)) + b"".join(_line_body([tuple(map(float, p)) for p in r]) for r in rings)


def linestring_wkb(pts) -> bytes:
    return _hdr(2) + _line_body([tuple(map(float, p)) for p in pts])


def point_wkb(x, y) -> bytes:
    return _hdr(1) + struct.pack("<dd", float(x), float(y))


def multipoint_wkb(pts) -> bytes:
    return _hdr(4) + struct.pack("<I", len(pts)) + b"".join(point_wkb(*p) for p in pts)


def multilinestring_wkb(lines) -> bytes:
    return _hdr(5) + struct.pack("<I", len(lines)) + b"".join(linestring_wkb(l) for l in lines)


def multipolygon_wkb(polys) -> bytes:
    return _hdr(6) + struct.pack("<I", len(polys)) + b"".join(polygon_wkb(p) for p in polys)


def collection_wkb(members) -> bytes:
    return _hdr(7) + struct.pack("<I", len(members)) + b"".join(members)
