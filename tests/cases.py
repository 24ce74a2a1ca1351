"""Shared test inputs: the reference's own fixtures, restated as WKT lists.

GEOMS / values: /root/reference/python/test/test_many.py:19-28.
R helpers sq()/raster_info(): /root/reference/R/rusterize/tests/testthat/helper-geoms.R.
"""
import numpy as np

GEOMS = [
    "POLYGON ((-180 -20, -140 55, 10 0, -140 -60, -180 -20), (-150 -20, -100 -10, -110 20, -150 -20))",
    "POLYGON ((-10 0, 140 60, 160 0, 140 -55, -10 0))",
    "POLYGON ((-125 0, 0 60, 40 5, 15 -45, -125 0))",
    "MULTILINESTRING ((-180 -70, -140 -50), (-140 -50, -100 -70), (-100 -70, -60 -50), (-60 -50, -20 -70), "
    "(-20 -70, 20 -50), (20 -50, 60 -70), (60 -70, 100 -50), (100 -50, 140 -70), (140 -70, 180 -50))",
    "GEOMETRYCOLLECTION (POINT (50 -40), POLYGON ((75 -40, 75 -30, 100 -30, 100 -40, 75 -40)), "
    "LINESTRING (60 -40, 80 0), GEOMETRYCOLLECTION (POLYGON ((100 20, 100 30, 110 30, 110 20, 100 20))))",
]
VALUES = np.arange(1, 6)

# GDF.explode().explode() (test_many.py:329-346): multi-parts and collection members become rows,
# each keeping its parent's value.
_MLS = [(-180, -70), (-140, -50), (-100, -70), (-60, -50), (-20, -70), (20, -50), (60, -70), (100, -50), (140, -70), (180, -50)]
GEOMS_EXPLODED = GEOMS[:3] + [
    f"LINESTRING ({a[0]} {a[1]}, {b[0]} {b[1]})" for a, b in zip(_MLS[:-1], _MLS[1:])
] + [
    "POINT (50 -40)",
    "POLYGON ((75 -40, 75 -30, 100 -30, 100 -40, 75 -40))",
    "LINESTRING (60 -40, 80 0)",
    "POLYGON ((100 20, 100 30, 110 30, 110 20, 100 20))",
]
VALUES_EXPLODED = np.array([1, 2, 3] + [4] * 9 + [5] * 4)


def sq(x0, y0, x1, y1):
    return f"POLYGON (({x0} {y0}, {x1} {y0}, {x1} {y1}, {x0} {y1}, {x0} {y0}))"


R_INFO = dict(out_shape=(4, 4), extent=(0, 0, 4, 4))


def rmat(vals):
    """R's matrix(c(...), 4, 4) is column-major."""
    return np.array(vals, dtype=np.float64).reshape(4, 4).T
