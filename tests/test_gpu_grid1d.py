"""tile_apply launches a 2-D grid (column groups x tile rows*bands) and falls back to a flattened 1-D grid when
tile rows x bands exceeds the 65535 limit of gridDim.y - a size no test raster reaches.  RZ_APPLY_1D forces the
fallback (read once per process, hence a subprocess)."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent("""
    import sys
    sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
    import numpy as np, oracle, synth
    from rusterize_b200 import core, _lib
    x, y, off = synth.star_polygons(3, 2500, 8, 40, 60.0, 1500, 700)
    n = len(off) - 1
    by = [str(i % 3) for i in range(n)]
    band, names = core.group_keys(by)
    kw = dict(shape=(700, 1500), extent=(0, 0, 1500, 700))
    og, g = oracle.Geoms.from_rings(x, y, off), core.Geoms.from_polygons(x, y, off)
    for fun, dtype, bg in (("sum", "float32", np.nan), ("last", "float64", 0.0), ("count", "uint16", 0)):
        vals = (np.arange(n) % 11 + 1).astype(dtype)
        exp, _ = oracle.rasterize_dense(og, oracle.raster_info(None, **kw), fun, dtype, vals, None, by, bg, threads=3)
        got, st = core.rasterize_dense(g, core.raster_info(None, **kw), fun, dtype, vals, None, band, len(names), bg,
                                       flags=_lib.FLAG_FORCE_TILE_ENGINE)
        assert st["engine"] == 1 and np.array_equal(exp, got, equal_nan=True), fun
    print("GRID1D_OK")
""")


def test_tile_apply_flattened_grid(tmp_path):
    script = tmp_path / "grid1d.py"
    script.write_text(SCRIPT.format(root=ROOT))
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, RZ_APPLY_1D="1"))
    assert r.returncode == 0 and "GRID1D_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
