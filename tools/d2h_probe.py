"""End-to-end config 4 with one or two copy streams per window (RZ_COPY_STREAMS), same process, alternating."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import bench  # noqa: E402
from rusterize_b200 import _lib, core  # noqa: E402

w, x, y, off, vals = bench.make_workload("c4")
g = core.Geoms.from_polygons(x, y, off)
ri = core.raster_info(None, shape=(w["rows"], w["cols"]), extent=(0.0, 0.0, float(w["cols"]), float(w["rows"])))
h = torch.empty((1, w["rows"], w["cols"]), dtype=torch.float32).pin_memory().numpy()
for rep in range(3):
    for n in ("1", "2"):
        os.environ["RZ_COPY_STREAMS"] = n
        t = time.perf_counter()
        st = core.rasterize_dense(g, ri, "sum", "float32", vals, background=np.nan, out=h,
                                  flags=_lib.FLAG_FORCE_H2D | _lib.FLAG_SYNC_STAGES)[1]
        print(f"copy streams {n}: e2e {1e3*(time.perf_counter()-t):.1f} ms  h2d {st['h2d_ms']:.1f}  d2h span {st['d2h_ms']:.1f}", flush=True)
