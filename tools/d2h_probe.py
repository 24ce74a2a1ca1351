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
ref_tail = None
for rep in range(2):
    for n in ("1", "2", "2"):
        os.environ["RZ_COPY_STREAMS"] = n
        h[0, -4096:] = 7.0  # scramble the rows that land last: a call returning before its copies end would show
        t = time.perf_counter()
        st = core.rasterize_dense(g, ri, "sum", "float32", vals, background=np.nan, out=h,
                                  flags=_lib.FLAG_FORCE_H2D | _lib.FLAG_SYNC_STAGES)[1]
        print(f"copy streams {n}: e2e {1e3*(time.perf_counter()-t):.1f} ms  lib total {st['total_ms']:.1f}  h2d {st['h2d_ms']:.1f}  "
              f"d2h span {st['d2h_ms']:.1f}  stages mask {st['count_ms']:.1f} fill {st['fill_ms']:.1f} bin {st['emit_ms']:.1f} "
              f"sort {st['sort_ms']:.1f}  windows {st['n_windows']}", flush=True)
        tail = h[0, -4096:].copy()
        if ref_tail is None:
            ref_tail = tail
        print("   tail rows equal to the first call's:", bool(np.array_equal(tail, ref_tail, equal_nan=True)),
              " still scrambled:", int((tail == 7.0).sum()), flush=True)
