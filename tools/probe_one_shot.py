"""Where does a one-shot call spend its time on small workloads?  python tools/probe_one_shot.py c2"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench
from rusterize_b200 import _lib, core

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = bench.make_workload(name, 1.0)
fun, dtype, bgv = w["funs"][0]
bg = bench.bg_of(bgv)
rows, cols = w["rows"], w["cols"]
n_b = 1 if w["by"] is None else len(set(w["by"]))
band, names = (None, None) if w["by"] is None else core.group_keys(w["by"])
ri = core.raster_info(None, shape=(rows, cols), extent=(0.0, 0.0, float(cols), float(rows)))
h_out = torch.empty((n_b, rows, cols), dtype=getattr(torch, np.dtype(dtype).name)).pin_memory()
h_np = h_out.numpy()
print(name, [(a.dtype, a.flags.c_contiguous, a.shape) for a in w["soa"]])
for it in range(7):
    t0 = time.perf_counter()
    st = core.rasterize_dense_soa(w["soa"], ri, fun, dtype, w["field"], None, band, n_b, bg, out=h_np, devices=[0], flags=_lib.FLAG_SYNC_STAGES)[1]
    t1 = time.perf_counter()
    p = st["per_device"][0]
    print(f"one-shot {it}: python {1e3 * (t1 - t0):7.2f}  C call {st['wall_ms']:7.2f}  flatten {p['shard_ms']:6.2f}  burn+copy {p['wall_ms']:6.2f}")
for it in range(7):
    t0 = time.perf_counter()
    g = core.Geoms.from_soa(*w["soa"], device=0)
    t1 = time.perf_counter()
    st = core.rasterize_dense(g, ri, fun, dtype, w["field"], None, band, n_b, bg, out=h_np, devices=[0], flags=_lib.FLAG_SYNC_STAGES)[1]
    t2 = time.perf_counter()
    del g
    t3 = time.perf_counter()
    print(f"two-step {it}: python {1e3 * (t3 - t0):7.2f}  flatten {1e3 * (t1 - t0):6.2f}  call {1e3 * (t2 - t1):6.2f} (C {st['wall_ms']:6.2f})  free {1e3 * (t3 - t2):6.2f}")
