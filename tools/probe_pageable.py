"""End to end into a PAGEABLE destination (what the Rust / numpy bindings allocate themselves): the driver's staged
copy (RZ_BOUNCE=0) against the library's bounce blocks + host copy threads.  python tools/probe_pageable.py [c4]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench
from rusterize_b200 import _lib, core

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
w = bench.make_workload(name, 1.0)
fun, dtype, bgv = w["funs"][0]
bg = bench.bg_of(bgv)
rows, cols = w["rows"], w["cols"]
ri = core.raster_info(None, shape=(rows, cols), extent=(0.0, 0.0, float(cols), float(rows)))
ref = None
for mode, env in (("pinned (rz_host_alloc)", None), ("pageable, bounce blocks", {"RZ_BOUNCE": "1"}),
                  ("pageable, bounce 16 threads", {"RZ_BOUNCE": "1", "RZ_BOUNCE_THREADS": "16"}),
                  ("pageable, driver-staged copy", {"RZ_BOUNCE": "0"})):
    for k in ("RZ_BOUNCE", "RZ_BOUNCE_THREADS"):
        os.environ.pop(k, None)
    os.environ.update(env or {})
    for it in range(2):
        out = core.host_empty((1, rows, cols), dtype) if env is None else np.empty((1, rows, cols), dtype)  # fresh pages
        t0 = time.perf_counter()
        st = core.rasterize_dense_soa(w["soa"], ri, fun, dtype, w["field"], None, None, 1, bg, out=out, devices=[0], flags=_lib.FLAG_SYNC_STAGES)[1]
        t1 = time.perf_counter()
        chk = float(np.nansum(out[0, :: max(1, rows // 64)].astype(np.float64)))
        if ref is None:
            ref = chk
        p = st["per_device"][0]
        print(f"{mode:32s} call {it}: {1e3 * (t1 - t0):8.1f} ms  (flatten {p['shard_ms']:6.1f}, burn+copy wall {p['wall_ms']:7.1f}, d2h {p['d2h_ms']:7.1f})  checksum {'ok' if chk == ref else 'DIFFERS'}")
        del out
