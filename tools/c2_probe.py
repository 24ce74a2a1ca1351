"""Timing probe for BASELINE configs 2 and 3 (parity-test configurations, not bench lines): device-resident
milliseconds per call and the engine that ran."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
import synth  # noqa: E402
from rusterize_b200 import _lib, core  # noqa: E402
from test_gpu_fullsize import _config2  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        st = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3, st


size = 16384
g = core.Geoms.from_wkb(_config2())
ri = core.raster_info(None, shape=(size, size), extent=(0, 0, size, size))
for fun, dt in (("count", "uint32"), ("any", "uint8")):
    out = torch.empty((1, size, size), dtype={"uint32": torch.int32, "uint8": torch.uint8}[dt], device="cuda")
    ms, st = timed(lambda: core.rasterize_dense(g, ri, fun, dt, 1, background=0, out=out.data_ptr(), flags=_lib.FLAG_SYNC_STAGES)[1])
    print(f"c2 {fun}/{dt}: {ms:.2f} ms  engine={st['engine']} records={st['n_records']} "
          f"count={st['count_ms']:.2f} emit={st['emit_ms']:.2f} sort={st['sort_ms']:.2f} fill={st['fill_ms']:.2f}")

size, n = 8192, 100_000
x, y, off = synth.star_polygons(3, n, 64, 64, 256.0, size, size)
field = (1 + (np.arange(n, dtype=np.int64) * 2654435761 % 10**6)).astype(np.int32)
band, names = core.group_keys([str(i % 32) for i in range(n)])
g3 = core.Geoms.from_polygons(x, y, off)
ri3 = core.raster_info(None, shape=(size, size), extent=(0, 0, size, size))
out3 = torch.empty((32, size, size), dtype=torch.int32, device="cuda")
for extra in (0, _lib.FLAG_FORCE_TILE_ENGINE):
    for fun in ("first", "last", "min", "max"):
        ms, st = timed(lambda: core.rasterize_dense(g3, ri3, fun, "int32", field, None, band, 32, 0, out=out3.data_ptr(),
                                                    flags=_lib.FLAG_SYNC_STAGES | extra)[1], reps=3)
        print(f"c3 {fun}/int32 x32 bands: {ms:.2f} ms  engine={st['engine']} records={st['n_records']} "
              f"mask/count={st['count_ms']:.2f} sort={st['sort_ms']:.2f} fill={st['fill_ms']:.2f}")

# config 5, one GPU's share of the 8-GPU plan: 1.25M of the 10M parcels (a contiguous geometry range) over the
# full 131072 x 131072 extent, sparse encoding (triplets copied back to the host inside the call)
size, n = 131072, 1_250_000
x, y, off = synth.parcels(5, n, size, size)
vals = synth.splitmix_u(5, n, 20).astype(np.float32)
g5 = core.Geoms.from_polygons(x, y, off)
g5.upload(0)
ri5 = core.raster_info(None, shape=(size, size), extent=(0, 0, size, size))
for rep in range(3):  # results are dropped between calls: from the second call on their host blocks are recycled
    t = time.perf_counter()
    sp = core.rasterize_sparse(g5, ri5, "sum", "float32", vals, background=np.nan)
    dt = time.perf_counter() - t
    st, ntrip = sp["stats"], len(sp["rows"])
    del sp
    print(f"c5 share, call {rep}: {n} parcels -> {ntrip} triplets in {1e3*dt:.1f} ms "
          f"(device+D2H total_ms={st['total_ms']:.1f}: crossings+sort {st['sort_ms']:.1f}, scans {st['index_ms']:.1f}, "
          f"expand {st['fill_ms']:.1f}, alloc+D2H {st['d2h_ms']:.1f}; d2h bytes={st['d2h_bytes']/1e9:.2f} GB, launches={st['kernel_launches']})")
