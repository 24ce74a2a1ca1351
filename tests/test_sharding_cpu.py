"""Host-side logic of the multi-GPU path on CPU: row-band planning, the library's per-band part selection
(rz_geoms_row_shard) and assembly, run as a world-size-2 gloo job.  The compute stand-in is the oracle (this is
tests/): the CUDA path itself needs a GPU and is covered by test_gpu_parity.py::test_row_band_shards_equal_full_raster
and tests/test_gpu_multi.py."""
import os
import subprocess
import sys
import textwrap

import numpy as np

import bench
import oracle
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_band_plan_covers_rows_exactly():
    for rows in (1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            bands = [bench.band_of(r, world, rows) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == rows
            assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))


def _shard_rings(w, r0, r1):
    """This rank's share of a row-band sharded job, as the LIBRARY cuts it (rz_geoms_row_shard, host code): the
    shard's polygon rings and the field value of each (by its unchanged geometry index)."""
    from rusterize_b200 import core

    full = core.Geoms.from_soa(*w["soa"])
    ri = core.raster_info(None, shape=(w["rows"], w["cols"]), extent=(0, 0, w["cols"], w["rows"]))
    sh = full.row_shard(ri, r0, r1)
    x, y, tag = sh.pool(0)
    ends = np.flatnonzero(tag & 0x80000000)
    off = np.concatenate([[0], ends + 1]).astype(np.uint64)
    _, geom = sh.parts()
    part_of_ring = (tag[ends] & 0x3FFFFFFF).astype(np.int64)
    return x, y, off, w["field"][geom[part_of_ring].astype(np.int64)], sh.n_parts


def test_band_selection_keeps_every_touching_polygon_in_order():
    w = bench.make_workload("tiny")
    _, _, _, off, x, y = w["soa"]
    full = oracle.rasterize_dense(oracle.Geoms.from_rings(x, y, off), oracle.raster_info(
        None, shape=(w["rows"], w["cols"]), extent=(0, 0, w["cols"], w["rows"])), "sum", "float32", w["field"],
        background=np.nan)[0]
    for r0, r1 in [(0, 100), (100, 612), (612, 1024)]:
        bx, by, boff, bvals, n_parts = _shard_rings(w, r0, r1)
        assert n_parts < w["n"]
        ri = oracle.raster_info(None, shape=(r1 - r0, w["cols"]), extent=(0, w["rows"] - r1, w["cols"], w["rows"] - r0))
        band = oracle.rasterize_dense(oracle.Geoms.from_rings(bx, by, boff), ri, "sum", "float32", bvals,
                                      background=np.nan)[0]
        assert np.array_equal(band[0], full[0, r0:r1], equal_nan=True)
    # the oracle-side cut bench.py uses for its parity slices keeps at least what the library keeps
    keep = bench.geoms_touching_rows(w, 100, 612)
    assert keep.sum() >= _shard_rings(w, 100, 612)[4]


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import numpy as np, torch, torch.distributed as dist
    import bench, oracle
    from test_sharding_cpu import _shard_rings
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    w = bench.make_workload("tiny")
    r0, r1 = bench.band_of(rank, world, w["rows"])
    bx, by, boff, bvals, _ = _shard_rings(w, r0, r1)
    ri = oracle.raster_info(None, shape=(r1 - r0, w["cols"]), extent=(0, w["rows"] - r1, w["cols"], w["rows"] - r0))
    band = oracle.rasterize_dense(oracle.Geoms.from_rings(bx, by, boff), ri, "sum", "float32", bvals, background=np.nan)[0]
    parts = [torch.empty((1, b[1] - b[0], w["cols"])) for b in (bench.band_of(r, world, w["rows"]) for r in range(world))]
    dist.all_gather(parts, torch.from_numpy(band))
    if rank == 0:
        _, _, _, off, x, y = w["soa"]
        full = oracle.rasterize_dense(oracle.Geoms.from_rings(x, y, off), oracle.raster_info(
            None, shape=(w["rows"], w["cols"]), extent=(0, 0, w["cols"], w["rows"])), "sum", "float32", w["field"],
            background=np.nan)[0]
        got = torch.cat(parts, 1).numpy()
        assert np.array_equal(got, full, equal_nan=True)
        print("SHARDS_OK")
    dist.destroy_process_group()
""")


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARDS_OK" in r.stdout


def test_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    import json

    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0


def test_sparse_geometry_range_shards_concatenate():
    """Sparse output is ordered band -> geometry -> burn order (writers.rs:101-131), so the multi-GPU plan for
    sparse jobs is contiguous geometry ranges whose streams concatenate.  Host logic checked with
    the oracle standing in for the device."""
    n, size = 3000, 1024
    x, y, off = synth.parcels(11, n, size, size)
    vals = synth.splitmix_u(11, n, 20).astype(np.float32)
    ri = oracle.raster_info(None, shape=(size, size), extent=(0, 0, size, size))
    full = oracle.rasterize_sparse(oracle.Geoms.from_rings(x, y, off), ri, "sum", "float32", vals, None, None, np.nan)
    parts = []
    for d in range(3):
        a, b = n * d // 3, n * (d + 1) // 3
        o = off[a:b + 1]
        g = oracle.Geoms.from_rings(x[int(o[0]):int(o[-1])], y[int(o[0]):int(o[-1])], o - o[0])
        parts.append(oracle.rasterize_sparse(g, ri, "sum", "float32", vals[a:b], None, None, np.nan))
    for k in ("rows", "cols", "data"):
        assert np.array_equal(np.concatenate([p[k] for p in parts]), full[k]), k
