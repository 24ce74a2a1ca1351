#!/usr/bin/env python
"""bench.py — benchmark of the B200 burn path on every BASELINE.json config.

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference (oracle)

Headline workload (BASELINE.json configs[3], the one `north_star`'s target is quoted on): 1M synthetic star
polygons (100..300 vertices, seed 4) -> 65536 x 65536 float32 grid, fun=sum, background=NaN.  The other configs
(c1 10k polygons -> 4096^2 f64 sum; c2 100k mixed geometries -> 16384^2 count/any; c3 100k polygons, by = 32
layers, first/last/min/max int32 on 8192^2; c5 10M parcels -> sparse triplets over 131072^2) ride along in the same
JSON line under "other_configs", each with its own ms, roofline, cpu_baseline, e2e and parity_vs_oracle.

With N GPUs a dense raster is split into N row bands (no collective; strong scaling: total work is fixed), each rank
burning the parts the library cuts out for its band (rz_geoms_row_shard); the sparse config is split into N
contiguous geometry ranges inside ONE library call (rz_rasterize_sparse_multi).

One "step" = one full rasterisation.  `value` is device-resident throughput (geometry already in HBM, raster left
in HBM).  `e2e` is the whole call a user makes, from host coordinate arrays to a host raster: flattening, the
upload, the burn and the copy back into ONE pinned host array, all inside one library call
(rz_rasterize_dense_soa) driven by rank 0 over all N GPUs.  Every rank checks rows of its own band (or a sampled
geometry range of the sparse stream) against the CPU oracle, bit for bit, at every N.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

WORKLOADS = {
    "c4": dict(kind="stars", n=1_000_000, vmin=100, vmax=300, rho=48.0, rows=65536, cols=65536, seed=4,
               funs=[("sum", "float32", "nan")], cpu_rows=4096, cpu_threads=1,
               desc="1M star polygons (100-300 vertices) -> 65536x65536 f32, fun=sum, bg=NaN"),
    "c1": dict(kind="stars", n=10_000, vmin=64, vmax=64, rho=82.0, rows=4096, cols=4096, seed=1,
               funs=[("sum", "float64", "nan")], cpu_rows=4096, cpu_threads=1,
               desc="10k 64-vertex star polygons -> 4096x4096 f64, fun=sum, bg=NaN"),
    "c2": dict(kind="mixed", n=100_000, rows=16384, cols=16384, seed=2,
               funs=[("count", "uint32", 0), ("any", "uint8", 0)], cpu_rows=2048, cpu_threads=1,
               desc="100k mixed geometries (60% polygons, 25% lines, 15% points) -> 16384x16384, fun=count u32 / any u8"),
    "c3": dict(kind="layers", n=100_000, vmin=64, vmax=64, rho=256.0, rows=8192, cols=8192, seed=3,
               funs=[("first", "int32", 0), ("last", "int32", 0), ("min", "int32", 0), ("max", "int32", 0)], cpu_rows=512,
               cpu_threads=32,
               desc="100k star polygons (rho 256), by = 32 layers -> 32x8192x8192 i32, fun=first/last/min/max"),
    "c5": dict(kind="parcels", n=10_000_000, rows=131072, cols=131072, seed=5, funs=[("sum", "float32", "nan")],
               cpu_geoms=200_000, cpu_threads=1,
               desc="10M parcels (jittered quads, side 6-14 px) over 131072x131072, sparse triplets, f32 sum"),
    "tiny": dict(kind="stars", n=2_000, vmin=16, vmax=48, rho=24.0, rows=1024, cols=1024, seed=7,
                 funs=[("sum", "float32", "nan")], cpu_rows=1024, cpu_threads=1,
                 desc="2k star polygons -> 1024x1024 f32 (CI smoke size)"),
}


def bg_of(v):
    return np.nan if v == "nan" else v


def make_workload(name: str, scale: float = 1.0):
    """-> dict with the rz_geom_soa arrays (one part and one sequence per geometry), field values, `by` keys."""
    import synth

    w = dict(WORKLOADS[name])
    if scale != 1.0:  # shrink geometry count and grid together (keeps density)
        w["n"] = max(1, int(w["n"] * scale))
        side = max(64, int(w["rows"] * scale ** 0.5) // 64 * 64)
        w["rows"] = w["cols"] = side
        if "cpu_geoms" in w:
            w["cpu_geoms"] = max(1, int(w["cpu_geoms"] * scale))
    n, rows, cols, seed = w["n"], w["rows"], w["cols"], w["seed"]
    w["by"] = None
    if w["kind"] in ("stars", "layers"):
        x, y, off = synth.star_polygons(seed, n, w["vmin"], w["vmax"], w["rho"], cols, rows)
        ar = np.arange(n + 1, dtype=np.uint64)
        w["soa"] = (ar, np.zeros(n, np.uint8), ar, off, x, y)
        if w["kind"] == "layers":
            idx = np.arange(n, dtype=np.int64)
            w["field"] = (1 + (idx * 2654435761 % 10**6)).astype(np.int32)
            w["by"] = [str(i % 32) for i in range(n)]
        else:
            u = synth.splitmix_u(seed, n, 9)
            w["field"] = (100.0 * u if name == "c1" else u).astype(w["funs"][0][1])
    elif w["kind"] == "mixed":
        w["soa"] = synth.config2_soa(seed, n, rows)
        w["field"] = 1
    elif w["kind"] == "parcels":
        x, y, off = synth.parcels(seed, n, cols, rows)
        ar = np.arange(n + 1, dtype=np.uint64)
        w["soa"] = (ar, np.zeros(n, np.uint8), ar, off, x, y)
        w["field"] = synth.splitmix_u(seed, n, 20).astype(np.float32)
    w["n_vertices"] = int(len(w["soa"][4]))
    return w


def band_of(rank: int, world: int, rows: int):
    return rows * rank // world, rows * (rank + 1) // world


def geoms_touching_rows(w, r0, r1, margin=3):
    """Mask of the geometries whose y-extent can touch raster rows [r0, r1) (extent (0,0,cols,rows), res 1: world
    y = rows - pixel y).  Used only to cut the ORACLE's input down to a row slice."""
    sco, y = w["soa"][3].astype(np.int64), w["soa"][5]
    ymin = np.minimum.reduceat(y, sco[:-1])
    ymax = np.maximum.reduceat(y, sco[:-1])
    return (w["rows"] - ymax <= r1 + margin) & (w["rows"] - ymin >= r0 - margin)


def oracle_geoms(w, keep):
    """oracle.Geoms of the kept geometries + their field values and `by` keys (order preserved)."""
    import oracle
    import synth

    sel = synth.soa_select(w["soa"], keep)
    if w["kind"] == "mixed":
        g = oracle.Geoms.from_wkb(synth.soa_to_wkb(sel))
    else:
        g = oracle.Geoms.from_rings(sel[4], sel[5], sel[3])
    field = w["field"][keep] if np.ndim(w["field"]) else w["field"]
    by = None if w["by"] is None else [b for b, k in zip(w["by"], keep) if k]
    return g, field, by, int(keep.sum())


def oracle_rows(w, fun, dtype, bg, r0, r1, threads=1, repeat=1):
    """The oracle's raster rows [r0, r1) of the workload (all bands) -> (array [B][r1-r0][cols], seconds, #geometries)."""
    import oracle

    keep = geoms_touching_rows(w, r0, r1)
    g, field, by, n_s = oracle_geoms(w, keep)
    ri = oracle.raster_info(None, shape=(r1 - r0, w["cols"]), extent=(0.0, float(w["rows"] - r1), float(w["cols"]), float(w["rows"] - r0)))
    if by is not None:  # every band must exist even if the slice misses one: keep the full key set's order
        full = sorted(set(w["by"]))
        have = sorted(set(by))
    best = None
    for _ in range(repeat):
        t = time.perf_counter()
        out, names = oracle.rasterize_dense(g, ri, fun, dtype, field, None, by, bg, False, threads)
        sec = time.perf_counter() - t
        best = sec if best is None else min(best, sec)
    if by is not None and have != full:  # bands without a geometry in the slice stay background
        big = np.full((len(full),) + out.shape[1:], bg, out.dtype)
        for i, nm in enumerate(names):
            big[full.index(nm)] = out[i]
        out = big
    return out, best, n_s


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line).
    NVML is polled in-process every few ms (a timed region of a handful of ~10 ms steps is over before
    `nvidia-smi -lms` prints its first line); nvidia-smi is the fallback when pynvml is missing."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index: int):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                gpu_index = int(vis.split(",")[gpu_index])
            except (ValueError, IndexError):
                pass
        self.idx = gpu_index
        self.lines, self.sm, self.mx, self.reasons = [], [], [], set()
        self.proc = self.thread = self.nvml = None
        self.stop_flag = threading.Event()

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.nvml = pynvml
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)))
            self._poll_once()
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll_once(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for nm, bit in zip(self.NAMES, (0x8, 0x40, 0x20, 0x4)):  # nvml.h: HwSlowdown, HwThermal, SwThermal, SwPowerCap
            if r & bit:
                self.reasons.add(nm)

    def _poll(self):
        while not self.stop_flag.is_set():
            try:
                self._poll_once()
            except Exception:
                break
            time.sleep(0.02)  # (a faster poll competes with the timed loop for the GIL)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": max(self.mx), "reasons": sorted(self.reasons),
                    "samples": len(self.sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        for ln in self.lines:
            p = [s.strip() for s in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                self.sm.append(float(p[1]))
                self.mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(self.NAMES, p[5:9]):
                if v.lower().startswith("active"):
                    self.reasons.add(nm)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvidia-smi"}


def bind_to_gpu_numa_node(gpu_index: int):
    """Pin this rank to the CPUs NVML reports as local to its GPU, so that the pinned host buffers of the
    end-to-end path (geometry pools, output raster) are first-touched on the NUMA node the GPU's PCIe link
    hangs off.  Best effort: returns the CPU list or None."""
    try:
        import pynvml

        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            gpu_index = int(vis.split(",")[gpu_index])
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def config_of(name, w, scale, world, engine):
    """The `config` object of a JSON line (the reference arm prints the same keys)."""
    return {"workload": f"{name}: {w['desc']}", "scale": scale,
            "parallelism": (f"geometry-ranges x{world}" if w["kind"] == "parcels" else f"row-bands x{world}"),
            "l2": "inputs (vertex pools + record buffers + raster) are far larger than the 126 MB L2" if name in ("c4", "c5", "c3")
                  else "every step rewrites the whole raster (larger than L2) from cold record buffers",
            "engine": engine}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (C++ restatement of the reference's CPU algorithm) on a bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_threads_for(w):
    # the reference parallelises over `by` bands only (rust/src/rasterize.rs:89-101)
    return 1 if w["by"] is None else max(1, min(w["cpu_threads"], os.cpu_count() or 1, len(set(w["by"]))))


def cpu_dense_sample(w, fun, dtype, bg, repeat=1):
    rows = min(w["rows"], w["cpu_rows"])
    threads = cpu_threads_for(w)
    out, sec, n_s = oracle_rows(w, fun, dtype, bg, 0, rows, threads, repeat)
    n_b = 1 if w["by"] is None else len(set(w["by"]))
    mpx = n_b * rows * w["cols"] / sec / 1e6
    sample = (f"rows [0,{rows}) of the {n_b}x{w['rows']}x{w['cols']} grid with the {n_s} geometries touching them; C++ "
              f"restatement of the rusterize CPU algorithm (oracle/rz_oracle.cpp), {threads} thread(s): the reference "
              "parallelises over `by` bands only")
    return out, {"value": mpx, "unit": "Mpixel/s", "cores": threads, "kind": "port", "sample": sample, "seconds": sec,
                 "host_cores_available": os.cpu_count()}, rows


def cpu_sparse_sample(w, dtype, bg):
    import oracle

    m = min(w["n"], w["cpu_geoms"])
    keep = np.zeros(w["n"], bool)
    keep[:m] = True
    g, field, by, _ = oracle_geoms(w, keep)
    ri = oracle.raster_info(None, shape=(w["rows"], w["cols"]), extent=(0.0, 0.0, float(w["cols"]), float(w["rows"])))
    t = time.perf_counter()
    sp = oracle.rasterize_sparse(g, ri, "sum", dtype, field, None, None, bg)
    sec = time.perf_counter() - t
    # rate-normalised to the whole extent: the sample's geometries are 1/k of the job
    mpx = w["rows"] * w["cols"] * (m / w["n"]) / sec / 1e6
    return sp, {"value": mpx, "unit": "Mpixel/s", "cores": 1, "kind": "port", "seconds": sec,
                "sample": f"the first {m} of {w['n']} parcels on the full extent (sparse stream), rate-normalised by geometry "
                          "count; oracle/rz_oracle.cpp, 1 thread", "polygons_per_s": m / sec,
                "triplets_per_s": len(sp["rows"]) / sec, "host_cores_available": os.cpu_count()}, m


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name = args.workload
    w = make_workload(name, args.scale)
    fun, dtype, bgv = w["funs"][0]
    bg = bg_of(bgv)
    secs = []
    if w["kind"] == "parcels":
        for i in range(args.warmup + args.steps):
            _, cpu, _ = cpu_sparse_sample(w, dtype, bg)
            if i >= args.warmup:
                secs.append(cpu["seconds"])
    else:
        for i in range(args.warmup + args.steps):
            _, cpu, _ = cpu_dense_sample(w, fun, dtype, bg)
            if i >= args.warmup:
                secs.append(cpu["seconds"])
    sec = float(np.mean(secs))
    mpx = cpu["value"] * cpu["seconds"] / sec
    cpu = dict(cpu, value=mpx, seconds=sec)
    line = {
        "impl": "reference", "metric": "output_Mpixels_per_s", "value": mpx, "unit": "Mpixel/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64 geometry / %s values" % dtype,
        "data": "synthetic", "config": config_of(name, w, args.scale, args.gpus, "tile-binned" if w["kind"] != "mixed" else "crossing-records"),
        "polygons_per_s": w["n"] * (mpx * 1e6 / (w["rows"] * w["cols"] * (1 if w["by"] is None else len(set(w["by"]))))),
        "cpu_baseline": cpu,
        "e2e": {"value": mpx, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Env:
    """Per-process state of the GPU arm: ranks, streams, barriers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the burn path has no CPU fallback (use --impl reference)")
        torch.cuda.set_device(self.local)
        self.numa = bind_to_gpu_numa_node(self.local)  # before any large host allocation
        self.cpu_group = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.cpu_group = dist.new_group(backend="gloo")  # host-side waits and object gathers
        # A dedicated (non-default) stream: the library launches on the stream handle it is given and the timing
        # events are recorded on the same stream.
        self.tstream = torch.cuda.Stream()
        torch.cuda.set_stream(self.tstream)
        self.stream = self.tstream.cuda_stream
        assert self.stream != 0
        self.peak, self.peak_src = peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def host_barrier(self):  # ranks wait on the CPU (an NCCL barrier would spin on the GPUs rank 0 is about to use)
        if self.world > 1:
            self.dist.barrier(group=self.cpu_group)

    def max_over_ranks(self, v):
        t = self.torch.tensor([float(x) for x in np.atleast_1d(v)], device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def sum_over_ranks(self, v):
        t = self.torch.tensor([float(x) for x in np.atleast_1d(v)], device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t)
        return [float(x) for x in t.tolist()]

    def gather_objects(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.cpu_group)
        return out


def roofline_of(env, w, fun, dtype, s0, fill_ms, mask_ms):
    dt = np.dtype(dtype)
    if s0["engine"] == 1:
        # tile engine, span-fill kernel = tile_apply: compact inside-mask blocks (4 B / word) + one 16-byte block
        # descriptor per (part,tile) pair read once, raster written once
        tile_r = 64 if dt.itemsize <= 4 else 32
        kernel = "tile_apply_kernel<%s, %s, %d>" % (dtype, fun, tile_r)
        fill_bytes = s0["n_mask_words"] * 4.0 + s0["n_records"] * 16.0 + s0["out_bytes"]
        # tile_mask: world vertices (16 B) + tags (4 B) read once per mask unit, compact mask blocks written once
        mask_bytes = 20.0 * s0["n_poly_vertices"] + s0["n_mask_words"] * 4.0
        other = {"kernel": "tile_mask_kernel<%d>" % tile_r, "ms_per_launch": mask_ms, "bytes_per_launch": mask_bytes,
                 "achieved": mask_bytes / (mask_ms / 1e3) / 1e9 if mask_ms else None,
                 "frac": mask_bytes / (mask_ms / 1e3) / 1e9 / env.peak if mask_ms else None,
                 "note": "instruction-issue bound (f64 edge math, one crossing per lane)"}
    else:
        # record pipeline: crossing / pixel records (8 B) read once + raster written once
        kernel = "fill_kernel<%s, %s>" % (dtype, fun)
        fill_bytes = 8.0 * s0["n_records"] + s0["out_bytes"]
        other = None
    achieved = fill_bytes / (fill_ms / 1e3) / 1e9 if fill_ms else None
    return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": env.peak, "unit": "GB/s",
            "frac": achieved / env.peak if achieved else None, "traffic": None, "peak_source": env.peak_src,
            "bytes_per_launch": fill_bytes, "ms_per_launch": fill_ms, "second_kernel": other}


def run_dense(env, name, w, headline):
    """One dense config on this rank's row band.  Returns (on rank 0) the result dict."""
    torch = env.torch
    from rusterize_b200 import _lib, core

    args, rank, world, local = env.args, env.rank, env.world, env.local
    rows, cols = w["rows"], w["cols"]
    n_b = 1 if w["by"] is None else len(set(w["by"]))
    r0, r1 = band_of(rank, world, rows)
    band, names = (None, None) if w["by"] is None else core.group_keys(w["by"])
    t_f = time.perf_counter()
    full = core.Geoms.from_soa(*w["soa"])
    flatten_ms = (time.perf_counter() - t_f) * 1e3
    ri = core.raster_info(None, shape=(rows, cols), extent=(0.0, 0.0, float(cols), float(rows)))
    t_s = time.perf_counter()
    geoms = full.row_shard(ri, r0, r1) if world > 1 else full
    shard_ms = (time.perf_counter() - t_s) * 1e3
    geoms.upload(local)
    eng_flag = {"auto": 0, "records": _lib.FLAG_NO_TILE_ENGINE, "tiles": _lib.FLAG_FORCE_TILE_ENGINE}[args.engine]
    steps, warmup = (args.steps, args.warmup) if headline else (max(3, min(args.steps, 5)), max(3, args.warmup))
    per_fun, clocks, each_ms_head, wall_ms_head = {}, None, None, None
    parity = {}
    cpu = None
    s_rows = min(512, r1 - r0)
    for fi, (fun, dtype, bgv) in enumerate(w["funs"]):
        bg = bg_of(bgv)
        dt = np.dtype(dtype)
        tdt = getattr(torch, dt.name)
        d_out = torch.empty((n_b, r1 - r0, cols), dtype=tdt, device="cuda")
        # device-resident arm: every input of the call already lives in HBM - geometry (uploaded above), field values
        # and band ids (RZ_FLAG_INPUTS_ON_DEVICE) - and the raster stays there
        f_np = np.ascontiguousarray(np.asarray(w["field"]).astype(dt)) if np.ndim(w["field"]) else np.array([w["field"]], dt)
        d_field = torch.from_numpy(f_np.view(np.uint8)).cuda()
        d_band = None if band is None else torch.from_numpy(np.ascontiguousarray(band, np.int32)).cuda()
        torch.cuda.synchronize()
        inputs_dev = dict(field=d_field.data_ptr(), scalar=np.ndim(w["field"]) == 0, band=None if d_band is None else d_band.data_ptr())

        def step(flags=0):
            return core.rasterize_dense(geoms, ri, fun, dtype, 0, None, None, n_b, bg, device=local,
                                        rows=(r0, r1), out=d_out.data_ptr(), stream=env.stream, flags=flags | eng_flag,
                                        tile_bytes=args.tile_bytes, inputs_dev=inputs_dev)[1]

        for _ in range(warmup):
            step()
        sampler = ClockSampler(local) if (headline and fi == 0 and rank == 0) else None
        env.barrier()
        if sampler:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall = time.perf_counter()
        ev0.record()
        marks = []
        for _ in range(steps):
            st_timed = step()
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record()
        ev1.record()
        env.barrier()
        wall_ms = (time.perf_counter() - t_wall) * 1e3 / steps  # cross-check of the CUDA-event time
        each_ms = [a.elapsed_time(b) for a, b in zip([ev0] + marks[:-1], marks)]
        if sampler:
            clocks = sampler.stop()
            each_ms_head, wall_ms_head = each_ms, wall_ms
        ms = ev0.elapsed_time(ev1) / steps
        # per-stage CUDA-event timings come from a separate pass: the event synchronisations they need would
        # otherwise sit inside the timed region
        stats = [step(flags=_lib.FLAG_SYNC_STAGES) for _ in range(min(steps, 3))]
        ms_max = env.max_over_ranks(ms)[0]
        keys = ["n_records", "n_crossings", "out_bytes", "kernel_launches", "n_mask_words", "n_poly_vertices"]
        agg = env.sum_over_ranks([float(np.mean([s[k] for s in stats])) for k in keys])
        stage = env.max_over_ranks([float(np.mean([s[k] for s in stats])) for k in
                                    ["count_ms", "emit_ms", "sort_ms", "index_ms", "fill_ms", "total_ms"]])
        # ---- parity: the last rows of this rank's band against the oracle (every rank, every N) ----------
        a0, a1 = r1 - s_rows, r1
        exp, _, n_s = oracle_rows(w, fun, dtype, bg, a0, a1, cpu_threads_for(w))
        got = d_out[:, a0 - r0:a1 - r0].cpu().numpy()
        ok = bool(np.array_equal(exp, got, equal_nan=True))
        par = {"rank": rank, "rows": [int(a0), int(a1)], "geometries": n_s, "bit_exact": ok}
        # ---- CPU baseline (rank 0, N=1): the sample doubles as a parity check of the raster's first rows -------
        if world == 1 and not args.no_cpu and fi == 0:
            o_out, cpu, c_rows = cpu_dense_sample(w, fun, dtype, bg)
            gtop = d_out[:, :c_rows].cpu().numpy()
            cpu["parity_vs_gpu"] = {"bit_exact": bool(np.array_equal(o_out, gtop, equal_nan=True)), "rows": [0, int(c_rows)]}
            par["bit_exact"] = par["bit_exact"] and cpu["parity_vs_gpu"]["bit_exact"]
        parity[fun] = env.gather_objects(par)
        s0 = dict(stats[-1])
        for k, v in zip(keys, agg):
            s0[k] = v / world  # per GPU: a launch's bytes over that launch's duration (the slowest rank's)
        per_fun[fun] = {"dtype": dtype, "ms": ms_max, "Mpixel_per_s": n_b * rows * cols / (ms_max / 1e3) / 1e6,
                        "polygons_per_s": w["n"] / (ms_max / 1e3),
                        "engine": "tile-binned" if s0["engine"] == 1 else "crossing-records",
                        "stage_ms_max_over_ranks": dict(zip(["mask_build" if s0["engine"] == 1 else "count", "emit", "sort",
                                                             "index", "fill", "lib_total_staged_pass"], stage)),
                        "records": int(agg[0]), "crossings": int(agg[1]), "launches_per_step": int(round(agg[3])),
                        "host_syncs_per_step": int(st_timed["host_syncs"]),
                        "roofline": roofline_of(env, w, fun, dtype, s0, stage[4], stage[0]),
                        "whole_step_frac_of_hbm": (20.0 * agg[5] + agg[2]) / (ms_max / 1e3) / 1e9 / env.peak,
                        "_steps": steps, "_warmup": warmup}
        del d_out
        torch.cuda.empty_cache()
    res = {"per_fun": per_fun, "parity": parity, "cpu": cpu, "clocks": clocks, "each_ms": each_ms_head,
           "wall_ms": wall_ms_head, "flatten_ms_rank0": flatten_ms, "row_shard_ms_rank0": shard_ms,
           "n_vertices": w["n_vertices"]}
    del geoms, full
    return res


def run_dense_e2e(env, name, w):
    """Rank 0 only: the whole call from host coordinate arrays to ONE pinned host raster over all N GPUs."""
    torch = env.torch
    from rusterize_b200 import _lib, core

    args, world = env.args, env.world
    fun, dtype, bgv = w["funs"][0]
    bg = bg_of(bgv)
    rows, cols = w["rows"], w["cols"]
    n_b = 1 if w["by"] is None else len(set(w["by"]))
    band, names = (None, None) if w["by"] is None else core.group_keys(w["by"])
    ri = core.raster_info(None, shape=(rows, cols), extent=(0.0, 0.0, float(cols), float(rows)))
    # the caller-owned host raster: one page-locked array (cudaHostAlloc through torch; `--host-alloc rz` takes it from
    # the library's own allocator, rz_host_alloc, instead)
    if args.host_alloc == "rz":
        h_np = core.host_empty((n_b, rows, cols), dtype)
        h_out = torch.from_numpy(h_np)
    else:
        h_out = torch.empty((n_b, rows, cols), dtype=getattr(torch, np.dtype(dtype).name)).pin_memory()
        h_np = h_out.numpy()
    devices = list(range(world))
    eng_flag = {"auto": 0, "records": _lib.FLAG_NO_TILE_ENGINE, "tiles": _lib.FLAG_FORCE_TILE_ENGINE}[args.engine]
    if world > 1:  # this process now drives every GPU: it may run on every core again
        try:
            os.sched_setaffinity(0, range(os.cpu_count()))
        except OSError:
            pass

    def call(g, flags):
        return core.rasterize_dense(g, ri, fun, dtype, w["field"], None, band, n_b, bg, out=h_np, devices=devices,
                                    flags=flags | eng_flag)[1]

    n_e2e = max(1, args.e2e_steps)

    def one_shot(flags):  # rz_rasterize_dense_soa: flatten (each device only its band's parts) + H2D + burn + D2H
        return core.rasterize_dense_soa(w["soa"], ri, fun, dtype, w["field"], None, band, n_b, bg, out=h_np,
                                        devices=devices, flags=flags | eng_flag)[1]

    one_shot(_lib.FLAG_SYNC_STAGES)  # warm: page-locked pools / staging buffers exist, kernels are loaded
    whole, flat, st_last = [], [], None
    for _ in range(n_e2e):
        t0 = time.perf_counter()
        st_last = one_shot(_lib.FLAG_SYNC_STAGES)
        whole.append((time.perf_counter() - t0) * 1e3)
        flat.append(max(p["shard_ms"] for p in st_last["per_device"]))
    stride = max(1, rows // 64)
    a0 = rows - min(512, rows)
    # what the one-shot call left in the host raster (checked against the oracle by the caller); the cached arm and the
    # D2H ceiling probe below write into the same host array afterwards
    tail_one_shot = h_np[:, a0:rows].copy()
    checksum = float(np.nansum(h_np[0, ::stride].astype(np.float64)))
    h_np[:, a0:rows] = 0
    # the same job on an already flattened handle (what a caller who keeps the handle pays): upload forced
    g = core.Geoms.from_soa(*w["soa"])
    call(g, _lib.FLAG_SYNC_STAGES)
    cached, cached_lib = [], []
    for _ in range(n_e2e):
        t0 = time.perf_counter()
        st_c = call(g, _lib.FLAG_SYNC_STAGES | _lib.FLAG_FORCE_H2D)
        cached.append((time.perf_counter() - t0) * 1e3)
        cached_lib.append([round(st_c["h2d_ms"], 1), round(st_c["d2h_ms"], 1), round(st_c["total_ms"], 1), round(st_c["wall_ms"], 1)])
    del g
    tail_cached = h_np[:, a0:rows].copy()
    e_ms = float(np.mean(whole))
    per = st_last["per_device"]
    # the host's ceiling for this raster: the same bytes copied device -> the same pinned array by plain
    # cudaMemcpyAsync from all N devices at once, nothing else running (what the D2H phase of the call cannot beat)
    ceiling = None
    if name == args.workload:
        try:
            flat_out = h_out.view(-1)
            per_dev = flat_out.numel() // world
            bufs, streams = [], []
            for d in devices:
                with torch.cuda.device(d):
                    bufs.append(torch.empty(per_dev, dtype=h_out.dtype, device=f"cuda:{d}"))
                    streams.append(torch.cuda.Stream(device=d))
            best = None
            for _ in range(2):
                for d in devices:
                    torch.cuda.synchronize(d)
                t0 = time.perf_counter()
                for i, d in enumerate(devices):
                    with torch.cuda.device(d), torch.cuda.stream(streams[i]):
                        flat_out[i * per_dev:(i + 1) * per_dev].copy_(bufs[i], non_blocking=True)
                for d in devices:
                    torch.cuda.synchronize(d)
                dt_ = time.perf_counter() - t0
                best = dt_ if best is None else min(best, dt_)
            nbytes = per_dev * world * h_out.element_size()
            ceiling = {"ms": best * 1e3, "GB_per_s": nbytes / best / 1e9, "bytes": nbytes,
                       "note": f"{world} concurrent cudaMemcpyAsync device->pinned host of 1/{world} of the raster each"}
            del bufs
        except Exception as e:  # the probe must never take the bench down
            ceiling = {"error": str(e)[:200]}
    # the same call into a PAGEABLE destination, freshly allocated each time - what a binding that allocates its own
    # array (ndarray's Array3 in the Rust shim) hands over: the library drains the device through two page-locked
    # bounce blocks with host copy threads instead of letting the driver stage the copy
    pageable = None
    if name == args.workload and not args.no_pageable and world == 1:  # (measured on one GPU only)
        flat_out = None
        del h_out, h_np
        try:
            each, same = [], True
            for _ in range(2):
                dst = np.empty((n_b, rows, cols), dtype)
                t0 = time.perf_counter()
                core.rasterize_dense_soa(w["soa"], ri, fun, dtype, w["field"], None, band, n_b, bg, out=dst, devices=devices,
                                         flags=eng_flag)
                each.append(round((time.perf_counter() - t0) * 1e3, 1))
                same = same and bool(np.array_equal(dst[:, a0:rows], tail_one_shot, equal_nan=True))
                del dst
            pageable = {"ms_each_call": each, "last_rows_equal_the_pinned_run": same,
                        "note": "destination = np.empty (fresh pageable pages) per call; page faults of the destination included"}
        except Exception as e:  # never take the bench down
            pageable = {"error": str(e)[:200]}
    return {"value": n_b * rows * cols / (e_ms / 1e3) / 1e6, "unit": "Mpixel/s", "ms_per_step": e_ms, "host_d2h_ceiling": ceiling,
            "pageable_destination": pageable,
            "h2d_bytes_per_step": int(st_c["h2d_bytes"]), "d2h_bytes_per_step": int(st_last["d2h_bytes"]), "steps": n_e2e,
            "includes": "ONE library call from host coordinate arrays to the pinned host raster (rz_rasterize_dense_soa): every "
                        "device flattens the parts of its row band into page-locked pools with the H2D overlapped, burns, "
                        "copies its rows back",
            "flatten_ms": float(np.mean(flat)), "flatten_Gvert_per_s": w["n_vertices"] / (float(np.mean(flat)) / 1e3) / 1e9,
            "flatten_note": "slowest device: y-extent pass + flattening of its band's parts, upload overlapped",
            "ms_each_step": [round(v, 1) for v in whole],
            "e2e_handle_cached": {"ms_per_step": float(np.mean(cached)), "value": n_b * rows * cols / (float(np.mean(cached)) / 1e3) / 1e6,
                                  "ms_each_step": [round(v, 1) for v in cached], "lib_h2d_d2h_total_wall_ms_each_step": cached_lib,
                                  "note": "geometry handle kept between calls (no flattening, part subsets cached), upload forced"},
            "per_device": [{"device": d, "h2d_ms": round(p["h2d_ms"], 2), "d2h_ms": round(p["d2h_ms"], 2),
                            "total_ms": round(p["total_ms"], 2), "wall_ms": round(p["wall_ms"], 2),
                            "shard_ms": round(p["shard_ms"], 2), "h2d_MB": round(p["h2d_bytes"] / 1e6, 1),
                            "d2h_MB": round(p["d2h_bytes"] / 1e6, 1)} for d, p in enumerate(per)],
            "checksum": checksum, "_tails": (a0, tail_one_shot, tail_cached)}


def run_sparse(env, name, w):
    """Config 5: rank 0 drives all N GPUs through rz_rasterize_sparse_multi (contiguous geometry ranges balanced by
    estimated work, streams concatenated by offset in one set of host arrays)."""
    from rusterize_b200 import core

    args, world = env.args, env.world
    fun, dtype, bgv = w["funs"][0]
    bg = bg_of(bgv)
    rows, cols = w["rows"], w["cols"]
    isz = np.dtype(dtype).itemsize
    ri = core.raster_info(None, shape=(rows, cols), extent=(0.0, 0.0, float(cols), float(rows)))
    devices = list(range(world))
    steps = max(2, min(args.steps, 3))
    if world > 1:
        try:
            os.sched_setaffinity(0, range(os.cpu_count()))
        except OSError:
            pass
    g = core.Geoms.from_soa(*w["soa"])
    sp = core.rasterize_sparse(g, ri, fun, dtype, w["field"], background=bg, devices=devices)  # warm (pools, uploads)
    ms, sts = [], []
    for _ in range(steps):
        del sp
        t0 = time.perf_counter()
        sp = core.rasterize_sparse(g, ri, fun, dtype, w["field"], background=bg, devices=devices)
        ms.append((time.perf_counter() - t0) * 1e3)
        sts.append(sp["stats"])
    # the whole call incl. flattening and upload (a caller's previous geometry set is gone: its page-locked pools are
    # recycled by the next one)
    whole, flat = [], []
    del g
    for i in range(max(1, args.e2e_steps) + 1):
        del sp
        t0 = time.perf_counter()
        g = core.Geoms.from_soa(*w["soa"], device=0 if world == 1 else None)
        t1 = time.perf_counter()
        sp = core.rasterize_sparse(g, ri, fun, dtype, w["field"], background=bg, devices=devices)
        if i:  # (the first pass re-creates the pools the warm handle above held differently sized)
            whole.append((time.perf_counter() - t0) * 1e3)
            flat.append((t1 - t0) * 1e3)
        st_e2e = sp["stats"]
        if i < max(1, args.e2e_steps):
            del g
    st = sts[-1]
    P = int(len(sp["rows"]))
    X = int(st["n_crossings"])
    t_ms = float(np.mean(ms))
    fill_ms = float(np.mean([s["fill_ms"] for s in sts]))
    a_fill = 8.0 * X + P * (16.0 + isz)
    # ---- parity: sampled geometry ranges of the stream against the oracle --------------------------------
    import oracle

    n = w["n"]
    m = min(n, 20_000)
    checks = []
    starts = sorted(set([0] + [n * i // world for i in range(1, world)] + [n - m]))
    if len(starts) > 5:
        starts = starts[:2] + starts[len(starts) // 2:len(starts) // 2 + 1] + starts[-2:]
    ori = oracle.raster_info(None, shape=(rows, cols), extent=(0.0, 0.0, float(cols), float(rows)))
    import synth

    for a in starts:
        a = max(0, min(a, n - m))
        # offset of geometry a's first triplet = number of triplets of geometries [0, a): counted by a GPU call
        if a == 0:
            off = 0
        elif a == n - m:
            off = None  # the tail: aligned at the end of the stream
        else:
            keep = np.zeros(n, bool)
            keep[:a] = True
            sel = synth.soa_select(w["soa"], keep)
            gh = core.Geoms.from_soa(*sel)
            off = int(core.rasterize_sparse(gh, ri, fun, dtype, w["field"][:a], background=bg, devices=devices)["counts"][0])
            del gh
        keep = np.zeros(n, bool)
        keep[a:a + m] = True
        og, field, _, _ = oracle_geoms(w, keep)
        osp = oracle.rasterize_sparse(og, ori, fun, dtype, field, None, None, bg)
        k = len(osp["rows"])
        lo = P - k if off is None else off
        ok = all(np.array_equal(np.asarray(osp[key]), np.asarray(sp[key][lo:lo + k])) for key in ("rows", "cols", "data"))
        checks.append({"geometries": [int(a), int(a + m)], "triplets": int(k), "stream_offset": int(lo), "bit_exact": bool(ok)})
    cpu = None
    if not args.no_cpu:
        osp, cpu, m_c = cpu_sparse_sample(w, dtype, bg)
        k = len(osp["rows"])
        cpu["parity_vs_gpu"] = {"bit_exact": all(np.array_equal(np.asarray(osp[key]), np.asarray(sp[key][:k])) for key in ("rows", "cols", "data")),
                                "geometries": [0, int(m_c)], "triplets": int(k)}
    per = st["per_device"]
    e_ms = float(np.mean(whole))
    res = {"ms": t_ms, "Mpixel_per_s": rows * cols / (t_ms / 1e3) / 1e6, "polygons_per_s": n / (t_ms / 1e3),
           "triplets": P, "triplets_per_s": P / (t_ms / 1e3), "crossings": X,
           "note": "sparse output always lands in host memory: `ms` is the call on an uploaded handle (scans, sort, expand, D2H)",
           "stage_ms_max_over_devices": {"crossings+sort": st["sort_ms"], "scans": st["index_ms"], "expand": st["fill_ms"], "d2h": st["d2h_ms"]},
           "d2h_share": st["d2h_ms"] / max(st["total_ms"], 1e-9),
           "roofline": {"bound": "hbm", "kernel": "poly_expand_kernel<f32> (triplet expand)", "achieved": a_fill / world / (fill_ms / 1e3) / 1e9,
                        "peak": env.peak, "unit": "GB/s", "frac": a_fill / world / (fill_ms / 1e3) / 1e9 / env.peak, "traffic": None,
                        "peak_source": env.peak_src, "bytes_per_launch": a_fill / world, "ms_per_launch": fill_ms,
                        "formula": "A_fill(sparse) = 8*X + P*(16+s), per device"},
           "cpu_baseline": cpu,
           "e2e": {"value": rows * cols / (e_ms / 1e3) / 1e6, "unit": "Mpixel/s", "ms_per_step": e_ms,
                   "h2d_bytes_per_step": int(st_e2e["h2d_bytes"]), "d2h_bytes_per_step": int(st_e2e["d2h_bytes"]),
                   "includes": "rz_geoms_from_soa + geometry-range subsets + H2D + scans/sort/expand + D2H into one triplet stream",
                   "ms_each_step": [round(v, 1) for v in whole], "flatten_ms": float(np.mean(flat)),
                   "per_device_wall_shard_ms": [[round(p["wall_ms"], 1), round(p["shard_ms"], 1)] for p in st_e2e["per_device"]]},
           "per_device": [{"device": d, "total_ms": round(p["total_ms"], 2), "wall_ms": round(p["wall_ms"], 2),
                           "shard_ms": round(p["shard_ms"], 2), "d2h_ms": round(p["d2h_ms"], 2),
                           "expand_ms": round(p["fill_ms"], 2), "triplet_MB": round(p["out_bytes"] / 1e6, 1)} for d, p in enumerate(per)],
           "parity_vs_oracle": {"bit_exact": all(c["bit_exact"] for c in checks), "checks": checks},
           "config": config_of(name, w, args.scale, world, "sparse: scans + radix sort + expand"), "n_gpus": world}
    del sp, g
    return res


def summarize_dense(env, name, w, res, e2e):
    """Rank 0: the per-config object of the JSON line."""
    first = w["funs"][0][0]
    pf = res["per_fun"]
    head = pf[first]
    par = {f: {"bit_exact": all(p["bit_exact"] for p in ranks), "per_rank": ranks} for f, ranks in res["parity"].items()}
    out = {"ms": head["ms"], "Mpixel_per_s": head["Mpixel_per_s"], "polygons_per_s": head["polygons_per_s"],
           "fun": first, "roofline": head["roofline"], "cpu_baseline": res["cpu"], "e2e": e2e,
           "parity_vs_oracle": {"bit_exact": all(v["bit_exact"] for v in par.values()), "per_fun": par},
           "per_fun": {f: {k: v for k, v in d.items() if not k.startswith("_")} for f, d in pf.items()},
           "config": config_of(name, w, env.args.scale, env.world, head["engine"]), "n_gpus": env.world,
           "steps": head["_steps"], "warmup": head["_warmup"]}
    return out


def run_b200(args):
    env = Env(args)
    torch = env.torch
    names = [args.workload] + [c for c in args.others.split(",") if c and c != "none" and c != args.workload]
    results = {}
    for name in names:
        # every workload starts with an empty host pool: the page-locked blocks the previous one left behind (its 17 GB
        # raster, its vertex pools) go back to the system here, outside every timed region
        import gc

        from rusterize_b200 import core as _core

        gc.collect()
        pooled = _core.host_trim(1 << 62)
        _core.host_trim(0)
        if env.rank == 0:
            print(f"[bench] {name}: host pool held {pooled / 1e9:.2f} GB of free page-locked blocks, trimmed", file=sys.stderr)
        w = make_workload(name, args.scale)
        headline = name == args.workload
        if w["kind"] == "parcels":
            env.barrier()
            env.host_barrier()
            if env.rank == 0:
                results[name] = run_sparse(env, name, w)
            env.host_barrier()
            del w
            continue
        res = run_dense(env, name, w, headline)
        torch.cuda.empty_cache()
        env.barrier()
        env.host_barrier()
        e2e = None
        if env.rank == 0 and not args.no_e2e:
            e2e = run_dense_e2e(env, name, w)
            # the host raster must equal the oracle rows checked above (rank 0's last band rows) - via the device copy
            a0, tail_one_shot, tail_cached = e2e.pop("_tails")
            fun, dtype, bgv = w["funs"][0]
            a1 = w["rows"]
            exp, _, _ = oracle_rows(w, fun, dtype, bg_of(bgv), a0, a1, cpu_threads_for(w))
            e2e["host_raster_vs_oracle"] = {"rows": [int(a0), int(a1)],
                                            "bit_exact": bool(np.array_equal(exp, tail_one_shot, equal_nan=True)),
                                            "handle_cached_bit_exact": bool(np.array_equal(exp, tail_cached, equal_nan=True))}
            del tail_one_shot, tail_cached
        env.host_barrier()
        if env.rank == 0:
            results[name] = (summarize_dense(env, name, w, res, e2e), res)
        del w

    if env.rank != 0:
        if env.world > 1:
            env.dist.destroy_process_group()
        return 0

    hname = args.workload
    hw = WORKLOADS[hname]
    if hw["kind"] == "parcels":
        body = results[hname]
        line = {"metric": "output_Mpixels_per_s", "value": body["Mpixel_per_s"], "unit": "Mpixel/s", "n_gpus": env.world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": body["ms"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64 geometry / float32 values", "data": "synthetic",
                "gpu_launches": None}
        line.update(body)
    else:
        summ, res = results[hname]
        head = res["per_fun"][hw["funs"][0][0]]
        tp = ROOT / "profiles" / "fill_traffic.json"
        if tp.exists() and env.world == 1 and args.scale == 1.0:
            try:
                summ["roofline"]["traffic"] = json.loads(tp.read_text()).get(hname)
            except Exception:
                pass
        line = {
            "metric": "output_Mpixels_per_s", "value": head["Mpixel_per_s"], "unit": "Mpixel/s",
            "n_gpus": env.world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms"],
            "wall_ms_per_step_rank0": res["wall_ms"], "ms_each_step_rank0": [round(v, 3) for v in (res["each_ms"] or [])],
            "lib_total_ms_staged_pass": head["stage_ms_max_over_ranks"]["lib_total_staged_pass"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 geometry / %s values" % hw["funs"][0][1], "data": "synthetic",
            "config": summ["config"], "polygons_per_s": head["polygons_per_s"],
            "stage_ms_max_over_ranks": head["stage_ms_max_over_ranks"],
            "records": head["records"], "crossings": head["crossings"],
            "gpu_launches": head["launches_per_step"] * args.steps,
            "host_syncs_per_step": head["host_syncs_per_step"],
            "whole_step_frac_of_hbm": head["whole_step_frac_of_hbm"],
            "flatten_ms_rank0": res["flatten_ms_rank0"], "row_shard_ms_rank0": res["row_shard_ms_rank0"],
            "roofline": summ["roofline"], "cpu_baseline": summ["cpu_baseline"], "e2e": summ["e2e"],
            "parity_vs_oracle": summ["parity_vs_oracle"], "clocks": res["clocks"],
            "rank0_cpu_affinity": (f"{len(env.numa)} CPUs local to the GPU (NVML)" if env.numa else "unchanged"),
        }
    others = {}
    for name in names:
        if name == hname:
            continue
        r = results[name]
        others[name] = r if isinstance(r, dict) else r[0]
    line["other_configs"] = others
    print(json.dumps(line))
    if env.world > 1:
        env.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--others", default="c1,c2,c3,c5", help="configs reported under other_configs (comma list, or none)")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink geometry count and grid area by this factor")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--tile-bytes", type=int, default=0)
    ap.add_argument("--engine", default="auto", choices=["auto", "records", "tiles"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-pageable", action="store_true", help="skip the e2e variant that writes a pageable destination")
    ap.add_argument("--host-alloc", default="torch", choices=["torch", "rz"], help="who allocates the pinned host raster of the e2e arm")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
