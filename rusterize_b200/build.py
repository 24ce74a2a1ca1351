"""Builds librz_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
SO = HERE / "librz_b200.so"
SOURCES = [CSRC / "rz_engine.cu", CSRC / "rz_host.cpp"]
HEADERS = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.hpp")) + [HERE.parent / "include" / "rz_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",  # the reference never fuses mul+add (rust/src/geo/edges.rs:50-55)
    "-Xcompiler", "-fPIC,-O3,-Wall,-ffp-contract=off",
    "-shared", "-cudart", "static",
]


def needs_build() -> bool:
    if not SO.exists():
        return True
    t = SO.stat().st_mtime
    return any(p.exists() and p.stat().st_mtime > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", str(SO), *map(str, SOURCES)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building librz_b200.so")
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
