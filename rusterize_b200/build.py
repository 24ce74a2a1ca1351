"""Builds librz_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

The library is several translation units (orchestration + C ABI, host flattener, and the dtype x pixel-function
kernel families) compiled in parallel into rusterize_b200/_obj/ and linked; only stale objects are rebuilt."""
from __future__ import annotations

import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = HERE / "_obj"
SO = HERE / "librz_b200.so"
SOURCES = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cpp"))
PUBLIC_HEADER = HERE.parent / "include" / "rz_b200.h"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",  # the reference never fuses mul+add (rust/src/geo/edges.rs:50-55)
    "-Xcompiler", "-fPIC,-O3,-Wall,-ffp-contract=off,-pthread",
]
LINK_FLAGS = ["-shared", "-cudart", "static", "-Xcompiler", "-pthread"]

_INC = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def _deps(src: Path, seen=None) -> set:
    """Transitive quoted includes of a source file."""
    seen = set() if seen is None else seen
    for name in _INC.findall(src.read_text()):
        p = (src.parent / name).resolve()
        if p.exists() and p not in seen:
            seen.add(p)
            _deps(p, seen)
    return seen


def _stale(src: Path, obj: Path) -> bool:
    if not obj.exists():
        return True
    t = obj.stat().st_mtime
    return any(p.stat().st_mtime > t for p in [src, Path(__file__), *_deps(src)])


def needs_build() -> bool:
    """The library is stale when any source or header is newer than it (the objects are a local cache only: they
    do not travel to the GPU box, the built .so does)."""
    if not SO.exists():
        return True
    t = SO.stat().st_mtime
    files = set(SOURCES) | {PUBLIC_HEADER}
    for s_ in SOURCES:
        files |= _deps(s_)
    return any(p.stat().st_mtime > t for p in files)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    OBJ.mkdir(exist_ok=True)
    todo = [s for s in SOURCES if force or _stale(s, OBJ / (s.stem + ".o"))]

    def compile_one(src: Path):
        cmd = [nvcc, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-c", "-o", str(OBJ / (src.stem + ".o")), str(src)]
        return src, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        results = list(ex.map(compile_one, todo))
    for src, r in results:
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"nvcc failed compiling {src.name}")
        if verbose:
            sys.stderr.write(r.stdout + r.stderr)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", *LINK_FLAGS, "-o", str(SO),
           *[str(OBJ / (s.stem + ".o")) for s in SOURCES]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking librz_b200.so")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
