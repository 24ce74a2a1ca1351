"""include/rz_b200.h as a C caller sees it: tests/abi_smoke.c is compiled with gcc against the header and linked to
librz_b200.so; and the hand-written struct mirrors (ctypes in rusterize_b200/_lib.py, #[repr(C)] in
integration/rusterize-b200-sys/src/lib.rs) are checked against the library's own rz_abi_layout()."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def abi_smoke(tmp_path_factory):
    from rusterize_b200 import build

    so = build.build()
    exe = tmp_path_factory.mktemp("abi") / "abi_smoke"
    libdir = os.path.dirname(str(so))
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "abi_smoke.c"), "-o", str(exe), "-L", libdir, "-lrz_b200",
                        f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return str(exe)


def test_c_caller_compiles_links_and_sees_the_headers_layout(abi_smoke):
    """No GPU needed: struct sizes of the header as gcc lays them out == the library's; WKT ingestion, grid math and
    the reference's length-mismatch ValueError string through plain C."""
    r = subprocess.run([abi_smoke], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "abi ok" in r.stdout


@pytest.mark.gpu
def test_c_caller_burns_the_reference_fixture(abi_smoke):
    """C caller -> golden raster of the reference (histogram of python/test/data/standard_output_sum.tif) and the
    29 363-triplet sparse stream of python/docs/python.md."""
    r = subprocess.run([abi_smoke, "burn"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "matches the golden raster" in r.stdout


def test_ctypes_mirrors_match_the_library():
    from rusterize_b200 import _lib

    L = _lib.lib()  # (lib() itself refuses to load a library whose layout differs)
    theirs = (C.c_uint64 * 16)()
    assert L.rz_abi_layout(theirs, 16) == 16
    assert list(theirs) == _lib.abi_layout_of_bindings()


def test_rust_mirrors_list_the_same_fields_in_the_same_order():
    """The Rust crate cannot be compiled here (no cargo): compare its #[repr(C)] field lists with the header's,
    field by field, and its extern block with the header's function list."""
    hdr = open(os.path.join(ROOT, "include", "rz_b200.h")).read()
    rs = open(os.path.join(ROOT, "integration", "rusterize-b200-sys", "src", "lib.rs")).read()

    def c_fields(name):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(",")
            for i, nm in enumerate(names):
                nm = nm.strip().split()[-1] if i == 0 else nm.strip()
                out.append(re.sub(r"\[\d*\]|\*", "", nm))
        return out

    def rs_fields(name):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % name, rs, re.S).group(1)
        return re.findall(r"pub (\w+):", body)

    for c, r in [("rz_raster_info", "RzRasterInfo"), ("rz_raw_raster_info", "RzRawRasterInfo"),
                 ("rz_geom_soa", "RzGeomSoa"), ("rz_context", "RzContext"), ("rz_stats", "RzStats")]:
        assert c_fields(c) == rs_fields(r), (c, c_fields(c), rs_fields(r))
    c_funcs = set(re.findall(r"\b(rz_\w+)\(", re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)))
    rs_funcs = set(re.findall(r"pub fn (rz_\w+)\(", rs))
    assert c_funcs == rs_funcs, (c_funcs ^ rs_funcs)
