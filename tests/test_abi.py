"""CPU: libscpb.so loads and exports every symbol include/scpb.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "scpb.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(scpb_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_symbols():
    syms = _declared_symbols()
    assert "scpb_discretize" in syms and "scpb_create" in syms and len(syms) >= 8


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.lib.load()
    for s in _declared_symbols():
        assert hasattr(lib, s), f"libscpb.so does not export {s}"
        assert s in pkg.lib.SIGNATURES, f"lib.py has no ctypes signature for {s}"


def test_signature_table_matches_header(pkg):
    assert sorted(pkg.lib.SIGNATURES) == _declared_symbols()


def test_version(pkg):
    assert pkg.lib.load().scpb_version() >= 100


def test_no_cpu_fallback_without_gpu(pkg):
    """Without a CUDA device the product path must fail loudly, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(pkg.ScpbError):
        pkg.Handle(0)


def test_product_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "scptoolbox.jl_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU fallback", ""), f"{f} references the oracle"


def test_header_is_plain_c(tmp_path):
    """include/scpb.h is the drop-in boundary: it must compile as C99 (what ccall / cgo / ctypes bind), not only as C++."""
    import subprocess
    src = tmp_path / "use_scpb.c"
    src.write_text('#include "scpb.h"\n'
                   'int use(void) { scpb_handle h = 0; scpb_cone_opts o; scpb_ptr_desc d; scpb_scvx_desc v;\n'
                   '  (void)o; (void)d; (void)v; return scpb_create(0, &h) == SCPB_OK ? (int)sizeof(o) : (int)sizeof(d); }\n')
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only",
                        "-I", os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
