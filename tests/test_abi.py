"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/cgb200.h
declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

import cryogrid_jl_b200 as cg
from cryogrid_jl_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "cgb200.h")).read()
    declared = set(re.findall(r"\b(cgb_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"cgb_handle", "cgb_config", "cgb_stats"}
    assert len(declared) >= 24
    lib = C.CDLL(L.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in cgb200.h but not exported"
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)


def test_abi_version():
    assert L.load().cgb_abi_version() == L.ABI_VERSION


def test_config_struct_layout(tmp_path):
    # the ctypes mirrors must match the C structs of include/cgb200.h as a C compiler lays them out
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "cgb200.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(cgb_config), sizeof(cgb_stats), '
                   'offsetof(cgb_config, n_columns), offsetof(cgb_config, maxdelta), offsetof(cgb_config, lite_tol));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert [int(x) for x in out] == [C.sizeof(L.Config), C.sizeof(L.Stats), L.Config.n_columns.offset,
                                     L.Config.maxdelta.offset, L.Config.lite_tol.offset]


@pytest.mark.skipif(_has_gpu(), reason="GPU present")
def test_no_cpu_fallback():
    lib = L.load()
    cfg = L.Config(abi_version=L.ABI_VERSION, device=0, n_cells=10, n_layers=1, n_columns=4, maxdelta=5e4, courant=0.5,
                   dt0=60.0, dtmin=1.0, dtmax=3600.0, lite_miniters=2, lite_maxiters=1000, lite_tol=1e-3)
    h = C.c_void_p()
    rc = lib.cgb_create(C.byref(cfg), C.byref(h))
    assert rc == L.ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.cgb_last_error(None)
    with pytest.raises(L.CgbError):
        import common
        cg.Engine(common.fourier_tile())


def test_create_rejects_bad_config():
    lib = L.load()
    h = C.c_void_p()
    cfg = L.Config(abi_version=99, n_cells=10, n_layers=1, n_columns=1)
    assert lib.cgb_create(C.byref(cfg), C.byref(h)) == L.ERR_INVALID
    cfg = L.Config(abi_version=L.ABI_VERSION, n_cells=1, n_layers=1, n_columns=1)
    assert lib.cgb_create(C.byref(cfg), C.byref(h)) == L.ERR_INVALID
    assert lib.cgb_create(None, C.byref(h)) == L.ERR_INVALID


def test_product_path_does_not_import_oracle():
    pkg = os.path.join(ROOT, "cryogrid.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "cg_oracle" not in src and "import oracle" not in src and "from oracle" not in src, f
