"""CPU-side checks of the C-ABI boundary: the product library loads and exports every symbol that
include/vilba.h declares, the ctypes mirror matches the C struct layout, and without a CUDA device
the product refuses to run (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from mc_slam_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return capi.load_library()


def test_header_symbols_are_exported(built):
    hdr = open(os.path.join(ROOT, "include", "vilba.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)  # declarations only, not prose in comments
    declared = set(re.findall(r"\b(vilba_[a-z_]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTED_SYMBOLS), declared ^ set(capi.EXPORTED_SYMBOLS)
    for sym in declared:
        assert getattr(built, sym) is not None


def test_struct_layout_matches_c(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "vilba.h"\n'
        "int main(){printf(\"%zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(vilba_params), sizeof(vilba_window),"
        " sizeof(vilba_iter_record), sizeof(vilba_result), sizeof(vilba_stats), offsetof(vilba_window, fx),"
        " offsetof(vilba_result, trace), offsetof(vilba_result, solve_ms));return 0;}\n"
    )
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(capi.Params), C.sizeof(capi.CWindow), C.sizeof(capi.IterRecord), C.sizeof(capi.CResult),
            C.sizeof(capi.Stats), capi.CWindow.fx.offset, capi.CResult.trace.offset, capi.CResult.solve_ms.offset]
    assert got == want


def test_default_params_match_reference_literals(built):
    p = capi.Params()
    built.vilba_default_params(C.byref(p))
    q = capi.default_params()
    for name, _ in capi.Params._fields_:
        assert getattr(p, name) == getattr(q, name), name
    assert p.huber_mono == float(np.float32(np.sqrt(5.991)))  # float-rounded delta (Optimizer.cpp:2580)
    assert p.huber_mono ** 2 != 5.991 and p.chi2_gate == 5.991


def test_no_cpu_fallback_without_cuda(built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from mc_slam_b200 import api

    with pytest.raises(api.VilbaError):
        api.Context(0)
    assert not built.vilba_create(0, None)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mc_slam_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "vilba_oracle" not in txt and "oracle/" not in txt, f
