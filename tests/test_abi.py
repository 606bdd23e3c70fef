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
    for header, table in (("vilba.h", capi.EXPORTED_SYMBOLS), ("vilba_diag.h", capi.DIAG_SYMBOLS)):
        hdr = open(os.path.join(ROOT, "include", header)).read()
        hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)  # declarations only, not prose in comments
        declared = set(re.findall(r"\b(vilba_[a-z_]+)\s*\(", hdr))
        assert declared == set(table), declared ^ set(table)
        for sym in declared:
            assert getattr(built, sym) is not None
    assert sorted(os.listdir(os.path.join(ROOT, "include"))) == ["vilba.h", "vilba_diag.h"]


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


def test_blob_and_shard_entry_points_from_plain_c(tmp_path, built):
    """The host-only entry points called the way a C / C++ maintainer would: a C program linked against libvilba.so
    serialises a window it builds itself, reads it back as a zero-copy view, and asks for the point shards."""
    src = tmp_path / "blob.c"
    src.write_text(r'''
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "vilba.h"
int main(void) {
    enum { K = 3, NI = 2, P = 4, E = 7 };
    double kf[K * VILBA_NS_DOUBLES] = {0}, pre[NI * VILBA_PREINT_DOUBLES] = {0}, pts[P * 3];
    unsigned char flags[K] = {VILBA_KF_FIXED | VILBA_KF_HAS_BIAS, VILBA_KF_HAS_BIAS, VILBA_KF_HAS_BIAS};
    int64_t ids[K] = {10, 11, 12};
    int32_t ii[NI] = {0, 1}, jj[NI] = {1, 2}, begin[P + 1] = {0, 2, 4, 5, 7}, okf[E] = {0, 1, 1, 2, 0, 1, 2};
    float uv[2 * E], is2[E];
    for (int i = 0; i < K; ++i) kf[i * VILBA_NS_DOUBLES + 6] = 1.0;
    for (int i = 0; i < P * 3; ++i) pts[i] = 0.5 * i;
    for (int i = 0; i < E; ++i) { uv[2 * i] = 10.f * i; uv[2 * i + 1] = 3.f * i; is2[i] = 1.f; }
    vilba_window w;
    memset(&w, 0, sizeof(w));
    w.n_kf = K; w.n_imu = NI; w.n_pts = P; w.n_obs = E;
    w.kf_state = kf; w.kf_flags = flags; w.kf_id = ids; w.imu_kf_i = ii; w.imu_kf_j = jj; w.imu_preint = pre;
    w.pt_xyz = pts; w.pt_obs_begin = begin; w.obs_kf = okf; w.obs_uv = uv; w.obs_inv_sigma2 = is2;
    w.fx = 458.654; w.fy = 457.296; w.cx = 367.215; w.cy = 248.375;
    w.Rbc[0] = w.Rbc[4] = w.Rbc[8] = 1.0; w.gravity[2] = -9.81;
    size_t n = vilba_window_blob_size(&w), written = 0;
    uint64_t* buf = (uint64_t*)calloc((n + 7) / 8, 8);
    if (!n || vilba_window_serialize(&w, buf, n, &written) != VILBA_OK || written != n) return 1;
    if (vilba_window_serialize(&w, buf, n - 1, &written) == VILBA_OK) return 2;   /* capacity too small */
    vilba_window v;
    if (vilba_window_deserialize(buf, n, &v) != VILBA_OK) return 3;
    if (v.n_kf != K || v.n_obs != E || v.kf_id[2] != 12 || v.obs_kf[6] != 2 || v.pt_xyz[11] != 5.5 || v.obs_uv[13] != 18.f ||
        v.fx != w.fx || v.gravity[2] != -9.81 || memcmp(v.pt_obs_begin, begin, sizeof(begin)) != 0) return 4;
    ((unsigned char*)buf)[300] ^= 1;                                            /* corrupt the payload */
    if (vilba_window_deserialize(buf, n, &v) == VILBA_OK) return 5;
    ((unsigned char*)buf)[300] ^= 1;
    if (vilba_window_deserialize((char*)buf + 4, n - 4, &v) == VILBA_OK) return 6;  /* misaligned / not a blob */
    int32_t p0, p1, total = 0;
    for (int r = 0; r < 3; ++r) {
        if (vilba_shard_points(&w, r, 3, &p0, &p1) != VILBA_OK || p1 < p0) return 7;
        total += p1 - p0;
    }
    if (total != P || vilba_shard_points(&w, 3, 3, &p0, &p1) == VILBA_OK) return 8;
    printf("ok %zu %s\n", n, vilba_version());
    free(buf);
    return 0;
}
''')
    exe = tmp_path / "blob"
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-l:libvilba.so", f"-Wl,-rpath,{libdir}"])
    out = subprocess.check_output([str(exe)]).decode()
    assert out.startswith("ok ") and "vilba" in out
