"""ctypes binding of oracle/_ref/libref_imu.so: the reference's OWN so3 / IMUPreintegrator / NavState / imudata / g2otypes
sources compiled unmodified against oracle/eigen_stub and oracle/g2o_stub (oracle/Makefile target `ref`).  TEST INFRASTRUCTURE: the pin of the
oracle's restatement; only tests/ and the golden-vector generators use it."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_imu.so")
REFERENCE = os.environ.get("VILBA_REFERENCE", "/root/reference")
_dp = C.POINTER(C.c_double)
_lib = None


def available() -> bool:
    """True if the compiled reference is there, or can be built here (the reference tree exists in this container only)."""
    return os.path.exists(LIB_PATH) or os.path.isdir(os.path.join(REFERENCE, "src", "IMU"))


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref", f"REF={REFERENCE}"] + (["-B"] if force else []))
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(LIB_PATH)
        l.ref_preintegrate.argtypes = [C.c_int32] + [_dp] * 6
        l.ref_preintegrate.restype = None
        for name, n in [("ref_so3_exp", 2), ("ref_so3_log", 2), ("ref_so3_mul", 3), ("ref_so3_inverse", 2), ("ref_so3_matrix", 2),
                        ("ref_so3_from_matrix", 2), ("ref_so3_rotate", 3), ("ref_jacobian_r", 2), ("ref_jacobian_r_inv", 2),
                        ("ref_so3_jacobian_r", 2), ("ref_so3_jacobian_r_inv", 2), ("ref_navstate_inc_pvr", 2),
                        ("ref_navstate_inc_bias", 2), ("ref_imu_constants", 1)]:
            f = getattr(l, name)
            f.argtypes = [_dp] * n
            f.restype = None
        l.ref_build_info.restype = C.c_char_p
        l.ref_edge_pvr.argtypes = [C.c_int32] + [_dp] * 13
        l.ref_edge_pvr.restype = None
        l.ref_edge_bias.argtypes = [_dp] * 5
        l.ref_edge_bias.restype = None
        l.ref_edge_mono.argtypes = [_dp] * 7 + [C.POINTER(C.c_int32)]
        l.ref_edge_mono.restype = None
        l.ref_huber.argtypes = [C.c_double, C.c_double, _dp]
        l.ref_huber.restype = None
        for name in ("ref_vertex_pvr_oplus", "ref_vertex_bias_oplus"):
            getattr(l, name).argtypes = [_dp, _dp]
            getattr(l, name).restype = None
        _lib = l
    return _lib


def _a(x, n=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
    assert n is None or a.size == n
    return a


def _d(a):
    return a.ctypes.data_as(_dp)


def _call(name, out_n, *ins):
    out = np.zeros(out_n)
    getattr(lib(), name)(*[_d(_a(i)) for i in ins], _d(out))
    return out


def preintegrate_batch(sample_begin, gyro, acc, dt, bg, ba) -> np.ndarray:
    """IMUPreintegrator::reset + update per sample for every key-frame pair, like KeyFrame::ComputePreInt."""
    sb = np.asarray(sample_begin, dtype=np.int64)
    g, a, t = _a(gyro).reshape(-1, 3), _a(acc).reshape(-1, 3), _a(dt)
    b1, b2 = _a(bg).reshape(-1, 3), _a(ba).reshape(-1, 3)
    out = np.zeros((sb.size - 1, 142))
    for p in range(sb.size - 1):
        s0, s1 = int(sb[p]), int(sb[p + 1])
        gi, ai, ti = np.ascontiguousarray(g[s0:s1]), np.ascontiguousarray(a[s0:s1]), np.ascontiguousarray(t[s0:s1])
        lib().ref_preintegrate(s1 - s0, _d(gi), _d(ai), _d(ti), _d(np.ascontiguousarray(b1[p])), _d(np.ascontiguousarray(b2[p])),
                               _d(out[p]))
    return out


def so3_exp(w): return _call("ref_so3_exp", 4, w)
def so3_log(q): return _call("ref_so3_log", 3, q)
def so3_mul(a, b): return _call("ref_so3_mul", 4, a, b)
def so3_inverse(a): return _call("ref_so3_inverse", 4, a)
def so3_matrix(a): return _call("ref_so3_matrix", 9, a).reshape(3, 3)
def so3_from_matrix(R): return _call("ref_so3_from_matrix", 4, R)
def so3_rotate(a, v): return _call("ref_so3_rotate", 3, a, v)
def jacobian_r(w): return _call("ref_jacobian_r", 9, w).reshape(3, 3)
def jacobian_r_inv(w): return _call("ref_jacobian_r_inv", 9, w).reshape(3, 3)
def so3_jacobian_r(w): return _call("ref_so3_jacobian_r", 9, w).reshape(3, 3)
def so3_jacobian_r_inv(w): return _call("ref_so3_jacobian_r_inv", 9, w).reshape(3, 3)


def inc_pvr(ns, d):
    s = _a(ns, 22).copy()
    lib().ref_navstate_inc_pvr(_d(s), _d(_a(d, 9)))
    return s


def inc_bias(ns, d):
    s = _a(ns, 22).copy()
    lib().ref_navstate_inc_bias(_d(s), _d(_a(d, 6)))
    return s


def imu_constants(): return _call("ref_imu_constants", 4)
def build_info() -> str: return lib().ref_build_info().decode()


# ---- the reference's own factor classes (src/IMU/g2otypes.cpp), see oracle/ref_harness_edges.cpp ----
def edge_pvr(gyro, acc, dt, bg, ba, ns_i, ns_j, ns_bias_i, g):
    """EdgeNavStatePVR with the pre-integration of the given samples as measurement: (err 9, Ji 9x9, Jj 9x9, Jb 9x6)."""
    gy, ac, t = _a(gyro), _a(acc), _a(dt)
    err, Ji, Jj, Jb = np.zeros(9), np.zeros(81), np.zeros(81), np.zeros(54)
    lib().ref_edge_pvr(t.size, _d(gy), _d(ac), _d(t), _d(_a(bg, 3)), _d(_a(ba, 3)), _d(_a(ns_i, 22)), _d(_a(ns_j, 22)),
                       _d(_a(ns_bias_i, 22)), _d(_a(g, 3)), _d(err), _d(Ji), _d(Jj), _d(Jb))
    return err, Ji.reshape(9, 9), Jj.reshape(9, 9), Jb.reshape(9, 6)


def edge_bias(ns_i, ns_j):
    err, Ji, Jj = np.zeros(6), np.zeros(36), np.zeros(36)
    lib().ref_edge_bias(_d(_a(ns_i, 22)), _d(_a(ns_j, 22)), _d(err), _d(Ji), _d(Jj))
    return err, Ji.reshape(6, 6), Jj.reshape(6, 6)


def edge_mono(ns, pw, calib, uv):
    """EdgeNavStatePVRPointXYZ: (err 2, Jpoint 2x3, Jpvr 2x9, isDepthPositive)."""
    err, Jp, Jn = np.zeros(2), np.zeros(6), np.zeros(18)
    dp = C.c_int32(0)
    lib().ref_edge_mono(_d(_a(ns, 22)), _d(_a(pw, 3)), _d(_a(calib, 16)), _d(_a(uv, 2)), _d(err), _d(Jp), _d(Jn), C.byref(dp))
    return err, Jp.reshape(2, 3), Jn.reshape(2, 9), bool(dp.value)


def vertex_pvr_oplus(ns, d):
    s = _a(ns, 22).copy()
    lib().ref_vertex_pvr_oplus(_d(s), _d(_a(d, 9)))
    return s


def vertex_bias_oplus(ns, d):
    s = _a(ns, 22).copy()
    lib().ref_vertex_bias_oplus(_d(s), _d(_a(d, 6)))
    return s


# ---- g2o's own Levenberg-Marquardt step and robust kernel, see oracle/ref_harness_lm.cpp ----
def huber(e2: float, delta: float) -> np.ndarray:
    """RobustKernelHuber::robustify after setDelta(delta): rho, rho', rho''."""
    rho = np.zeros(3)
    lib().ref_huber(float(e2), float(delta), _d(rho))
    return rho


def lm_local_ba(win, params=None, stop_flag=None):
    """oracle_local_ba with the reference's OptimizationAlgorithmLevenberg::solve() as the LM step."""
    from mc_slam_b200.capi import CResult, CWindow, Params, Result, default_params
    f = lib().ref_lm_local_ba
    f.argtypes = [C.POINTER(CWindow), C.POINTER(Params), C.POINTER(CResult), C.POINTER(C.c_uint8)]
    f.restype = C.c_int
    res = Result.alloc(win)
    cw, cr = win.as_c(), res.as_c()
    p = params or default_params()
    sf = stop_flag.ctypes.data_as(C.POINTER(C.c_uint8)) if stop_flag is not None else None
    st = f(C.byref(cw), C.byref(p), C.byref(cr), sf)
    res.take(cr)
    if st < 0:
        raise RuntimeError(f"ref_lm_local_ba failed with status {st}")
    return res
