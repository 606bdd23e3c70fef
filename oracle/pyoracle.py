"""ctypes binding of the CPU oracle (oracle/libvilba_oracle.so).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package (mc_slam_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from mc_slam_b200.capi import CResult, CWindow, Params, Result, Window, default_params, PREINT_DOUBLES

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvilba_oracle.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (g++ only, no external dependency)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        l = C.CDLL(LIB_PATH)
        l.oracle_local_ba.argtypes = [C.POINTER(CWindow), C.POINTER(Params), C.POINTER(CResult), C.POINTER(C.c_uint8)]
        l.oracle_local_ba.restype = C.c_int
        l.oracle_preintegrate_batch.argtypes = [C.POINTER(Params), C.c_int32, _ip, _dp, _dp, _dp, _dp, _dp, _dp]
        l.oracle_preintegrate_batch.restype = C.c_int
        l.oracle_debug_system.argtypes = [C.POINTER(CWindow), C.POINTER(Params), C.c_int, C.c_double] + [_dp] * 10
        l.oracle_debug_system.restype = C.c_int
        l.oracle_build_info.restype = C.c_char_p
        for name, n in [
            ("oracle_so3_exp", 2), ("oracle_so3_log", 2), ("oracle_quat_to_matrix", 2), ("oracle_matrix_to_quat", 2),
            ("oracle_jacobian_r", 2), ("oracle_jacobian_r_inv", 2), ("oracle_inverse9", 2),
            ("oracle_navstate_oplus_pvr", 2), ("oracle_navstate_oplus_bias", 2), ("oracle_bias_edge", 3),
            ("oracle_pvr_edge", 9),
        ]:
            f = getattr(l, name)
            f.argtypes = [_dp] * n
            f.restype = None
        l.oracle_mono_edge.argtypes = [_dp] * 7 + [C.POINTER(C.c_int)]
        l.oracle_mono_edge.restype = None
        _lib = l
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _arr(x, n=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
    if n is not None:
        assert a.size == n, (a.size, n)
    return a


def local_ba(win: Window, params: Optional[Params] = None, stop_flag: Optional[np.ndarray] = None) -> Result:
    res = Result.alloc(win)
    cw, cr = win.as_c(), res.as_c()
    p = params or default_params()
    sf = stop_flag.ctypes.data_as(C.POINTER(C.c_uint8)) if stop_flag is not None else None
    st = lib().oracle_local_ba(C.byref(cw), C.byref(p), C.byref(cr), sf)
    res.take(cr)
    if st < 0:
        raise RuntimeError(f"oracle_local_ba failed with status {st}")
    return res


def preintegrate_batch(sample_begin, gyro, acc, dt, bg, ba, params: Optional[Params] = None) -> np.ndarray:
    sb = np.ascontiguousarray(sample_begin, dtype=np.int32)
    n = sb.size - 1
    g, a, t = _arr(gyro), _arr(acc), _arr(dt)
    b1, b2 = _arr(bg, 3 * n), _arr(ba, 3 * n)
    out = np.zeros((n, PREINT_DOUBLES), np.float64)
    p = params or default_params()
    st = lib().oracle_preintegrate_batch(C.byref(p), n, sb.ctypes.data_as(_ip), _d(g), _d(a), _d(t), _d(b1), _d(b2), _d(out))
    if st != 0:
        raise RuntimeError(f"oracle_preintegrate_batch failed with status {st}")
    return out


def so3_exp(w):
    q = np.zeros(4)
    lib().oracle_so3_exp(_d(_arr(w, 3)), _d(q))
    return q


def so3_log(q):
    w = np.zeros(3)
    lib().oracle_so3_log(_d(_arr(q, 4)), _d(w))
    return w


def quat_to_matrix(q):
    R = np.zeros(9)
    lib().oracle_quat_to_matrix(_d(_arr(q, 4)), _d(R))
    return R.reshape(3, 3)


def matrix_to_quat(R):
    q = np.zeros(4)
    lib().oracle_matrix_to_quat(_d(_arr(R, 9)), _d(q))
    return q


def jacobian_r(w):
    J = np.zeros(9)
    lib().oracle_jacobian_r(_d(_arr(w, 3)), _d(J))
    return J.reshape(3, 3)


def jacobian_r_inv(w):
    J = np.zeros(9)
    lib().oracle_jacobian_r_inv(_d(_arr(w, 3)), _d(J))
    return J.reshape(3, 3)


def inverse9(A):
    o = np.zeros(81)
    lib().oracle_inverse9(_d(_arr(A, 81)), _d(o))
    return o.reshape(9, 9)


def oplus_pvr(ns, d):
    s = _arr(ns, 22).copy()
    lib().oracle_navstate_oplus_pvr(_d(s), _d(_arr(d, 9)))
    return s


def oplus_bias(ns, d):
    s = _arr(ns, 22).copy()
    lib().oracle_navstate_oplus_bias(_d(s), _d(_arr(d, 6)))
    return s


def calib_vec(win: Window) -> np.ndarray:
    return np.concatenate([[win.fx, win.fy, win.cx, win.cy], win.Rbc.reshape(-1), win.Pbc.reshape(-1)]).astype(np.float64)


def mono_edge(ns, pw, calib, uv):
    err, Jp, Jn = np.zeros(2), np.zeros(6), np.zeros(18)
    dp = C.c_int(0)
    lib().oracle_mono_edge(_d(_arr(ns, 22)), _d(_arr(pw, 3)), _d(_arr(calib, 16)), _d(_arr(uv, 2)), _d(err), _d(Jp),
                           _d(Jn), C.byref(dp))
    return err, Jp.reshape(2, 3), Jn.reshape(2, 9), bool(dp.value)


def pvr_edge(ns_i, ns_j, ns_bias_i, preint, g):
    err, Ji, Jj, Jb = np.zeros(9), np.zeros(81), np.zeros(81), np.zeros(54)
    lib().oracle_pvr_edge(_d(_arr(ns_i, 22)), _d(_arr(ns_j, 22)), _d(_arr(ns_bias_i, 22)), _d(_arr(preint, 142)),
                          _d(_arr(g, 3)), _d(err), _d(Ji), _d(Jj), _d(Jb))
    return err, Ji.reshape(9, 9), Jj.reshape(9, 9), Jb.reshape(9, 6)


def bias_edge(ns_i, ns_j):
    err = np.zeros(6)
    lib().oracle_bias_edge(_d(_arr(ns_i, 22)), _d(_arr(ns_j, 22)), _d(err))
    return err


def huber(e2: float, delta: float) -> np.ndarray:
    rho = np.zeros(3)
    f = lib().oracle_huber
    f.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_double)]
    f.restype = None
    f(float(e2), float(delta), rho.ctypes.data_as(C.POINTER(C.c_double)))
    return rho


def debug_system(win: Window, lam: float, robust_mono: bool = True, params: Optional[Params] = None) -> dict:
    n = 15 * win.n_free
    P, E = win.n_pts, win.n_obs
    o = dict(
        Hpp=np.zeros((n, n)), bp=np.zeros(n), Hll=np.zeros((P, 3, 3)), bl=np.zeros((P, 3)), Hpl=np.zeros((E, 6, 3)),
        S=np.zeros((n, n)), bs=np.zeros(n), x=np.zeros(n + 3 * P), chi2=np.zeros(1), obs_chi2=np.zeros(E),
    )
    cw = win.as_c()
    p = params or default_params()
    r = lib().oracle_debug_system(C.byref(cw), C.byref(p), int(robust_mono), float(lam), *[_d(o[k]) for k in
                                  ("Hpp", "bp", "Hll", "bl", "Hpl", "S", "bs", "x", "chi2", "obs_chi2")])
    o["n"] = r
    return o


def build_info() -> str:
    return lib().oracle_build_info().decode()
