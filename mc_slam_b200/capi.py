"""ctypes mirror of include/vilba.h (the C-ABI drop-in boundary).

Nothing here computes: it lays numpy arrays out as the plain structs the C ABI takes and loads
``libvilba.so`` (CUDA kernels + host C++ controller).  If that library cannot be loaded, or no CUDA
device is usable, the product API raises -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

NS_DOUBLES = 22
PREINT_DOUBLES = 142
KF_FIXED = 1
KF_HAS_BIAS = 2
MAX_TRACE = 64

OK = 0
ABORTED = 1

_c_double_p = C.POINTER(C.c_double)
_c_float_p = C.POINTER(C.c_float)
_c_int32_p = C.POINTER(C.c_int32)
_c_int64_p = C.POINTER(C.c_int64)
_c_uint8_p = C.POINTER(C.c_uint8)


class Params(C.Structure):
    _fields_ = [
        ("iters_stage1", C.c_int32),
        ("iters_stage2", C.c_int32),
        ("max_trials", C.c_int32),
        ("mode", C.c_int32),
        ("huber_mono", C.c_double),
        ("huber_pvr", C.c_double),
        ("huber_bias", C.c_double),
        ("chi2_gate", C.c_double),
        ("lm_tau", C.c_double),
        ("lm_good_lo", C.c_double),
        ("lm_good_hi", C.c_double),
        ("gyr_bias_rw2", C.c_double),
        ("acc_bias_rw2", C.c_double),
        ("gyr_meas_cov", C.c_double),
        ("acc_meas_cov", C.c_double),
    ]


MODE_SINGLE_STAGE = 1  # VILBA_MODE_SINGLE_STAGE
MODE_MONO_NOT_ROBUST = 2  # VILBA_MODE_MONO_NOT_ROBUST


def global_ba_params(n_iterations: int, robust: bool, base: Optional["Params"] = None) -> "Params":
    """Host-side restatement of vilba_global_ba_params: the settings of Optimizer::GlobalBundleAdjustmentNavState
    (src/Optimizer.cpp:1438-1439,1541,1590-1595,1624).  The tests hand these to the oracle."""
    p = Params()
    C.memmove(C.byref(p), C.byref(base if base is not None else default_params()), C.sizeof(Params))
    p.mode = MODE_SINGLE_STAGE | (0 if robust else MODE_MONO_NOT_ROBUST)
    p.iters_stage1 = int(n_iterations)
    p.iters_stage2 = 0
    if robust:
        p.huber_pvr = float(np.float32(np.sqrt(21.666)))
        p.huber_bias = float(np.float32(np.sqrt(16.812)))
        p.huber_mono = float(np.float32(np.sqrt(5.99)))
    else:
        p.huber_pvr = p.huber_bias = float("inf")
    return p


def default_params() -> Params:
    """The literals of the reference (src/Optimizer.cpp:2487-2488,2580,2648,2667,2676; imudata.cpp:25-31)."""
    p = Params()
    p.iters_stage1 = 5
    p.iters_stage2 = 10
    p.max_trials = 10
    p.huber_mono = float(np.float32(np.sqrt(5.991)))
    p.huber_pvr = float(np.float32(np.sqrt(100 * 21.666)))
    p.huber_bias = float(np.float32(np.sqrt(100 * 16.812)))
    p.chi2_gate = 5.991
    p.lm_tau = 1e-5
    p.lm_good_lo = 1.0 / 3.0
    p.lm_good_hi = 2.0 / 3.0
    p.gyr_bias_rw2 = 2.0e-5 * 2.0e-5
    p.acc_bias_rw2 = 5.0e-3 * 5.0e-3
    p.gyr_meas_cov = 1.7e-4 * 1.7e-4 / 0.005
    p.acc_meas_cov = 2.0e-3 * 2.0e-3 / 0.005 * 100
    return p


class CWindow(C.Structure):
    _fields_ = [
        ("n_kf", C.c_int32),
        ("n_imu", C.c_int32),
        ("n_pts", C.c_int32),
        ("n_obs", C.c_int32),
        ("kf_state", _c_double_p),
        ("kf_flags", _c_uint8_p),
        ("kf_id", _c_int64_p),
        ("imu_kf_i", _c_int32_p),
        ("imu_kf_j", _c_int32_p),
        ("imu_preint", _c_double_p),
        ("pt_xyz", _c_double_p),
        ("pt_obs_begin", _c_int32_p),
        ("obs_kf", _c_int32_p),
        ("obs_uv", _c_float_p),
        ("obs_inv_sigma2", _c_float_p),
        ("fx", C.c_double),
        ("fy", C.c_double),
        ("cx", C.c_double),
        ("cy", C.c_double),
        ("Rbc", C.c_double * 9),
        ("Pbc", C.c_double * 3),
        ("gravity", C.c_double * 3),
    ]


class IterRecord(C.Structure):
    _fields_ = [
        ("stage", C.c_int32),
        ("iteration", C.c_int32),
        ("trials", C.c_int32),
        ("result", C.c_int32),
        ("n_active_edges", C.c_int32),
        ("accepted", C.c_int32),
        ("chi2_initial", C.c_double),
        ("chi2_final", C.c_double),
        ("lambda_", C.c_double),
        ("lambda_first_trial", C.c_double),
    ]


class CResult(C.Structure):
    _fields_ = [
        ("kf_state", _c_double_p),
        ("pt_xyz", _c_double_p),
        ("obs_outlier", _c_uint8_p),
        ("obs_chi2", _c_double_p),
        ("status", C.c_int32),
        ("stage2_ran", C.c_int32),
        ("n_trace", C.c_int32),
        ("n_outliers_stage1", C.c_int32),
        ("trace", IterRecord * MAX_TRACE),
        ("solve_ms", C.c_double),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_int64),
        ("lm_iterations", C.c_int64),
        ("lm_trials", C.c_int64),
        ("edges_linearized", C.c_int64),
        ("linearize_ms", C.c_double),
        ("linearize_launches", C.c_int64),
        ("schur_ms", C.c_double),
        ("schur_launches", C.c_int64),
        ("solve_ms", C.c_double),
        ("solve_launches", C.c_int64),
        ("update_ms", C.c_double),
        ("update_launches", C.c_int64),
        ("preint_ms", C.c_double),
        ("preint_launches", C.c_int64),
    ]


def _ptr(a: Optional[np.ndarray], ctype):
    if a is None:
        return C.cast(None, C.POINTER(ctype))
    return a.ctypes.data_as(C.POINTER(ctype))


@dataclass
class Window:
    """One local-BA window as numpy struct-of-arrays (layout: include/vilba.h vilba_window)."""

    kf_state: np.ndarray  # (K,22) f64
    kf_flags: np.ndarray  # (K,) u8
    kf_id: np.ndarray  # (K,) i64
    imu_kf_i: np.ndarray  # (NI,) i32
    imu_kf_j: np.ndarray  # (NI,) i32
    imu_preint: np.ndarray  # (NI,142) f64
    pt_xyz: np.ndarray  # (P,3) f64 (float32-valued)
    pt_obs_begin: np.ndarray  # (P+1,) i32
    obs_kf: np.ndarray  # (E,) i32
    obs_uv: np.ndarray  # (E,2) f32
    obs_inv_sigma2: np.ndarray  # (E,) f32
    fx: float
    fy: float
    cx: float
    cy: float
    Rbc: np.ndarray  # (3,3)
    Pbc: np.ndarray  # (3,)
    gravity: np.ndarray  # (3,)
    truth: dict = field(default_factory=dict)  # generator ground truth (not part of the ABI)

    def __post_init__(self):
        self.kf_state = np.ascontiguousarray(self.kf_state, dtype=np.float64).reshape(-1, NS_DOUBLES)
        self.kf_flags = np.ascontiguousarray(self.kf_flags, dtype=np.uint8)
        self.kf_id = np.ascontiguousarray(self.kf_id, dtype=np.int64)
        self.imu_kf_i = np.ascontiguousarray(self.imu_kf_i, dtype=np.int32)
        self.imu_kf_j = np.ascontiguousarray(self.imu_kf_j, dtype=np.int32)
        self.imu_preint = np.ascontiguousarray(self.imu_preint, dtype=np.float64).reshape(-1, PREINT_DOUBLES)
        self.pt_xyz = np.ascontiguousarray(self.pt_xyz, dtype=np.float64).reshape(-1, 3)
        self.pt_obs_begin = np.ascontiguousarray(self.pt_obs_begin, dtype=np.int32)
        self.obs_kf = np.ascontiguousarray(self.obs_kf, dtype=np.int32)
        self.obs_uv = np.ascontiguousarray(self.obs_uv, dtype=np.float32).reshape(-1, 2)
        self.obs_inv_sigma2 = np.ascontiguousarray(self.obs_inv_sigma2, dtype=np.float32)
        self.Rbc = np.ascontiguousarray(self.Rbc, dtype=np.float64).reshape(3, 3)
        self.Pbc = np.ascontiguousarray(self.Pbc, dtype=np.float64).reshape(3)
        self.gravity = np.ascontiguousarray(self.gravity, dtype=np.float64).reshape(3)

    @property
    def n_kf(self) -> int:
        return int(self.kf_state.shape[0])

    @property
    def n_imu(self) -> int:
        return int(self.imu_kf_i.shape[0])

    @property
    def n_pts(self) -> int:
        return int(self.pt_xyz.shape[0])

    @property
    def n_obs(self) -> int:
        return int(self.obs_kf.shape[0])

    @property
    def n_free(self) -> int:
        return int(np.count_nonzero((self.kf_flags & KF_FIXED) == 0))

    # ---- wire / on-disk format (include/vilba.h: vilba_window_serialize / _deserialize) ----------------
    def to_bytes(self) -> bytes:
        lib = load_library()
        cw = self.as_c()
        n = lib.vilba_window_blob_size(C.byref(cw))
        if n == 0:
            raise ValueError("invalid window")
        buf = np.zeros(n // 8, np.uint64)  # 8-byte aligned
        written = C.c_size_t(0)
        st = lib.vilba_window_serialize(C.byref(cw), buf.ctypes.data_as(C.c_void_p), n, C.byref(written))
        if st != 0 or written.value != n:
            raise ValueError(f"vilba_window_serialize failed ({st})")
        return buf.tobytes()

    @staticmethod
    def from_bytes(blob: bytes) -> "Window":
        lib = load_library()
        n = len(blob)
        buf = np.zeros((n + 7) // 8, np.uint64)
        C.memmove(buf.ctypes.data, blob, n)
        cw = CWindow()
        st = lib.vilba_window_deserialize(buf.ctypes.data_as(C.c_void_p), n, C.byref(cw))
        if st != 0:
            raise ValueError(f"vilba_window_deserialize failed ({st}): not a window blob, truncated or corrupt")
        K, NI, P, E = cw.n_kf, cw.n_imu, cw.n_pts, cw.n_obs

        def arr(ptr, count, dtype):
            if count == 0:
                return np.zeros(0, dtype)
            return np.ctypeslib.as_array(ptr, shape=(count,)).astype(dtype, copy=True)

        return Window(
            kf_state=arr(cw.kf_state, NS_DOUBLES * K, np.float64), kf_flags=arr(cw.kf_flags, K, np.uint8),
            kf_id=arr(cw.kf_id, K, np.int64), imu_kf_i=arr(cw.imu_kf_i, NI, np.int32), imu_kf_j=arr(cw.imu_kf_j, NI, np.int32),
            imu_preint=arr(cw.imu_preint, PREINT_DOUBLES * NI, np.float64), pt_xyz=arr(cw.pt_xyz, 3 * P, np.float64),
            pt_obs_begin=arr(cw.pt_obs_begin, P + 1, np.int32), obs_kf=arr(cw.obs_kf, E, np.int32),
            obs_uv=arr(cw.obs_uv, 2 * E, np.float32), obs_inv_sigma2=arr(cw.obs_inv_sigma2, E, np.float32),
            fx=cw.fx, fy=cw.fy, cx=cw.cx, cy=cw.cy, Rbc=np.array(list(cw.Rbc)), Pbc=np.array(list(cw.Pbc)),
            gravity=np.array(list(cw.gravity)))

    def as_c(self) -> CWindow:
        w = CWindow()
        w.n_kf, w.n_imu, w.n_pts, w.n_obs = self.n_kf, self.n_imu, self.n_pts, self.n_obs
        w.kf_state = _ptr(self.kf_state, C.c_double)
        w.kf_flags = _ptr(self.kf_flags, C.c_uint8)
        w.kf_id = _ptr(self.kf_id, C.c_int64)
        w.imu_kf_i = _ptr(self.imu_kf_i, C.c_int32)
        w.imu_kf_j = _ptr(self.imu_kf_j, C.c_int32)
        w.imu_preint = _ptr(self.imu_preint, C.c_double)
        w.pt_xyz = _ptr(self.pt_xyz, C.c_double)
        w.pt_obs_begin = _ptr(self.pt_obs_begin, C.c_int32)
        w.obs_kf = _ptr(self.obs_kf, C.c_int32)
        w.obs_uv = _ptr(self.obs_uv, C.c_float)
        w.obs_inv_sigma2 = _ptr(self.obs_inv_sigma2, C.c_float)
        w.fx, w.fy, w.cx, w.cy = self.fx, self.fy, self.cx, self.cy
        for i in range(9):
            w.Rbc[i] = float(self.Rbc.reshape(-1)[i])
        for i in range(3):
            w.Pbc[i] = float(self.Pbc[i])
            w.gravity[i] = float(self.gravity[i])
        return w


@dataclass
class Result:
    """Outputs of one local-BA call (layout: include/vilba.h vilba_result)."""

    kf_state: np.ndarray
    pt_xyz: np.ndarray
    obs_outlier: np.ndarray
    obs_chi2: Optional[np.ndarray]
    status: int = 0
    stage2_ran: int = 0
    n_outliers_stage1: int = 0
    trace: List[dict] = field(default_factory=list)
    solve_ms: float = 0.0

    @staticmethod
    def alloc(win: Window, chi2: bool = True) -> "Result":
        """`chi2=False`: no per-edge chi2 array (the reference's function does not hand one back either); the library
        then leaves that part of the results on the device."""
        return Result(
            kf_state=np.zeros((win.n_kf, NS_DOUBLES), np.float64),
            pt_xyz=np.zeros((win.n_pts, 3), np.float64),
            obs_outlier=np.zeros((win.n_obs,), np.uint8),
            obs_chi2=np.zeros((win.n_obs,), np.float64) if chi2 else None,
        )

    def as_c(self) -> CResult:
        r = CResult()
        r.kf_state = _ptr(self.kf_state, C.c_double)
        r.pt_xyz = _ptr(self.pt_xyz, C.c_double)
        r.obs_outlier = _ptr(self.obs_outlier, C.c_uint8)
        r.obs_chi2 = _ptr(self.obs_chi2, C.c_double)
        return r

    def take(self, r: CResult) -> "Result":
        self.status = int(r.status)
        self.stage2_ran = int(r.stage2_ran)
        self.n_outliers_stage1 = int(r.n_outliers_stage1)
        self.solve_ms = float(r.solve_ms)
        self.trace = []
        for i in range(int(r.n_trace)):
            t = r.trace[i]
            self.trace.append(
                dict(
                    stage=int(t.stage),
                    iteration=int(t.iteration),
                    trials=int(t.trials),
                    result=int(t.result),
                    n_active_edges=int(t.n_active_edges),
                    accepted=int(t.accepted),
                    chi2_initial=float(t.chi2_initial),
                    chi2_final=float(t.chi2_final),
                    lambda_=float(t.lambda_),
                    lambda_first_trial=float(t.lambda_first_trial),
                )
            )
        return self


# ------------------------------------------------------------------------------------------------
# product library
# ------------------------------------------------------------------------------------------------
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_REPO, "mc_slam_b200", "libvilba.so")

EXPORTED_SYMBOLS = [  # every symbol include/vilba.h declares
    "vilba_default_params",
    "vilba_create",
    "vilba_destroy",
    "vilba_last_error",
    "vilba_version",
    "vilba_local_ba",
    "vilba_global_ba",
    "vilba_global_ba_params",
    "vilba_local_ba_batch",
    "vilba_window_upload",
    "vilba_window_solve_resident",
    "vilba_window_download",
    "vilba_max_batch",
    "vilba_batch_upload",
    "vilba_batch_groups",
    "vilba_batch_solve_resident",
    "vilba_batch_download",
    "vilba_comm_unique_id",
    "vilba_comm_init",
    "vilba_shard_points",
    "vilba_window_blob_size",
    "vilba_window_serialize",
    "vilba_window_deserialize",
    "vilba_preintegrate_batch",
    "vilba_preintegrate_batch_dev",
    "vilba_get_stats",
    "vilba_reset_stats",
    "vilba_set_profiling",
]
DIAG_SYMBOLS = [
    "vilba_diag_first_trial",  # every symbol include/vilba_diag.h declares (diagnostics, not the reference boundary)
    "vilba_diag_dense_solve",
    "vilba_diag_dense_supported",
]

_lib = None


def load_library(path: Optional[str] = None) -> C.CDLL:
    """dlopen libvilba.so and declare the prototypes.  Raises if it is missing (no fallback)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the VI local-BA path has no CPU fallback)"
        )
    lib = C.CDLL(p)
    lib.vilba_default_params.argtypes = [C.POINTER(Params)]
    lib.vilba_default_params.restype = None
    lib.vilba_create.argtypes = [C.c_int, C.POINTER(Params)]
    lib.vilba_create.restype = C.c_void_p
    lib.vilba_destroy.argtypes = [C.c_void_p]
    lib.vilba_destroy.restype = None
    lib.vilba_last_error.argtypes = [C.c_void_p]
    lib.vilba_last_error.restype = C.c_char_p
    lib.vilba_version.argtypes = []
    lib.vilba_version.restype = C.c_char_p
    lib.vilba_local_ba.argtypes = [C.c_void_p, C.POINTER(CWindow), C.POINTER(CResult), _c_uint8_p]
    lib.vilba_local_ba.restype = C.c_int
    lib.vilba_global_ba.argtypes = [C.c_void_p, C.POINTER(CWindow), C.c_int32, C.c_int32, C.POINTER(CResult), _c_uint8_p]
    lib.vilba_global_ba.restype = C.c_int
    lib.vilba_global_ba_params.argtypes = [C.POINTER(Params), C.c_int32, C.c_int32, C.POINTER(Params)]
    lib.vilba_global_ba_params.restype = None
    lib.vilba_local_ba_batch.argtypes = [C.c_void_p, C.c_int32, C.POINTER(CWindow), C.POINTER(CResult)]
    lib.vilba_local_ba_batch.restype = C.c_int
    lib.vilba_window_upload.argtypes = [C.c_void_p, C.POINTER(CWindow)]
    lib.vilba_window_upload.restype = C.c_int
    lib.vilba_window_solve_resident.argtypes = [C.c_void_p, C.POINTER(CResult)]
    lib.vilba_window_solve_resident.restype = C.c_int
    lib.vilba_window_download.argtypes = [C.c_void_p, C.POINTER(CResult)]
    lib.vilba_window_download.restype = C.c_int
    lib.vilba_max_batch.argtypes = []
    lib.vilba_max_batch.restype = C.c_int
    lib.vilba_batch_groups.argtypes = [C.c_void_p]
    lib.vilba_batch_groups.restype = C.c_int
    lib.vilba_batch_upload.argtypes = [C.c_void_p, C.c_int32, C.POINTER(CWindow)]
    lib.vilba_batch_upload.restype = C.c_int
    lib.vilba_batch_solve_resident.argtypes = [C.c_void_p, C.c_int32, C.POINTER(CResult)]
    lib.vilba_batch_solve_resident.restype = C.c_int
    lib.vilba_batch_download.argtypes = [C.c_void_p, C.c_int32, C.POINTER(CResult)]
    lib.vilba_batch_download.restype = C.c_int
    lib.vilba_comm_unique_id.argtypes = [C.c_void_p]
    lib.vilba_comm_unique_id.restype = C.c_int
    lib.vilba_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
    lib.vilba_comm_init.restype = C.c_int
    lib.vilba_shard_points.argtypes = [C.POINTER(CWindow), C.c_int32, C.c_int32, _c_int32_p, _c_int32_p]
    lib.vilba_shard_points.restype = C.c_int
    lib.vilba_window_blob_size.argtypes = [C.POINTER(CWindow)]
    lib.vilba_window_blob_size.restype = C.c_size_t
    lib.vilba_window_serialize.argtypes = [C.POINTER(CWindow), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.vilba_window_serialize.restype = C.c_int
    lib.vilba_window_deserialize.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(CWindow)]
    lib.vilba_window_deserialize.restype = C.c_int
    lib.vilba_preintegrate_batch.argtypes = [
        C.c_void_p, C.c_int32, _c_int32_p, _c_double_p, _c_double_p, _c_double_p, _c_double_p, _c_double_p,
        _c_double_p,
    ]
    lib.vilba_preintegrate_batch.restype = C.c_int
    lib.vilba_preintegrate_batch_dev.argtypes = [
        C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
        C.c_void_p, C.c_void_p,
    ]
    lib.vilba_preintegrate_batch_dev.restype = C.c_int
    lib.vilba_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    lib.vilba_get_stats.restype = None
    lib.vilba_reset_stats.argtypes = [C.c_void_p]
    lib.vilba_reset_stats.restype = None
    lib.vilba_set_profiling.argtypes = [C.c_void_p, C.c_int]
    lib.vilba_set_profiling.restype = None
    lib.vilba_diag_dense_solve.argtypes = [
        C.c_int32, C.c_int32, _c_double_p, _c_double_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _c_double_p,
        _c_int32_p, _c_double_p,
    ]
    lib.vilba_diag_dense_solve.restype = C.c_int
    lib.vilba_diag_first_trial.argtypes = [C.c_void_p, C.POINTER(CWindow)] + [_c_double_p] * 8
    lib.vilba_diag_first_trial.restype = C.c_int
    lib.vilba_diag_dense_supported.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    lib.vilba_diag_dense_supported.restype = C.c_int
    if path is None:
        _lib = lib
    return lib


def diag_dense_solve(S, b, variant=1, cluster=8, n_windows=1, reps=1, device=0):
    """Solve S x = b with one of the library's reduced-system kernels (include/vilba_diag.h).
    Returns (x, fail_flag, average microseconds per launch)."""
    import numpy as np

    lib = load_library()
    S = np.ascontiguousarray(S, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    n = S.shape[0]
    x = np.zeros(n)
    fail = C.c_int32(0)
    us = C.c_double(0.0)
    st = lib.vilba_diag_dense_solve(
        device, n, S.ctypes.data_as(_c_double_p), b.ctypes.data_as(_c_double_p), variant, cluster, n_windows, reps,
        x.ctypes.data_as(_c_double_p), C.byref(fail), C.byref(us))
    if st != 0:
        raise RuntimeError(f"vilba_diag_dense_solve(n={n}, variant={variant}, cluster={cluster}) returned {st}")
    return x, int(fail.value), float(us.value)
