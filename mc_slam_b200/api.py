"""Python face of the C-ABI (include/vilba.h): thin ctypes calls into libvilba.so.

Every call here ends in the CUDA kernels of mc_slam_b200/csrc; if the library or a CUDA device is
missing the constructor raises (there is no CPU path in the product).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import capi
from .capi import PREINT_DOUBLES, CResult, CWindow, Params, Result, Stats, Window, default_params


class VilbaError(RuntimeError):
    pass


class Context:
    """vilba_ctx: owns the CUDA stream, device arenas and the LM controller state of one caller thread."""

    def __init__(self, device: int = 0, params: Optional[Params] = None):
        self._lib = capi.load_library()
        self.params = params or default_params()
        self._h = self._lib.vilba_create(int(device), C.byref(self.params))
        if not self._h:
            raise VilbaError(
                f"vilba_create(device={device}) failed: no usable CUDA device or kernels not loadable "
                "(the VI local-BA path has no CPU fallback)"
            )
        self.device = int(device)
        self._keep = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vilba_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, st: int, what: str):
        if st < 0:
            raise VilbaError(f"{what} failed ({st}): {self._lib.vilba_last_error(self._h).decode()}")

    # ---- entry 1: Optimizer::LocalBundleAdjustmentNavState phases C..E ---------------------------
    def local_ba(self, win: Window, stop_flag: Optional[np.ndarray] = None) -> Result:
        res = Result.alloc(win)
        cw, cr = win.as_c(), res.as_c()
        sf = stop_flag.ctypes.data_as(C.POINTER(C.c_uint8)) if stop_flag is not None else None
        st = self._lib.vilba_local_ba(self._h, C.byref(cw), C.byref(cr), sf)
        self._check(st, "vilba_local_ba")
        return res.take(cr)

    # ---- entry 1b: Optimizer::GlobalBundleAdjustmentNavState (src/Optimizer.cpp:1392-1668) -------
    def global_ba(self, win: Window, n_iterations: int = 10, robust: bool = False,
                  stop_flag: Optional[np.ndarray] = None) -> Result:
        """`win` is the whole map as one window: every good key-frame (mnId 0 fixed, all with a bias vertex) and
        every map point with an observation; one optimize(n_iterations), no cull."""
        res = Result.alloc(win)
        cw, cr = win.as_c(), res.as_c()
        sf = stop_flag.ctypes.data_as(C.POINTER(C.c_uint8)) if stop_flag is not None else None
        st = self._lib.vilba_global_ba(self._h, C.byref(cw), int(n_iterations), int(bool(robust)), C.byref(cr), sf)
        self._check(st, "vilba_global_ba")
        return res.take(cr)

    # ---- diagnostics (include/vilba_diag.h): the normal equations of the first LM trial ----------
    def first_trial_system(self, win: Window) -> dict:
        n, P, E = 15 * win.n_free, win.n_pts, win.n_obs
        o = dict(lam=np.zeros(1), Hpp=np.zeros((n, n)), bp=np.zeros(n), Hll=np.zeros((P, 6)), bl=np.zeros((P, 3)),
                 W=np.zeros((E, 6, 3)), S=np.zeros((n, n)), bs=np.zeros(n))
        cw = win.as_c()
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
        st = self._lib.vilba_diag_first_trial(self._h, C.byref(cw), *[dp(o[k]) for k in ("lam", "Hpp", "bp", "Hll", "bl", "W", "S", "bs")])
        self._check(st, "vilba_diag_first_trial")
        o["lam"] = float(o["lam"][0])
        return o

    def local_ba_batch(self, wins: Sequence[Window]) -> List[Result]:
        n = len(wins)
        results = [Result.alloc(w) for w in wins]
        cws = (CWindow * n)(*[w.as_c() for w in wins])
        crs = (CResult * n)(*[r.as_c() for r in results])
        st = self._lib.vilba_local_ba_batch(self._h, n, cws, crs)
        self._check(st, "vilba_local_ba_batch")
        return [r.take(crs[i]) for i, r in enumerate(results)]

    # ---- device-resident variant (bench: inputs already in HBM) ---------------------------------
    def upload(self, win: Window):
        self._keep = win
        cw = win.as_c()
        self._check(self._lib.vilba_window_upload(self._h, C.byref(cw)), "vilba_window_upload")

    def solve_resident(self) -> Result:
        res = Result(kf_state=np.zeros((0, 22)), pt_xyz=np.zeros((0, 3)), obs_outlier=np.zeros(0, np.uint8),
                     obs_chi2=np.zeros(0))
        cr = CResult()
        st = self._lib.vilba_window_solve_resident(self._h, C.byref(cr))
        self._check(st, "vilba_window_solve_resident")
        return res.take(cr)

    def download(self) -> Result:
        res = Result.alloc(self._keep)
        cr = res.as_c()
        self._check(self._lib.vilba_window_download(self._h, C.byref(cr)), "vilba_window_download")
        return res

    # ---- the same two calls with caller-owned, reusable buffers (what a C++ caller does) -----------
    def prepare(self, wins: Sequence[Window], chi2: bool = True) -> "PreparedBatch":
        """Allocates the result arrays and the ctypes views of `wins` once; run() is then only the C-ABI call.
        `chi2=False` asks for what the reference's function returns (states, points, outlier flags) and no per-edge chi2."""
        return PreparedBatch(self, list(wins), chi2)

    # ---- resident batch of independent windows (one batched launch per kernel) --------------------
    def upload_batch(self, wins: Sequence[Window]):
        self._keep_batch = list(wins)
        n = len(wins)
        cws = (CWindow * n)(*[w.as_c() for w in wins])
        self._check(self._lib.vilba_batch_upload(self._h, n, cws), "vilba_batch_upload")

    def batch_groups(self) -> int:
        return int(self._lib.vilba_batch_groups(self._h))

    def solve_batch_resident(self) -> List[Result]:
        n = len(self._keep_batch)
        crs = (CResult * n)()
        self._check(self._lib.vilba_batch_solve_resident(self._h, n, crs), "vilba_batch_solve_resident")
        empty = lambda: Result(kf_state=np.zeros((0, 22)), pt_xyz=np.zeros((0, 3)), obs_outlier=np.zeros(0, np.uint8),
                               obs_chi2=np.zeros(0))
        return [empty().take(crs[i]) for i in range(n)]

    def download_batch(self) -> List[Result]:
        n = len(self._keep_batch)
        results = [Result.alloc(w) for w in self._keep_batch]
        crs = (CResult * n)(*[r.as_c() for r in results])
        self._check(self._lib.vilba_batch_download(self._h, n, crs), "vilba_batch_download")
        return results

    # ---- one large window sharded by map point over the GPUs of a node ------------------------------
    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """Collective: joins the NCCL communicator of the sharded solve (unique_id from comm_unique_id() on rank 0)."""
        if len(unique_id) != 128:
            raise ValueError("the NCCL unique id is 128 bytes")
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self._lib.vilba_comm_init(self._h, buf, int(rank), int(world)), "vilba_comm_init")

    # ---- entry 2: IMUPreintegrator::update loop, batched ----------------------------------------
    def preintegrate_batch(self, sample_begin, gyro, acc, dt, bg, ba) -> np.ndarray:
        sb = np.ascontiguousarray(sample_begin, dtype=np.int32)
        n = sb.size - 1
        g = np.ascontiguousarray(gyro, dtype=np.float64).reshape(-1)
        a = np.ascontiguousarray(acc, dtype=np.float64).reshape(-1)
        t = np.ascontiguousarray(dt, dtype=np.float64).reshape(-1)
        b1 = np.ascontiguousarray(bg, dtype=np.float64).reshape(-1)
        b2 = np.ascontiguousarray(ba, dtype=np.float64).reshape(-1)
        if g.size != 3 * sb[-1] or a.size != g.size or t.size != sb[-1] or b1.size != 3 * n or b2.size != 3 * n:
            raise ValueError("inconsistent IMU batch shapes")
        out = np.zeros((n, PREINT_DOUBLES), np.float64)
        dp = C.POINTER(C.c_double)
        st = self._lib.vilba_preintegrate_batch(
            self._h, n, sb.ctypes.data_as(C.POINTER(C.c_int32)), g.ctypes.data_as(dp), a.ctypes.data_as(dp),
            t.ctypes.data_as(dp), b1.ctypes.data_as(dp), b2.ctypes.data_as(dp), out.ctypes.data_as(dp))
        self._check(st, "vilba_preintegrate_batch")
        return out

    def preintegrate_batch_dev(self, n_pairs: int, n_samples: int, sample_begin_ptr: int, gyro_ptr: int, acc_ptr: int,
                               dt_ptr: int, bg_ptr: int, ba_ptr: int, out_ptr: int):
        st = self._lib.vilba_preintegrate_batch_dev(self._h, n_pairs, n_samples, sample_begin_ptr, gyro_ptr, acc_ptr,
                                                    dt_ptr, bg_ptr, ba_ptr, out_ptr)
        self._check(st, "vilba_preintegrate_batch_dev")

    # ---- introspection ---------------------------------------------------------------------------
    def stats(self) -> Stats:
        s = Stats()
        self._lib.vilba_get_stats(self._h, C.byref(s))
        return s

    def reset_stats(self):
        self._lib.vilba_reset_stats(self._h)

    def set_profiling(self, on: bool):
        self._lib.vilba_set_profiling(self._h, int(bool(on)))


class PreparedBatch:
    """Host buffers of a batch, owned by the caller and reused across calls: the inputs as ctypes views of the
    numpy arrays, the outputs pre-allocated.  run() = vilba_local_ba_batch (or vilba_local_ba for one window)."""

    def __init__(self, ctx: Context, wins: List[Window], chi2: bool = True):
        self.ctx, self.wins = ctx, wins
        n = len(wins)
        self.results = [Result.alloc(w, chi2) for w in wins]
        for r in self.results:  # touch the pages now: the first write is otherwise paid inside the call
            r.kf_state.fill(0), r.pt_xyz.fill(0), r.obs_outlier.fill(0)
            if r.obs_chi2 is not None:
                r.obs_chi2.fill(0)
        self._cws = (CWindow * n)(*[w.as_c() for w in wins])
        self._crs = (CResult * n)(*[r.as_c() for r in self.results])

    def run(self) -> List[Result]:
        lib, h, n = self.ctx._lib, self.ctx._h, len(self.wins)
        if n == 1:
            st = lib.vilba_local_ba(h, C.byref(self._cws[0]), C.byref(self._crs[0]), None)
        else:
            st = lib.vilba_local_ba_batch(h, n, self._cws, self._crs)
        self.ctx._check(st, "vilba_local_ba(_batch)")
        return self.results

    def collect(self) -> List[Result]:
        """Copies status / trace out of the C structs (not part of the timed call)."""
        return [r.take(self._crs[i]) for i, r in enumerate(self.results)]


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the C ABI (call on rank 0, broadcast the 128 bytes)."""
    buf = C.create_string_buffer(128)
    if capi.load_library().vilba_comm_unique_id(buf) != 0:
        raise VilbaError("vilba_comm_unique_id failed (libnccl.so.2 not loadable)")
    return buf.raw


def max_batch() -> int:
    return int(capi.load_library().vilba_max_batch())


def version() -> str:
    return capi.load_library().vilba_version().decode()
