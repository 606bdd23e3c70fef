"""Synthetic EuRoC-shaped local-BA windows and IMU batches.

Pure numpy input generation: camera/IMU calibration from the reference's config/euroc.yaml:40-65,
a smooth figure-8 trajectory, 200 Hz IMU with the reference's noise constants
(src/IMU/imudata.cpp:25-31), map points observed by contiguous runs of key-frames, float32
quantisation where the reference stores CV_32F (src/Converter.cpp:122-129,153-160).

The pre-integrated measurements that a window carries are produced here by an independent numpy
restatement of the Forster recurrence (so generating inputs needs neither the CUDA library nor the
oracle); tests cross-check it against both.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from .capi import KF_FIXED, KF_HAS_BIAS, PREINT_DOUBLES, Window

SEED_BASE = 20261018

# config/euroc.yaml:54-65
FX, FY, CX, CY = 458.654, 457.296, 367.215, 248.375
IMG_W, IMG_H = 752, 480
# config/euroc.yaml:40-44 (Camera.Tbc)
_TBC = np.array(
    [
        [0.0148655429818, -0.999880929698, 0.00414029679422, -0.0216401454975],
        [0.999557249008, 0.0149672133247, 0.025715529948, -0.064676986768],
        [-0.0257744366974, 0.00375618835797, 0.999660727178, 0.00981073058949],
        [0.0, 0.0, 0.0, 1.0],
    ]
)
IMU_DT = 0.005  # 200 Hz
KF_DT = 0.2  # key-frame every 0.2 s = 40 IMU samples
GYR_MEAS_COV = 1.7e-4 * 1.7e-4 / 0.005
ACC_MEAS_COV = 2.0e-3 * 2.0e-3 / 0.005 * 100
SIGMA_G = 1.7e-4 / np.sqrt(0.005)
SIGMA_A = 2.0e-3 / np.sqrt(0.005)


# ------------------------------------------------------------------------------------------------
# small rotation helpers (numpy, batched on the leading axes)
# ------------------------------------------------------------------------------------------------
def hat(v: np.ndarray) -> np.ndarray:
    v = np.asarray(v, dtype=np.float64)
    o = np.zeros(v.shape[:-1] + (3, 3))
    o[..., 0, 1], o[..., 0, 2] = -v[..., 2], v[..., 1]
    o[..., 1, 0], o[..., 1, 2] = v[..., 2], -v[..., 0]
    o[..., 2, 0], o[..., 2, 1] = -v[..., 1], v[..., 0]
    return o


def exp_so3(w: np.ndarray) -> np.ndarray:
    """Rodrigues; returns rotation matrices."""
    w = np.asarray(w, dtype=np.float64)
    th = np.linalg.norm(w, axis=-1)[..., None, None]
    W = hat(w)
    small = th < 1e-8
    ths = np.where(small, 1.0, th)
    A = np.where(small, 1.0 - th**2 / 6.0, np.sin(ths) / ths)
    B = np.where(small, 0.5 - th**2 / 24.0, (1.0 - np.cos(ths)) / ths**2)
    return np.eye(3) + A * W + B * (W @ W)


def log_so3(R: np.ndarray) -> np.ndarray:
    q = mat_to_quat(R)
    n = np.linalg.norm(q[..., 1:], axis=-1)
    w = q[..., 0]
    f = np.where(n < 1e-10, 2.0 / np.where(w == 0, 1, w), 2.0 * np.arctan2(n, w) / np.where(n < 1e-10, 1.0, n))
    return f[..., None] * q[..., 1:]


def right_jacobian(w: np.ndarray) -> np.ndarray:
    w = np.asarray(w, dtype=np.float64)
    th = np.linalg.norm(w, axis=-1)[..., None, None]
    small = th < 1e-5
    ths = np.where(small, 1.0, th)
    K = hat(w / ths[..., 0])
    J = np.eye(3) - (1 - np.cos(ths)) / ths * K + (1 - np.sin(ths) / ths) * (K @ K)
    return np.where(small, np.eye(3), J)


def mat_to_quat(R: np.ndarray) -> np.ndarray:
    """Rotation matrix -> unit quaternion (w,x,y,z), w >= 0; batched."""
    R = np.asarray(R, dtype=np.float64)
    shp = R.shape[:-2]
    Rf = R.reshape(-1, 3, 3)
    q = np.zeros((Rf.shape[0], 4))
    for n in range(Rf.shape[0]):
        m = Rf[n]
        t = np.trace(m)
        if t > 0:
            s = np.sqrt(t + 1.0)
            w = 0.5 * s
            s = 0.5 / s
            q[n] = [w, (m[2, 1] - m[1, 2]) * s, (m[0, 2] - m[2, 0]) * s, (m[1, 0] - m[0, 1]) * s]
        else:
            i = int(np.argmax(np.diag(m)))
            j, k = (i + 1) % 3, (i + 2) % 3
            s = np.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
            v = np.zeros(3)
            v[i] = 0.5 * s
            s = 0.5 / s
            w = (m[k, j] - m[j, k]) * s
            v[j] = (m[j, i] + m[i, j]) * s
            v[k] = (m[k, i] + m[i, k]) * s
            q[n] = [w, v[0], v[1], v[2]]
        if q[n, 0] < 0:
            q[n] = -q[n]
        q[n] /= np.linalg.norm(q[n])
    return q.reshape(shp + (4,))


def quat_to_mat(q: np.ndarray) -> np.ndarray:
    q = np.asarray(q, dtype=np.float64)
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z)
    R[..., 0, 1] = 2 * (x * y - z * w)
    R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w)
    R[..., 1, 1] = 1 - 2 * (x * x + z * z)
    R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w)
    R[..., 2, 1] = 2 * (y * z + x * w)
    R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def calib_tbc() -> Tuple[np.ndarray, np.ndarray]:
    """Rbc re-orthonormalised through a quaternion like configparam.cpp:55-56, and Pbc."""
    R = quat_to_mat(mat_to_quat(_TBC[:3, :3]))
    return R, _TBC[:3, 3].copy()


def inv_level_sigma2(n_levels: int = 8, scale: float = 1.2) -> np.ndarray:
    """float32 pyramid constants built like ORBextractor.cpp:427-442."""
    sf = np.ones(n_levels, np.float32)
    s2 = np.ones(n_levels, np.float32)
    for i in range(1, n_levels):
        sf[i] = np.float32(sf[i - 1] * np.float32(scale))
        s2[i] = np.float32(sf[i] * sf[i])
    return (np.float32(1.0) / s2).astype(np.float32)


# ------------------------------------------------------------------------------------------------
# trajectory (body frame in world): figure-8, heading follows the velocity, +-10 deg roll/pitch
# ------------------------------------------------------------------------------------------------
_R0 = np.array([[0.0, 0.0, 1.0], [0.0, -1.0, 0.0], [1.0, 0.0, 0.0]])  # columns: body x=up, y=right(-Y), z=forward(+X)


TRAJ_RATE = 0.3  # time scale of the figure-8 (keeps ~8 co-visible key-frames per point at 5 KF/s)


def _pos(t):
    t = TRAJ_RATE * np.asarray(t, dtype=np.float64)
    return np.stack([2 * np.sin(0.5 * t), 1.5 * np.sin(t), 0.3 * np.sin(0.7 * t) + 1.0], -1)


def _vel(t):
    t = TRAJ_RATE * np.asarray(t, dtype=np.float64)
    return TRAJ_RATE * np.stack([np.cos(0.5 * t), 1.5 * np.cos(t), 0.21 * np.cos(0.7 * t)], -1)


def _acc(t):
    t = TRAJ_RATE * np.asarray(t, dtype=np.float64)
    return TRAJ_RATE**2 * np.stack([-0.5 * np.sin(0.5 * t), -1.5 * np.sin(t), -0.147 * np.sin(0.7 * t)], -1)


def _rot(t):
    t = np.asarray(t, dtype=np.float64)
    v = _vel(t)
    yaw = np.arctan2(v[..., 1], v[..., 0])
    roll = np.deg2rad(10.0) * np.sin(0.9 * TRAJ_RATE * t)
    pitch = np.deg2rad(10.0) * np.sin(0.6 * TRAJ_RATE * t + 0.5)
    z = np.zeros_like(t)
    Rz = exp_so3(np.stack([z, z, yaw], -1))
    Ry = exp_so3(np.stack([z, pitch, z], -1))
    Rx = exp_so3(np.stack([roll, z, z], -1))
    return Rz @ Ry @ Rx @ _R0


# ------------------------------------------------------------------------------------------------
# numpy pre-integration (batched over pairs, sequential over samples)
# ------------------------------------------------------------------------------------------------
def preintegrate_numpy(gyro: np.ndarray, acc: np.ndarray, dt: np.ndarray, bg: np.ndarray, ba: np.ndarray,
                       gyr_cov: float = GYR_MEAS_COV, acc_cov: float = ACC_MEAS_COV) -> np.ndarray:
    """gyro/acc: (N,S,3), dt: (N,S), bg/ba: (N,3) -> (N,142) records (include/vilba.h layout).

    Independent restatement of Forster et al. pre-integration as the reference iterates it
    (src/IMU/IMUPreintegrator.cpp:63-112); used only to manufacture the windows' measurements."""
    gyro = np.asarray(gyro, np.float64)
    acc = np.asarray(acc, np.float64)
    N, S = gyro.shape[0], gyro.shape[1]
    dP = np.zeros((N, 3))
    dV = np.zeros((N, 3))
    dR = np.tile(np.eye(3), (N, 1, 1))
    JPg, JPa, JVg, JVa, JRg = (np.zeros((N, 3, 3)) for _ in range(5))
    cov = np.zeros((N, 9, 9))
    T = np.zeros(N)
    I3 = np.eye(3)
    for s in range(S):
        h = dt[:, s][:, None, None]
        h1 = dt[:, s][:, None]
        w = gyro[:, s] - bg
        a = acc[:, s] - ba
        dRk = exp_so3(w * h1)
        Jr = right_jacobian(w * h1)
        ah = hat(a)
        A = np.tile(np.eye(9), (N, 1, 1))
        A[:, 6:9, 6:9] = np.swapaxes(dRk, 1, 2)
        A[:, 3:6, 6:9] = -dR @ ah * h
        A[:, 0:3, 6:9] = -0.5 * dR @ ah * h * h
        A[:, 0:3, 3:6] = I3 * h
        Bg = np.zeros((N, 9, 3))
        Bg[:, 6:9] = Jr * h
        Ca = np.zeros((N, 9, 3))
        Ca[:, 3:6] = dR * h
        Ca[:, 0:3] = 0.5 * dR * h * h
        cov = A @ cov @ np.swapaxes(A, 1, 2) + gyr_cov * Bg @ np.swapaxes(Bg, 1, 2) + acc_cov * Ca @ np.swapaxes(Ca, 1, 2)
        JPa_n = JPa + JVa * h - 0.5 * dR * h * h
        JPg_n = JPg + JVg * h - 0.5 * dR @ ah @ JRg * h * h
        JVa_n = JVa - dR * h
        JVg_n = JVg - dR @ ah @ JRg * h
        JRg_n = np.swapaxes(dRk, 1, 2) @ JRg - Jr * h
        JPa, JPg, JVa, JVg, JRg = JPa_n, JPg_n, JVa_n, JVg_n, JRg_n
        Ra = np.einsum("nij,nj->ni", dR, a)
        dP = dP + dV * h1 + 0.5 * Ra * h1 * h1
        dV = dV + Ra * h1
        dR = quat_to_mat(mat_to_quat(dR @ dRk))
        T = T + dt[:, s]
    out = np.zeros((N, PREINT_DOUBLES))
    out[:, 0:3], out[:, 3:6] = dP, dV
    out[:, 6:15] = dR.reshape(N, 9)
    out[:, 15:24], out[:, 24:33] = JPg.reshape(N, 9), JPa.reshape(N, 9)
    out[:, 33:42], out[:, 42:51] = JVg.reshape(N, 9), JVa.reshape(N, 9)
    out[:, 51:60] = JRg.reshape(N, 9)
    out[:, 60:141] = cov.reshape(N, 81)
    out[:, 141] = T
    return out


# ------------------------------------------------------------------------------------------------
# IMU batches (BASELINE config C2)
# ------------------------------------------------------------------------------------------------
@dataclass
class ImuBatch:
    sample_begin: np.ndarray  # (N+1,) i32
    gyro: np.ndarray  # (Ns,3)
    acc: np.ndarray  # (Ns,3)
    dt: np.ndarray  # (Ns,)
    bg: np.ndarray  # (N,3)
    ba: np.ndarray  # (N,3)

    @property
    def n_pairs(self) -> int:
        return int(self.sample_begin.size - 1)


def _imu_samples(rng: np.random.Generator, t0: np.ndarray, n_samples: int, bg_true: np.ndarray, ba_true: np.ndarray,
                 g_w: np.ndarray, noise: bool = True):
    """Ideal body-frame gyro/acc on t0[:,None] + i*IMU_DT plus bias and white noise; shapes (N,S,3)."""
    ts = t0[:, None] + IMU_DT * np.arange(n_samples)[None, :]
    R0 = _rot(ts.reshape(-1)).reshape(ts.shape + (3, 3))
    R1 = _rot((ts + IMU_DT).reshape(-1)).reshape(ts.shape + (3, 3))
    # piecewise-constant body rate that takes R(t) exactly to R(t+dt)
    gyro = log_so3(np.swapaxes(R0, -1, -2) @ R1) / IMU_DT
    # mid-point specific force, body frame
    a_w = _acc(ts + 0.5 * IMU_DT) - g_w
    acc = np.einsum("nsji,nsj->nsi", R0, a_w)
    gyro = gyro + bg_true[:, None, :]
    acc = acc + ba_true[:, None, :]
    if noise:
        gyro = gyro + rng.normal(0.0, SIGMA_G, gyro.shape)
        acc = acc + rng.normal(0.0, SIGMA_A, acc.shape)
    return gyro, acc


def make_imu_batch(n_pairs: int = 4096, n_samples: int = 40, seed: int = SEED_BASE + 2, ragged: bool = False,
                   leading_partial: bool = False) -> ImuBatch:
    """C2: n_pairs key-frame pairs x n_samples IMU samples.  ragged=True draws 10..100 samples per
    pair (the product's realistic range); leading_partial=True prepends the partial interval sample
    that KeyFrame::ComputePreInt feeds first (src/KeyFrame.cpp:214-218)."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    if ragged:
        counts = rng.integers(10, 101, size=n_pairs)
    else:
        counts = np.full(n_pairs, n_samples)
    smax = int(counts.max())
    t0 = rng.uniform(0.0, 60.0, size=n_pairs)
    bg_true = rng.normal(0.0, 0.01, (n_pairs, 3))
    ba_true = rng.normal(0.0, 0.05, (n_pairs, 3))
    g_w = np.array([0.0, 0.0, -9.81])
    gyro, acc = _imu_samples(rng, t0, smax, bg_true, ba_true, g_w)
    bg = bg_true + rng.normal(0.0, 1e-3, (n_pairs, 3))
    ba = ba_true + rng.normal(0.0, 1e-2, (n_pairs, 3))
    G, A, D, begin = [], [], [], [0]
    for p in range(n_pairs):
        c = int(counts[p])
        g, a, d = gyro[p, :c], acc[p, :c], np.full(c, IMU_DT)
        if leading_partial:
            lead = rng.uniform(0.0, IMU_DT)
            g = np.concatenate([g[:1], g])
            a = np.concatenate([a[:1], a])
            d = np.concatenate([[lead], d])
        G.append(g)
        A.append(a)
        D.append(d)
        begin.append(begin[-1] + g.shape[0])
    return ImuBatch(
        sample_begin=np.asarray(begin, np.int32),
        gyro=np.ascontiguousarray(np.concatenate(G)),
        acc=np.ascontiguousarray(np.concatenate(A)),
        dt=np.ascontiguousarray(np.concatenate(D)),
        bg=np.ascontiguousarray(bg),
        ba=np.ascontiguousarray(ba),
    )


# ------------------------------------------------------------------------------------------------
# windows (BASELINE configs C1, C3, C4, C5)
# ------------------------------------------------------------------------------------------------
CONFIGS = {
    # name: (n_kf, n_pts, mean run length, seed offset)
    # (the run-length parameter is tuned so that visibility clipping leaves ~15k / ~40k / ~600k edges)
    "c1": (10, 2000, 20.0, 1),
    "c3": (20, 5000, 10.0, 3),
    "c4": (100, 50000, 18.0, 4),
    "tiny": (4, 40, 3.0, 11),
    "small": (6, 300, 4.0, 12),
}


def make_window(n_kf: int = 20, n_pts: int = 5000, mean_run: float = 8.0, seed: int = SEED_BASE + 3,
                n_fixed_extra: int = 0, outlier_frac: float = 0.02, t_start: Optional[float] = None) -> Window:
    """One synthetic window: KF0 is the fixed anchor (plays pKFPrevLocal, src/Optimizer.cpp:2366) with a
    fixed bias vertex, KF1..K-1 are free; `n_fixed_extra` additional fixed covisible key-frames (older
    than the anchor) observe some of the points."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    Rbc, Pbc = calib_tbc()
    g_w = np.array([0.0, 0.0, -9.81], np.float32).astype(np.float64)  # gw is CV_32F
    F = int(n_fixed_extra)
    Kc = n_kf  # chain key-frames (anchor + free)
    K = Kc + F
    t0 = float(rng.uniform(0.0, 30.0)) if t_start is None else float(t_start)
    # chain times; extra fixed key-frames sit before the anchor
    t_chain = t0 + KF_DT * np.arange(Kc)
    t_extra = t0 - KF_DT * (1 + np.arange(F))[::-1]
    # array order: extra fixed (oldest first), anchor, free ...  (increasing KeyFrame id)
    t_kf = np.concatenate([t_extra, t_chain])
    Rwb = _rot(t_kf)
    Pwb = _pos(t_kf)
    Vwb = _vel(t_kf)
    bg_true = rng.normal(0.0, 0.01, 3)
    ba_true = rng.normal(0.0, 0.05, 3)
    bg_nom = bg_true + rng.normal(0.0, 1e-3, 3)
    ba_nom = ba_true + rng.normal(0.0, 1e-2, 3)

    # --- IMU + pre-integration for the chain ---
    NI = Kc - 1
    n_s = int(round(KF_DT / IMU_DT))
    gyro, acc = _imu_samples(rng, t_chain[:-1], n_s, np.tile(bg_true, (NI, 1)), np.tile(ba_true, (NI, 1)), g_w)
    preint = preintegrate_numpy(gyro, acc, np.full((NI, n_s), IMU_DT), np.tile(bg_nom, (NI, 1)), np.tile(ba_nom, (NI, 1)))
    imu_i = F + np.arange(NI, dtype=np.int32)
    imu_j = imu_i + 1

    # --- camera poses (truth) ---
    Rwc = Rwb @ Rbc
    Pwc = np.einsum("kij,j->ki", Rwb, Pbc) + Pwb

    # --- points and observations ---
    inv_s2 = inv_level_sigma2()
    p_oct = 0.6 ** np.arange(8)
    p_oct /= p_oct.sum()
    geo_p = 1.0 / max(mean_run - 2.0 + 1.0, 1.0)
    pts = np.zeros((n_pts, 3))
    vis = np.zeros((n_pts, K), bool)
    uv_all = np.zeros((n_pts, K, 2))
    todo = np.arange(n_pts)
    for _round in range(200):
        if todo.size == 0:
            break
        m = todo.size
        L = np.minimum(K, 2 + rng.geometric(geo_p, m) - 1)
        s = np.floor(rng.uniform(0.0, 1.0, m) * (K - L + 1)).astype(np.int64)
        mid = s + L // 2
        u, v = rng.uniform(0, IMG_W, m), rng.uniform(0, IMG_H, m)
        d = rng.uniform(2.0, 8.0, m)
        pc = np.stack([(u - CX) / FX * d, (v - CY) / FY * d, d], -1)
        pw = np.einsum("mij,mj->mi", Rwc[mid], pc) + Pwc[mid]
        pcs = np.einsum("kji,mkj->mki", Rwc, pw[:, None, :] - Pwc[None, :, :])  # (m,K,3)
        z = pcs[..., 2]
        ok = z > 0.5
        zs = np.where(ok, z, 1.0)
        uu = FX * pcs[..., 0] / zs + CX
        vv = FY * pcs[..., 1] / zs + CY
        kk = np.arange(K)[None, :]
        ok &= (uu >= 0) & (uu < IMG_W) & (vv >= 0) & (vv < IMG_H) & (kk >= s[:, None]) & (kk < (s + L)[:, None])
        good = (ok.sum(1) >= 2) & ok[:, F + 1:].any(1)
        idx = todo[good]
        pts[idx] = pw[good]
        vis[idx] = ok[good]
        uv_all[idx, :, 0] = uu[good]
        uv_all[idx, :, 1] = vv[good]
        todo = todo[~good]
    if todo.size:
        raise RuntimeError("could not place all points")
    pp, kk = np.nonzero(vis)  # row-major: point-major, key-frame index ascending
    n_obs = pp.size
    begin = np.zeros(n_pts + 1, np.int32)
    np.cumsum(vis.sum(1), out=begin[1:])
    octv = rng.choice(8, size=n_obs, p=p_oct)
    sig = 1.2 ** octv
    uvn = uv_all[pp, kk] + rng.normal(0.0, 1.0, (n_obs, 2)) * sig[:, None]
    out = rng.uniform(0, 1, n_obs) < outlier_frac
    ang = rng.uniform(0, 2 * np.pi, n_obs)
    mag = rng.uniform(10.0, 40.0, n_obs)
    uvn[:, 0] += out * mag * np.cos(ang)
    uvn[:, 1] += out * mag * np.sin(ang)
    obs_kf = [kk.astype(np.int32)]
    obs_uv = [uvn.astype(np.float32)]
    obs_is2 = [inv_s2[octv]]

    # --- initial estimates: truth + noise; anchor and extra fixed KFs stay at truth ---
    kf_state = np.zeros((K, 22))
    flags = np.zeros(K, np.uint8)
    for k in range(K):
        free = k > F
        P0, V0, R0 = Pwb[k].copy(), Vwb[k].copy(), Rwb[k].copy()
        if free:
            P0 += rng.normal(0.0, 0.02, 3)
            V0 += rng.normal(0.0, 0.05, 3)
            R0 = R0 @ exp_so3(rng.normal(0.0, np.deg2rad(0.5), 3))
        kf_state[k, 0:3], kf_state[k, 3:6] = P0, V0
        kf_state[k, 6:10] = mat_to_quat(R0)
        kf_state[k, 10:13], kf_state[k, 13:16] = bg_nom, ba_nom
        flags[k] = (0 if free else KF_FIXED) | (KF_HAS_BIAS if k >= F else 0)
    pts0 = (pts + rng.normal(0.0, 0.03, pts.shape)).astype(np.float32).astype(np.float64)

    truth = dict(Pwb=Pwb, Vwb=Vwb, Rwb=Rwb, pts=pts, bg=bg_true, ba=ba_true, bg_nom=bg_nom, ba_nom=ba_nom,
                 imu_gyro=gyro, imu_acc=acc, t_kf=t_kf)
    return Window(
        kf_state=kf_state, kf_flags=flags, kf_id=np.arange(1, K + 1, dtype=np.int64),
        imu_kf_i=imu_i, imu_kf_j=imu_j, imu_preint=preint,
        pt_xyz=pts0, pt_obs_begin=np.asarray(begin, np.int32),
        obs_kf=np.concatenate(obs_kf), obs_uv=np.concatenate(obs_uv), obs_inv_sigma2=np.concatenate(obs_is2),
        fx=float(np.float32(FX)), fy=float(np.float32(FY)), cx=float(np.float32(CX)), cy=float(np.float32(CY)),
        Rbc=Rbc, Pbc=Pbc, gravity=g_w, truth=truth,
    )


def make_config(name: str, window_index: int = 0, **overrides) -> Window:
    """Named BASELINE configs.  `window_index` selects the w-th independent window of C5 (seed + 1000*w)."""
    n_kf, n_pts, run, off = CONFIGS[name]
    kw = dict(n_kf=n_kf, n_pts=n_pts, mean_run=run, seed=SEED_BASE + off + 1000 * window_index)
    kw.update(overrides)
    return make_window(**kw)
