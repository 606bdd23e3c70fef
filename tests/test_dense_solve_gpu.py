"""GPU: the reduced-system kernels (replacing LinearSolverEigen::solve, g2o/solvers/linear_solver_eigen.h:94-124)
against numpy on dense symmetric systems, called through the C ABI (include/vilba_diag.h).  Covers the sizes of the
BASELINE configs (n = 135 / 285 / 1485), ragged sizes around the block and tile edges, every cluster size, and the
SimplicialLDLT failure rule: an indefinite matrix is factored through, only a zero pivot fails."""
import numpy as np
import pytest

from mc_slam_b200 import capi

pytestmark = pytest.mark.gpu


def spd(n, seed, cond=1e8):
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    ev = np.logspace(0, np.log10(cond), n)
    s = (q * ev) @ q.T
    return 0.5 * (s + s.T), rng.standard_normal(n)


def rel_err(S, b, x):
    ref = np.linalg.solve(S, b)
    return np.linalg.norm(x - ref) / np.linalg.norm(ref)


@pytest.mark.parametrize("n", [1, 3, 4, 15, 31, 32, 33, 60, 64, 135, 150, 256, 284, 285, 287, 288, 300])
@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
def test_lookahead_cluster_kernel_matches_numpy(n, cluster):
    lib = capi.load_library()
    if not lib.vilba_diag_dense_supported(n, 1, cluster):
        assert n > 170 and cluster < 8  # the tiles of C1 fit one CTA, those of C3 / n = 300 a cluster of 4
        pytest.skip("tiles do not fit the cluster's shared memory")
    S, b = spd(n, 100 + n)
    x, fail, _ = capi.diag_dense_solve(S, b, variant=1, cluster=cluster)
    assert fail == 0
    assert rel_err(S, b, x) < 1e-7  # cond 1e8 * eps ~ 2e-8
    assert lib is not None


def test_removed_variant_is_rejected():
    S, b = spd(30, 5)
    with pytest.raises(Exception):
        capi.diag_dense_solve(S, b, variant=0, cluster=1)


@pytest.mark.parametrize("n", [15, 100, 360, 405, 465, 480, 495, 1000, 1485])  # 345 < n <= 480: beyond chol_la in the product
def test_whole_gpu_kernel_matches_numpy(n):
    S, b = spd(n, 300 + n, cond=1e6)
    x, fail, _ = capi.diag_dense_solve(S, b, variant=2, cluster=1)
    assert fail == 0 and rel_err(S, b, x) < 1e-8


@pytest.mark.parametrize("n", [40, 285])
def test_indefinite_system_is_factored_through_like_simplicial_ldlt(n):
    # symmetric, indefinite, all leading minors non-zero: LDL^T exists with negative pivots in several blocks
    rng = np.random.default_rng(7 + n)
    S, b = spd(n, 400 + n, cond=1e3)
    flip = rng.choice(n, size=max(2, n // 10), replace=False)
    d = np.ones(n)
    d[flip] = -1.0
    low = np.linalg.cholesky(S)
    S = (low * d) @ low.T  # L diag(+-1) L^T
    S = 0.5 * (S + S.T)
    x, fail, _ = capi.diag_dense_solve(S, b, variant=1, cluster=8)
    assert fail == 0
    assert rel_err(S, b, x) < 1e-8
    batch_x, fail, _ = capi.diag_dense_solve(S, b, variant=1, cluster=2 if n < 100 else 4, n_windows=3)
    assert fail == 0 and rel_err(S, b, batch_x) < 1e-8


@pytest.mark.parametrize("n", [100, 600, 1485])
def test_whole_gpu_kernel_factors_an_indefinite_system_through(n):
    rng = np.random.default_rng(11 + n)
    S, b = spd(n, 500 + n, cond=1e3)
    d = np.ones(n)
    d[rng.choice(n, size=max(2, n // 10), replace=False)] = -1.0
    low = np.linalg.cholesky(S)
    S = (low * d) @ low.T
    S = 0.5 * (S + S.T)
    x, fail, _ = capi.diag_dense_solve(S, b, variant=2, cluster=1)
    assert fail == 0 and rel_err(S, b, x) < 1e-8


@pytest.mark.parametrize("n", [1, 31, 33, 63, 64, 65, 127, 129])
def test_whole_gpu_kernel_on_tile_edges(n):
    S, b = spd(n, 600 + n, cond=1e4)
    x, fail, _ = capi.diag_dense_solve(S, b, variant=2, cluster=1)
    assert fail == 0 and rel_err(S, b, x) < 1e-9


def test_zero_pivot_raises_the_failure_flag():
    S, b = spd(64, 9)
    S[0, :] = 0.0
    S[:, 0] = 0.0  # first pivot exactly zero
    _, fail, _ = capi.diag_dense_solve(S, b, variant=1, cluster=2)
    assert fail == 1
