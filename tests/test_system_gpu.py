"""The normal equations themselves: what the CUDA kernels assemble for the first Levenberg-Marquardt trial of a window
(linearise + accumulate = BlockSolver::buildSystem, then the Schur step with the initial lambda, block_solver.hpp:354-439)
against the oracle's system for the same lambda -- H_pp, b_p, H_ll, b_l, the H_pl blocks, S and b_s entry by entry, not
through the LM trajectory.  Through the C ABI (include/vilba_diag.h: vilba_diag_first_trial)."""
import numpy as np
import pytest

from mc_slam_b200 import synth

pytestmark = pytest.mark.gpu


def _close(got, ref, rel=1e-11):
    scale = max(1.0, float(np.abs(ref).max(initial=0.0)))
    return float(np.abs(got - ref).max(initial=0.0)) <= rel * scale


@pytest.mark.parametrize("name,kw", [("tiny", {}), ("small", dict(n_fixed_extra=2)), ("c1", {}), ("c3", {})])
def test_first_trial_system_matches_oracle(vilba, oracle, name, kw):
    w = synth.make_config(name, **kw)
    with vilba.Context(0) as ctx:
        g = ctx.first_trial_system(w)
    assert g["lam"] > 0
    o = oracle.debug_system(w, lam=g["lam"], robust_mono=True)
    n = 15 * w.n_free
    iu = np.triu_indices(n)
    # computeLambdaInit: tau * max |diag H| over the pose blocks and the landmark blocks
    maxdiag = max(np.abs(np.diag(o["Hpp"])).max(), np.abs(np.einsum("pii->pi", o["Hll"])).max())
    assert np.isclose(g["lam"], 1e-5 * maxdiag, rtol=1e-12)
    # buildSystem (the oracle restores the diagonals after its solve: H_pp and H_ll come back without lambda)
    assert _close(g["Hpp"][iu], o["Hpp"][iu]) and _close(g["bp"], o["bp"])
    Hll_g = np.zeros_like(o["Hll"])
    for k, (i, j) in enumerate(((0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2))):
        Hll_g[:, i, j] = Hll_g[:, j, i] = g["Hll"][:, k]
    assert _close(Hll_g, o["Hll"]) and _close(g["bl"], o["bl"])
    assert _close(g["W"], o["Hpl"])
    # Schur step: S = H_pp + lambda I - sum_l W D^-1 W^T, b_s = b_p - sum_l W D^-1 b_l
    assert _close(g["S"][iu], o["S"][iu], rel=1e-10) and _close(g["bs"], o["bs"], rel=1e-10)
