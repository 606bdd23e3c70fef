#!/usr/bin/env python
"""Times the reduced-system kernels alone (include/vilba_diag.h) on S-like SPD systems: variant 1 = look-ahead cluster
kernel with shared-memory tiles (chol_la.cu), 2 = multi-kernel blocked factorisation (chol_big.cu).  Usage: bench_chol.py [n,n,...] [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc_slam_b200 import capi  # noqa: E402

sizes = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "135,285,300").split(",")]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rng = np.random.default_rng(1)
for n in sizes:
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    S = (q * np.logspace(0, 6, n)) @ q.T
    S = 0.5 * (S + S.T)
    b = rng.standard_normal(n)
    ref = np.linalg.solve(S, b)
    for variant, clusters in ((1, (1, 2, 4, 8, 16)), (2, (1,))):
        for cl in clusters:
            for nw in (1, 16, 64):
                if variant == 2 and (nw > 1 or n < 300):
                    continue
                if variant != 2 and n > 600:
                    continue
                if nw * cl > 148 and variant == 1:
                    continue
                if not capi.load_library().vilba_diag_dense_supported(n, variant, cl):
                    continue
                try:
                    x, fail, us = capi.diag_dense_solve(S, b, variant=variant, cluster=cl, n_windows=nw, reps=reps)
                except RuntimeError as e:
                    print(f"n={n} variant={variant} cluster={cl} windows={nw}: {e}", flush=True)
                    continue
                err = np.linalg.norm(x - ref) / np.linalg.norm(ref)
                print(f"n={n} variant={variant} cluster={cl} windows={nw}: {us:8.1f} us  rel.err {err:.1e} fail={fail}", flush=True)
