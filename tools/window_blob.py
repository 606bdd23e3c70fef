#!/usr/bin/env python
"""Window blobs (the wire / on-disk format of include/vilba.h: vilba_window_serialize).

    python tools/window_blob.py make c3 out.blob [window_index]   # serialise a synthetic window
    python tools/window_blob.py info in.blob                      # counts, calibration, checksum status
    python tools/window_blob.py solve in.blob                     # replay through vilba_local_ba (needs a GPU)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc_slam_b200 import synth  # noqa: E402
from mc_slam_b200.capi import Window  # noqa: E402

cmd = sys.argv[1] if len(sys.argv) > 1 else "help"
if cmd == "make":
    w = synth.make_config(sys.argv[2], window_index=int(sys.argv[4]) if len(sys.argv) > 4 else 0)
    with open(sys.argv[3], "wb") as f:
        f.write(w.to_bytes())
    print(f"{sys.argv[3]}: {os.path.getsize(sys.argv[3])} bytes, {w.n_kf} KF / {w.n_pts} pts / {w.n_obs} obs / {w.n_imu} IMU edges")
elif cmd == "info":
    w = Window.from_bytes(open(sys.argv[2], "rb").read())  # raises on a bad magic / version / checksum / index
    print(f"{sys.argv[2]}: ok; {w.n_kf} KF ({w.n_free} free) / {w.n_pts} pts / {w.n_obs} obs / {w.n_imu} IMU edges; "
          f"fx {w.fx:.3f} fy {w.fy:.3f} cx {w.cx:.3f} cy {w.cy:.3f}; gravity {w.gravity.tolist()}")
elif cmd == "solve":
    from mc_slam_b200 import api
    w = Window.from_bytes(open(sys.argv[2], "rb").read())
    with api.Context(0) as ctx:
        r = ctx.local_ba(w)
    print(f"{len(r.trace)} LM iterations, chi2 {r.trace[0]['chi2_initial']:.3f} -> {r.trace[-1]['chi2_final']:.3f}, "
          f"{int(r.obs_outlier.sum())} outliers, {r.solve_ms:.3f} ms on the device")
else:
    print(__doc__)
