"""Where the end-to-end time of a batch goes: upload / solve / download of the resident API against the one-call entry."""
import sys
import time

sys.path.insert(0, '/root/repo')
from mc_slam_b200 import api, synth

nw = int(sys.argv[1]) if len(sys.argv) > 1 else 64
wins = [synth.make_config('c3', window_index=i) for i in range(nw)]
ctx = api.Context(0)
job = ctx.prepare(wins)  # ctypes arrays built once: the timed call is vilba_local_ba_batch itself
for _ in range(3):
    job.run()
for rep in range(3):
    t = time.perf_counter(); job.run(); e2e = time.perf_counter() - t
    t = time.perf_counter(); ctx.upload_batch(wins); up = time.perf_counter() - t
    t = time.perf_counter(); r = ctx.solve_batch_resident(); so = time.perf_counter() - t
    t = time.perf_counter(); ctx.download_batch(); dn = time.perf_counter() - t
    print(f"e2e {e2e*1e3:.2f} ms | upload {up*1e3:.2f}  solve call {so*1e3:.2f} (device {r[0].solve_ms:.2f})  download {dn*1e3:.2f}  sum {1e3*(up+so+dn):.2f}")
