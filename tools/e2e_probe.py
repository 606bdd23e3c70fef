import time, sys
sys.path.insert(0,'/root/repo')
from mc_slam_b200 import api, synth
w = synth.make_config('c3'); ctx = api.Context(0)
for _ in range(3): ctx.local_ba(w)
t=time.perf_counter(); r=ctx.local_ba(w); dt=time.perf_counter()-t
print('e2e ms', dt*1e3, 'solve_ms', r.solve_ms)
t=time.perf_counter(); ctx.upload(w); print('upload ms', (time.perf_counter()-t)*1e3)
t=time.perf_counter(); r=ctx.solve_resident(); print('solve call ms', (time.perf_counter()-t)*1e3, r.solve_ms)
t=time.perf_counter(); ctx.download(); print('download ms', (time.perf_counter()-t)*1e3)
