#!/usr/bin/env python
"""Digest of an ncu --set full report: headline metrics and the SASS regions ranked by executed instructions.
usage: tools/ncu_digest.py report.ncu-rep [min_exec]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps']
for r in rows[2:]:
    for w in want:
        if w in hdr:
            print(f"{w:70s} {r[hdr.index(w)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr)]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: 0 for s in stalls}
for r in data:
    for s in stalls:
        agg[s] += int(r[ix[s]])
tot = sum(agg.values())
print("stalls:", [(k, round(100 * v / tot, 1)) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]])
cur = None
def flush(end):
    if cur is not None and acc > 0:
        top = sorted(st.items(), key=lambda x: -x[1])[:3]
        print(f"{start:4d}-{end-1:4d} n={end-start:3d} exec/inst~{cur:>9d} total={acc:>10d} samples={samp:6d} smem_wf={wf:>9d} ideal={wfi:>9d} {[(k[6:], v) for k, v in top]}")
for n, r in enumerate(data):
    ie = int(r[ix["Instructions Executed"]])
    if cur is None or abs(ie - cur) > 0.15 * max(cur, 1):
        flush(n)
        cur, start, acc, samp, wf, wfi, st = ie, n, 0, 0, 0, 0, {s: 0 for s in stalls}
    acc += ie
    samp += int(r[ix["# Samples"]])
    wf += int(r[ix["L1 Wavefronts Shared"]])
    wfi += int(r[ix["L1 Wavefronts Shared Ideal"]])
    for s in stalls:
        st[s] += int(r[ix[s]])
flush(len(data))
