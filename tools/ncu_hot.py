#!/usr/bin/env python
"""Top stall locations of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K` (SASS view)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = []
for r in rows[2:]:
    if len(r) != len(hdr):
        if body:
            break  # next kernel of the report
        continue
    if r[ix["# Samples"]] == "# Samples":
        continue
    body.append(r)
tot = sum(float(r[ix["# Samples"]] or 0) for r in body)
inst = sum(float(r[ix["Instructions Executed"]] or 0) for r in body)
print(f"total samples {tot:.0f}, warp instructions {inst:.0f}")
top = sorted(body, key=lambda r: -float(r[ix["# Samples"]] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in top:
    s = float(r[ix["# Samples"]] or 0)
    main = sorted(((float(r[ix[h]] or 0), h) for h in stalls), reverse=True)[:2]
    print(f"{100*s/tot:5.1f}%  ex={r[ix['Instructions Executed']]:>9s} thr={r[ix['Avg. Threads Executed']]:>5s}  {r[ix['Source']][:70]:70s} {main[0][1]}:{main[0][0]:.0f} {main[1][1]}:{main[1][0]:.0f}")
