#!/bin/bash
# sweep of the Schur tile-kernel launch shape on the C5 bench (N=1); prints value / ms_per_step / kernels_us
run() {
  echo "== $*"
  env "$@" python bench.py --steps 6 --warmup 3 --no-extra 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        j=json.loads(line); k=j.get('kernels_us') or {}; print(round(j['value']), round(j['ms_per_step'],2), {a:round(b) for a,b in k.items() if a!='note'})"
}
for v in "$@"; do run $v; done
