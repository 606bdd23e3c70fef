#!/bin/bash
# timing ablations of the cluster Cholesky (needs a -DVILBA_CHOL_TIMING build; results are WRONG, only times matter)
# bits: 1 tiles, 2 panel fma, 4 potrf, 8 cluster barrier->syncthreads, 16 back-subst, 32 tile prefetch, 64 factor stores
for ab in ${@:-0 1 2 4 8 16 31}; do
  echo -n "ablate=$ab "
  VILBA_CHOL_ABLATE=$ab timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(d['kernels_us']['chol_solve'])"
done
