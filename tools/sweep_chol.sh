#!/bin/bash
# sweep the cluster size of chol_la for ONE window (batches take the smallest cluster that fits); run on the GPU box
for cs in 1 2 4 8 16; do
  echo -n "cluster=$cs "
  VILBA_CHOL_CLUSTER=$cs timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(d['value'], d['kernels_us']['chol_solve'])"
done
