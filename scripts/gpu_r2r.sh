#!/bin/bash
TAG=${1:-r02r}
mkdir -p gpurun_out
SHAPES="14208x5040 20000x5040 50000x5040"
for SM in 0 80000 120000; do
  echo "== PQB_FULLS_SMEM=$SM" | tee -a gpurun_out/occ_$TAG.log
  PQB_FULLS_SMEM=$SM python scripts/shape_sweep.py $SHAPES 2>&1 | tee -a gpurun_out/occ_$TAG.log
done
