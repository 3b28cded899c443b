#!/bin/bash
TAG=${1:-r02z}
SHAPES="6272x5040 7104x5040 8000x5040 6272x2520"
for P in 0 2 3 4 7; do
  echo "== PQB_SPLIT_ALL=$P" | tee -a gpurun_out/splitall_$TAG.log
  PQB_SPLIT_ALL=$P python scripts/shape_sweep.py $SHAPES 2>&1 | tee -a gpurun_out/splitall_$TAG.log
done
