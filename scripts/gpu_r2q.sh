#!/bin/bash
TAG=${1:-r02q}
mkdir -p gpurun_out
SHAPES="5000x2520 6272x2520 7104x2520 9472x2520 6272x5040 7104x5040 9472x5040 10016x5040 12800x5040 14208x5040"
for SB in 164 230 300 460; do
  echo "== PQB_SMALL_BLOCKS=$SB" | tee -a gpurun_out/small_$TAG.log
  PQB_SMALL_BLOCKS=$SB python scripts/shape_sweep.py $SHAPES 2>&1 | tee -a gpurun_out/small_$TAG.log
done
echo "== PQB_SMALL_BLOCKS=300 PQB_TAIL_SPLIT=0" | tee -a gpurun_out/small_$TAG.log
PQB_SMALL_BLOCKS=300 PQB_TAIL_SPLIT=0 python scripts/shape_sweep.py $SHAPES 2>&1 | tee -a gpurun_out/small_$TAG.log
