#!/bin/bash
TAG=${1:-r03b}
SHAPES="5000x2520 5500x2520 5650x2520 6272x5040 7104x5040 9472x5040 14208x5040 50000x5040"
echo "== default" | tee -a gpurun_out/base_$TAG.log
python scripts/shape_sweep.py $SHAPES 2>&1 | tee -a gpurun_out/base_$TAG.log
echo "== PQB_FORCE_BASE=1" | tee -a gpurun_out/base_$TAG.log
PQB_FORCE_BASE=1 python scripts/shape_sweep.py $SHAPES 2>&1 | tee -a gpurun_out/base_$TAG.log
