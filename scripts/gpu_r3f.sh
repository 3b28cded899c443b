#!/bin/bash
SHAPES="5000x2520 9472x5040 14208x5040 50000x5040"
for I in 0 1; do
  echo "== PQB_INTERLEAVE=$I" | tee -a gpurun_out/interleave_r03f.log
  PQB_INTERLEAVE=$I python scripts/shape_sweep.py $SHAPES 2>&1 | tee -a gpurun_out/interleave_r03f.log
done
