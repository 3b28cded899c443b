#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "one_block_per_sm or closes_at or flat_and_tied" 2>&1 | tail -2
SHAPES="5650x2520 6272x2520 6272x5040 7104x5040 8000x5040 9472x5040 10016x5040"
for M in 0 296; do
  echo "== PQB_MID_BLOCKS=$M" | tee -a gpurun_out/mid_r03v.log
  PQB_MID_BLOCKS=$M python scripts/shape_sweep.py $SHAPES 2>&1 | tee -a gpurun_out/mid_r03v.log
done
