#!/bin/bash
TAG=${1:-r03a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "one_block_per_sm or config2 or host_pipeline or closes_at" > gpurun_out/pytest_tail_$TAG.log 2>&1; grep -n "^E   \|passed\|failed" gpurun_out/pytest_tail_$TAG.log | cut -c1-300 | head -20
SHAPES="5000x2520 5500x2520 6272x2520 7104x2520 7872x2520 5000x5040 6272x5040 7104x5040 7872x5040 8000x5040"
for P in 3 2 5 9; do
  echo "== compact tail, PQB_TAIL_PARTS=$P" | tee -a gpurun_out/tail_$TAG.log
  PQB_TAIL_PARTS=$P python scripts/shape_sweep.py $SHAPES 2>&1 | tee -a gpurun_out/tail_$TAG.log
done
echo "== PQB_TAIL_COMPACT=0 (round-1 behaviour)" | tee -a gpurun_out/tail_$TAG.log
PQB_TAIL_COMPACT=0 python scripts/shape_sweep.py $SHAPES 2>&1 | tee -a gpurun_out/tail_$TAG.log
