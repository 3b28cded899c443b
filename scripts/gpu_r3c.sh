#!/bin/bash
TAG=${1:-r03c}
for RW in 0 1; do
PQB_WIN_RELAXED_WAIT=$RW timeout 600 python - <<'PY' 2>&1 | tee -a gpurun_out/c5_$TAG.log
import sys, os
sys.path.insert(0, ".")
import polars_quant_b200 as pq
from polars_quant_b200 import windows
eng = pq.get_engine(0)
wp = windows.WindowPanel(10_000, 5_040, engine=eng, host_staging=False); wp.fill_synthetic()
print("c5 relaxed_wait=%s" % os.environ["PQB_WIN_RELAXED_WAIT"], wp.time_device(warmup=2, iters=10)); wp.close()
wp = windows.WindowPanel(10_000, 5_040, engine=eng, host_staging=False, kdj=(9,), ext=(20, 55), atr=14); wp.fill_synthetic()
print("c5 small set relaxed_wait=%s" % os.environ["PQB_WIN_RELAXED_WAIT"], wp.time_device(warmup=2, iters=10)); wp.close()
PY
done
