#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_windows.py -m gpu -q -x 2>&1 | tail -3
for MG in 1 0; do
PQB_WIN_MERGE=$MG PQB_WIN_VERBOSE=1 timeout 600 python - <<'PY' 2>&1 | tee -a gpurun_out/merge_r03x.log
import sys, os
sys.path.insert(0, ".")
import polars_quant_b200 as pq
from polars_quant_b200 import windows
eng = pq.get_engine(0)
for U in (0, 3, 4):
    if U: os.environ["PQB_WIN_UNITS"] = str(U)
    wp = windows.WindowPanel(10_000, 5_040, engine=eng, host_staging=False); wp.fill_synthetic()
    print("c5 merge=%s U=%s" % (os.environ["PQB_WIN_MERGE"], U or "planner"), "%.3f" % wp.time_device(warmup=2, iters=10)[0], flush=True); wp.close()
PY
done
