#!/bin/bash
TAG=${1:-r02p}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_windows.py -m gpu -q -x > gpurun_out/pytest_win_$TAG.log 2>&1; grep -n "^E   \|passed\|failed" gpurun_out/pytest_win_$TAG.log | cut -c1-300 | head -20
PQB_WIN_PIPE=3 timeout 900 python -m pytest tests/test_gpu_windows.py -m gpu -q -x > gpurun_out/pytest_win_${TAG}_pipe3.log 2>&1; grep -n "^E   \|passed\|failed" gpurun_out/pytest_win_${TAG}_pipe3.log | cut -c1-300 | head -20
for ROOMY in 0 1; do
PQB_WIN_ROOMY=$ROOMY timeout 900 python - <<'PY' 2>&1 | tee -a gpurun_out/c5_$TAG.log
import sys, json, os
sys.path.insert(0, ".")
import polars_quant_b200 as pq
eng = pq.get_engine(0)
from polars_quant_b200 import windows
def run(tag, S=10_000, **kw):
    wp = windows.WindowPanel(S, 5_040, engine=eng, host_staging=False, **kw)
    wp.fill_synthetic()
    print("c5 roomy=%s" % os.environ["PQB_WIN_ROOMY"], tag, "%.3f" % wp.time_device()[0], flush=True)
    wp.close()
full = dict(kdj=(5, 9, 14, 60, 250), ext=(5, 20, 55, 250), atr=14)
for pipe in (0, 1, 2, 3):
    for U in (3, 4, 5):
        for st in (2, 3):
            os.environ["PQB_WIN_PIPE"] = str(pipe); os.environ["PQB_WIN_UNITS"] = str(U); os.environ["PQB_WIN_STAGES"] = str(st)
            run("pipe=%d U=%d stages=%d" % (pipe, U, st), **full)
PY
done
