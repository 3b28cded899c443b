#!/bin/bash
TAG=${1:-r02l}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_windows.py tests/test_gpu_ref_golden.py -m gpu -q -x > gpurun_out/pytest_win_$TAG.log 2>&1; grep -n "^E   \|passed\|failed" gpurun_out/pytest_win_$TAG.log | cut -c1-300 | head -20
timeout 900 python - <<'PY' 2>&1 | tee gpurun_out/c5_$TAG.log
import sys, json, os
sys.path.insert(0, ".")
import polars_quant_b200 as pq
eng = pq.get_engine(0)
from polars_quant_b200 import windows
os.environ["PQB_WIN_VERBOSE"] = "1"
def run(tag, **kw):
    wp = windows.WindowPanel(10_000, 5_040, engine=eng, host_staging=False, **kw)
    wp.fill_synthetic()
    print("c5", tag, wp.time_device(), flush=True)
    wp.close()
full = dict(kdj=(5, 9, 14, 60, 250), ext=(5, 20, 55, 250), atr=14)
run("planner", **full)
for U in (5, 4, 3, 2):
    for st in (2, 3):
        os.environ["PQB_WIN_UNITS"] = str(U); os.environ["PQB_WIN_STAGES"] = str(st)
        run("U=%d stages=%d" % (U, st), **full)
os.environ["PQB_WIN_UNITS"] = "3"; os.environ["PQB_WIN_STAGES"] = "2"
run("no250", kdj=(5, 9, 14, 60), ext=(5, 20, 55), atr=14)
run("only250", kdj=(250,), ext=(250,), atr=0)
os.environ["PQB_WIN_UNITS"] = "1"
run("kdj250 alone", kdj=(250,), ext=(), atr=0)
run("kdj9 alone", kdj=(9,), ext=(), atr=0)
run("wmd250 alone", kdj=(), ext=(250,), atr=0)
run("wmd20 alone", kdj=(), ext=(20,), atr=0)
PY
