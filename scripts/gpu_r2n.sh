#!/bin/bash
TAG=${1:-r02n}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_windows.py -m gpu -q -x > gpurun_out/pytest_win_$TAG.log 2>&1; grep -n "^E   \|passed\|failed" gpurun_out/pytest_win_$TAG.log | cut -c1-300 | head -20
timeout 900 python - <<'PY' 2>&1 | tee gpurun_out/c5_$TAG.log
import sys, json, os
sys.path.insert(0, ".")
import polars_quant_b200 as pq
eng = pq.get_engine(0)
from polars_quant_b200 import windows
os.environ["PQB_WIN_VERBOSE"] = "1"
def run(tag, S=10_000, **kw):
    wp = windows.WindowPanel(S, 5_040, engine=eng, host_staging=False, **kw)
    wp.fill_synthetic()
    print("c5", tag, wp.time_device(), flush=True)
    wp.close()
full = dict(kdj=(5, 9, 14, 60, 250), ext=(5, 20, 55, 250), atr=14)
run("planner", **full)
for U in (4, 3, 2, 1):
    for deal in (1, 0):
        os.environ["PQB_WIN_UNITS"] = str(U); os.environ["PQB_WIN_DEAL"] = str(deal)
        run("U=%d deal=%d" % (U, deal), **full)
os.environ["PQB_WIN_DEAL"] = "1"
# contention curve: n identical-cost KDJ units, one warp per CTA
os.environ["PQB_WIN_UNITS"] = "1"
for n in (1, 2, 3, 4, 5, 6):
    run("kdj x%d U=1" % n, kdj=tuple(range(9, 9 + n)), ext=(), atr=0)
os.environ["PQB_WIN_UNITS"] = "3"
run("kdj x6 U=3", kdj=tuple(range(9, 15)), ext=(), atr=0)
os.environ["PQB_WIN_UNITS"] = "6"
run("kdj x6 U=6", kdj=tuple(range(9, 15)), ext=(), atr=0)
os.environ["PQB_WIN_UNITS"] = "1"
run("kdj x6 U=1 4736 symbols", S=4736, kdj=tuple(range(9, 15)), ext=(), atr=0)
run("kdj x3 U=1 4736 symbols", S=4736, kdj=tuple(range(9, 12)), ext=(), atr=0)
run("kdj x1 U=1 4736 symbols", S=4736, kdj=tuple(range(9, 10)), ext=(), atr=0)
PY
