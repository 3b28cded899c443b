#!/bin/bash
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/c5dbg_$TAG.log
import sys, json, os
sys.path.insert(0, ".")
import polars_quant_b200 as pq
eng = pq.get_engine(0)
from polars_quant_b200 import windows
def run(tag, **kw):
    wp = windows.WindowPanel(10_000, 5_040, engine=eng, host_staging=False, **kw)
    wp.fill_synthetic()
    print("c5", tag, wp.time_device())
    wp.close()
os.environ["PQB_WIN_GROUPS"] = "1"
for dbg in (0, 1, 2, 4, 3, 7):
    os.environ["PQB_WIN_DBG"] = str(dbg)
    run("kdj250 alone dbg=%d" % dbg, kdj=(250,), ext=(), atr=0)
os.environ["PQB_WIN_DBG"] = "0"
run("wmd250 alone", kdj=(), ext=(250,), atr=0)
run("kdj60 global alone", kdj=(60,), ext=(), atr=0)
os.environ["PQB_WIN_SMEM_MAX"] = "64"
run("kdj60 smem alone", kdj=(60,), ext=(), atr=0)
PY
timeout 600 python -m pytest tests/test_gpu_ref_golden.py tests/test_gpu_extras.py -m gpu -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu_$TAG.log
