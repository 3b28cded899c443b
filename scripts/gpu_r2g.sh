#!/bin/bash
TAG=${1:-r02g}
mkdir -p gpurun_out
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/c5_$TAG.log
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
run("kdj250 alone", kdj=(250,), ext=(), atr=0)
run("wmd250 alone", kdj=(), ext=(250,), atr=0)
os.environ["PQB_WIN_SMEM_MAX"] = "64"
run("kdj60 smem alone", kdj=(60,), ext=(), atr=0)
os.environ["PQB_WIN_SMEM_MAX"] = "32"
for G, U in ((2, 5), (3, 4), (3, 5), (4, 3)):
    os.environ["PQB_WIN_GROUPS"] = str(G); os.environ["PQB_WIN_UNITS"] = str(U)
    run("all G=%d U=%d" % (G, U), kdj=(5, 9, 14, 60, 250), ext=(5, 20, 55, 250), atr=14)
PY
timeout 900 python -m pytest tests/test_gpu_ref_golden.py tests/test_gpu_extras.py tests/test_gpu_windows.py -m gpu -q 2>&1 | tail -60 | tee gpurun_out/pytest_gpu_$TAG.log
