#!/bin/bash
TAG=${1:-r02h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu_$TAG.log
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/c5_$TAG.log
import sys, json, os
sys.path.insert(0, ".")
import polars_quant_b200 as pq
eng = pq.get_engine(0)
from polars_quant_b200 import windows, longrows
lp = longrows.LongPanel(500, 1_000_000, engine=eng, host_staging=False); lp.fill_synthetic(); print("c3", lp.time_device()); lp.close()
def run(tag, **kw):
    wp = windows.WindowPanel(10_000, 5_040, engine=eng, host_staging=False, **kw)
    wp.fill_synthetic()
    print("c5", tag, wp.time_device())
    wp.close()
for deal in (1, 0):
    for G, U, SM in ((2, 5, 32), (2, 5, 64), (3, 4, 32), (3, 5, 32), (3, 4, 64)):
        os.environ["PQB_WIN_DEAL"] = str(deal); os.environ["PQB_WIN_GROUPS"] = str(G); os.environ["PQB_WIN_UNITS"] = str(U); os.environ["PQB_WIN_SMEM_MAX"] = str(SM)
        run("deal=%d G=%d U=%d smem_max=%d" % (deal, G, U, SM), kdj=(5, 9, 14, 60, 250), ext=(5, 20, 55, 250), atr=14)
PY
