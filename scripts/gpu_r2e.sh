#!/bin/bash
TAG=${1:-r02e}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_$TAG.log
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/c35_$TAG.log
import sys, json, os
sys.path.insert(0, ".")
import bench, polars_quant_b200 as pq
from polars_quant_b200 import _native as NV
eng = pq.get_engine(0)
from polars_quant_b200 import longrows, windows
for tile in (1024, 2048):
    lp = longrows.LongPanel(500, 1_000_000, engine=eng, tile_bars=tile, host_staging=False)
    lp.fill_synthetic()
    print("c3 tile", tile, lp.time_device())
    lp.close()
def run(tag, **kw):
    wp = windows.WindowPanel(10_000, 5_040, engine=eng, host_staging=False, **kw)
    wp.fill_synthetic()
    print("c5", tag, wp.time_device())
    wp.close()
os.environ["PQB_WIN_GROUPS"] = "2"
run("all", kdj=(5, 9, 14, 60, 250), ext=(5, 20, 55, 250), atr=14)
run("no250", kdj=(5, 9, 14, 60), ext=(5, 20, 55), atr=14)
run("only250", kdj=(250,), ext=(250,), atr=0)
run("only60/55 global", kdj=(60,), ext=(55,), atr=0)
os.environ["PQB_WIN_SMEM_MAX"] = "64"
run("only60/55 smem", kdj=(60,), ext=(55,), atr=0)
run("small smem only", kdj=(5, 9, 14), ext=(5, 20), atr=14)
os.environ["PQB_WIN_SMEM_MAX"] = "32"
os.environ["PQB_WIN_GROUPS"] = "1"
run("kdj9 alone", kdj=(9,), ext=(), atr=0)
run("wmd20 alone", kdj=(), ext=(20,), atr=0)
run("atr alone", kdj=(), ext=(), atr=14)
PY
ncu --set full --clock-control none --import-source on -k regex:window_suite -s 1 -c 1 -f -o gpurun_out/prof_win_$TAG python - <<'PY' > gpurun_out/ncu_win_$TAG.log 2>&1
import sys, os
sys.path.insert(0, ".")
os.environ["PQB_WIN_GROUPS"] = "2"
import polars_quant_b200 as pq
from polars_quant_b200 import windows
eng = pq.get_engine(0)
wp = windows.WindowPanel(10_000, 5_040, engine=eng, host_staging=False); wp.fill_synthetic(); wp.run(); wp.run(); wp.panel.sync(); wp.close()
PY
ncu -i gpurun_out/prof_win_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_win_${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
