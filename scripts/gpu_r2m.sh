#!/bin/bash
TAG=${1:-r02m}
mkdir -p gpurun_out
export PQB_WIN_UNITS=${2:-3} PQB_WIN_STAGES=2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:window_suite -s 1 -c 1 -f -o gpurun_out/prof_win_$TAG python - <<'PY' > gpurun_out/ncu_win_$TAG.log 2>&1
import sys, os
sys.path.insert(0, ".")
import polars_quant_b200 as pq
from polars_quant_b200 import windows
eng = pq.get_engine(0)
wp = windows.WindowPanel(10_000, 5_040, engine=eng, host_staging=False); wp.fill_synthetic(); wp.run(); wp.run(); wp.panel.sync(); wp.close()
PY
tail -3 gpurun_out/ncu_win_$TAG.log
ncu -i gpurun_out/prof_win_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_win_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_win_$TAG.ncu-rep --page source --csv > gpurun_out/prof_win_${TAG}_source.csv 2>/dev/null
ls -la gpurun_out | grep $TAG
