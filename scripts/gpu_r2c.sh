#!/bin/bash
TAG=${1:-r02c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_windows.py tests/test_gpu_longrows.py tests/test_gpu_ref_golden.py -m gpu -q 2>&1 | tail -30 | tee gpurun_out/pytest_gpu_$TAG.log
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/c35_$TAG.log
import sys, json, os
sys.path.insert(0, ".")
import bench, polars_quant_b200 as pq
from polars_quant_b200 import _native as NV
eng = pq.get_engine(0)
print(json.dumps(bench.bench_c3(pq, NV, eng, 6560.0)))
from polars_quant_b200 import longrows, windows
for tile in (2048, 4096, 8192):
    lp = longrows.LongPanel(500, 1_000_000, engine=eng, tile_bars=tile, host_staging=False)
    lp.fill_synthetic()
    print("c3 tile", tile, lp.time_device())
    lp.close()
for G in (3, 4):
    os.environ["PQB_WIN_GROUPS"] = str(G)
    print("c5 groups", G, json.dumps(bench.bench_c5(pq, NV, eng, 6560.0)))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c35_$TAG.csv python - <<'PY' > gpurun_out/ncu_c35_$TAG.log 2>&1
import sys
sys.path.insert(0, ".")
import polars_quant_b200 as pq
from polars_quant_b200 import longrows, windows
eng = pq.get_engine(0)
lp = longrows.LongPanel(500, 1_000_000, engine=eng, host_staging=False); lp.fill_synthetic(); lp.run(); lp.run(); lp.panel.sync(); lp.close()
wp = windows.WindowPanel(10_000, 5_040, engine=eng, host_staging=False); wp.fill_synthetic(); wp.run(); wp.run(); wp.panel.sync(); wp.close()
PY
grep -v "^==" gpurun_out/launches_c35_$TAG.csv | awk -F'","' '{print $5, $NF}' | tail -30
