#!/bin/bash
TAG=${1:-r02d}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_windows.py tests/test_gpu_longrows.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_$TAG.log
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/c35_$TAG.log
import sys, json, os
sys.path.insert(0, ".")
import bench, polars_quant_b200 as pq
from polars_quant_b200 import _native as NV
eng = pq.get_engine(0)
from polars_quant_b200 import longrows, windows
for tile in (1024, 2048, 4096):
    lp = longrows.LongPanel(500, 1_000_000, engine=eng, tile_bars=tile, host_staging=False)
    lp.fill_synthetic()
    print("c3 tile", tile, lp.time_device())
    lp.close()
for G, SM, U in ((2, 32, 5), (2, 64, 5), (3, 32, 4), (3, 64, 4), (4, 64, 3), (2, 20, 5), (3, 20, 4)):
    os.environ["PQB_WIN_GROUPS"] = str(G); os.environ["PQB_WIN_SMEM_MAX"] = str(SM); os.environ["PQB_WIN_UNITS"] = str(U)
    r = bench.bench_c5(pq, NV, eng, 6560.0)
    print("c5 groups", G, "smem_max", SM, "units", U, r["kernel_ms"], r["frac"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c3_$TAG.csv python - <<'PY' > gpurun_out/ncu_c3_$TAG.log 2>&1
import sys
sys.path.insert(0, ".")
import polars_quant_b200 as pq
from polars_quant_b200 import longrows
eng = pq.get_engine(0)
lp = longrows.LongPanel(500, 1_000_000, engine=eng, host_staging=False); lp.fill_synthetic(); lp.run(); lp.run(); lp.panel.sync(); lp.close()
PY
grep -v "^==" gpurun_out/launches_c3_$TAG.csv | awk -F'","' '{print $5, $NF}' | tail -12
