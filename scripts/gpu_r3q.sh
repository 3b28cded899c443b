#!/bin/bash
for BM in 0 1; do
PQB_WIN_BLOCK_MAJOR=$BM timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:window_suite -s 2 -c 1 python - <<'PY' 2>&1 | grep -i "dram__\|gpu__time"
import sys
sys.path.insert(0, ".")
import polars_quant_b200 as pq
from polars_quant_b200 import windows
wp = windows.WindowPanel(10_000, 5_040, engine=pq.get_engine(0), host_staging=False); wp.fill_synthetic(); wp.run(); wp.run(); wp.run(); wp.panel.sync(); wp.close()
PY
PQB_WIN_BLOCK_MAJOR=$BM timeout 600 python - <<'PY' 2>&1
import sys, os
sys.path.insert(0, ".")
import polars_quant_b200 as pq
from polars_quant_b200 import windows
for U in (3, 4, 5):
    os.environ["PQB_WIN_UNITS"] = str(U)
    wp = windows.WindowPanel(10_000, 5_040, engine=pq.get_engine(0), host_staging=False); wp.fill_synthetic()
    print("block_major", os.environ["PQB_WIN_BLOCK_MAJOR"], "U", U, wp.time_device(warmup=2, iters=10)); wp.close()
PY
done
