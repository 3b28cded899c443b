#!/bin/bash
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu_$TAG.log
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/c3_$TAG.log
import sys, json
sys.path.insert(0, ".")
import bench, polars_quant_b200 as pq
from polars_quant_b200 import _native as NV
eng = pq.get_engine(0)
print(json.dumps(bench.bench_c3(pq, NV, eng, 6560.0)))
from polars_quant_b200 import longrows
for tile in (512, 2048, 4096):
    lp = longrows.LongPanel(500, 1_000_000, engine=eng, tile_bars=tile, host_staging=False)
    lp.fill_synthetic()
    print("tile", tile, lp.time_device())
    lp.close()
PY
