#!/bin/bash
timeout 900 ncu --set full --clock-control none --import-source on -k suite_fused_kernel -s 1 -c 1 -f -o gpurun_out/prof_c4_r04h_${S:-50000} python - <<'PY' > gpurun_out/ncu_c4_r04h_${S:-50000}.log 2>&1
import sys
sys.path.insert(0, ".")
import polars_quant_b200 as pq
from polars_quant_b200 import _native as N
p = pq.Panel(int(__import__("os").environ.get("S", 50000)), 5_040, engine=pq.get_engine(0), host_staging=False)
p.fill_synthetic(seed=1, sigma=0.02)
prm = N.default_params()
p.run(prm); p.run(prm); p.sync()
PY
tail -2 gpurun_out/ncu_c4_r04h_${S:-50000}.log
