#!/bin/bash
# ncu --set full of a single-indicator launch (EMA alone, the partial-suite kernel) at 50,000 x 5,040
timeout 900 ncu --set full --clock-control none --import-source on -k suite_fused_kernel -s 1 -c 1 -f -o gpurun_out/prof_ema_alone_r06c python - <<'PY' > gpurun_out/ncu_ema_alone_r06c.log 2>&1
import sys
sys.path.insert(0, ".")
import polars_quant_b200 as pq
from polars_quant_b200 import _native as N
p = pq.Panel(50000, 5_040, engine=pq.get_engine(0), host_staging=False)
p.fill_synthetic(seed=1, sigma=0.02)
prm = N.default_params(indicators=N.IND["ema"])
p.run(prm); p.run(prm); p.sync()
PY
tail -2 gpurun_out/ncu_ema_alone_r06c.log
python scripts/ncu_summary.py gpurun_out/prof_ema_alone_r06c.ncu-rep gpurun_out/ncu_ema_alone_r06c.txt
