#!/bin/bash
# ncu --set full of the specialised null-aware kernel (fast stages) on 8,192 x 5,040 with a 3-bar halt in every symbol
timeout 900 ncu --set full --clock-control none --import-source on -k regex:suite_fused_kernel -s 1 -c 1 -f -o gpurun_out/prof_nulls_fast_r05z python - <<'PY' > gpurun_out/ncu_nulls_fast_r05z.log 2>&1
import sys
sys.path.insert(0, ".")
import numpy as np
import polars_quant_b200 as pq
from polars_quant_b200 import _native as N
S, NB = 8192, 5040
p = pq.Panel(S, NB, engine=pq.get_engine(0))
p.fill_synthetic(seed=5, to_host=True)
ok = np.ones(NB, dtype=bool); ok[2000:2003] = False
bits = np.packbits(ok, bitorder="little")
for s in range(S):
    p.set_column(s, "close", np.ascontiguousarray(p.host_field("close")[s]), validity=bits)
p.upload()
prm = N.default_params()
p.run(prm); p.run(prm); p.sync()
PY
tail -2 gpurun_out/ncu_nulls_fast_r05z.log
python scripts/ncu_summary.py gpurun_out/prof_nulls_fast_r05z.ncu-rep gpurun_out/ncu_nulls_fast_r05z.txt
