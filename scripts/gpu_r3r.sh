#!/bin/bash
python scripts/bench_touching_closes.py 2>&1 | tee gpurun_out/touching_r03r.log
python - <<'PY' 2>&1 | tee -a gpurun_out/touching_r03r.log
# the small-panel (pipelined) variant and the window kernel on touching closes
import sys, json
import numpy as np
sys.path.insert(0, ".")
import polars_quant_b200 as pq
from polars_quant_b200 import _native as N
for S, NB in ((5000, 2520),):
    p = pq.Panel(S, NB, engine=pq.get_engine(0))
    p.fill_synthetic(seed=5, to_host=True)
    prm = N.default_params()
    tot, fused, nl = p.time_device(prm, warmup=2, iters=10)
    print(json.dumps({"config": "config 2, synthetic", "kernel_ms": fused / 10}))
    rng = np.random.default_rng(1)
    c, h, l = p.host_field("close"), p.host_field("high"), p.host_field("low")
    m = rng.random((S, NB)) < 0.10; h[:, :NB][m] = c[:, :NB][m]
    m = rng.random((S, NB)) < 0.10; l[:, :NB][m] = c[:, :NB][m]
    p.upload()
    tot, fused, nl = p.time_device(prm, warmup=2, iters=10)
    print(json.dumps({"config": "config 2, close == high on 10% and close == low on 10% of the bars", "kernel_ms": fused / 10}))
PY
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "closes_at or flat_and_tied or config2" 2>&1 | tail -2
python -m pytest tests -m gpu -q -k "selftest or divsqrt or windows or closes_at or config5" 2>&1 | tail -2
python - <<'PY'
import sys, os
sys.path.insert(0, ".")
import polars_quant_b200 as pq
from polars_quant_b200 import windows
wp = windows.WindowPanel(10_000, 5_040, engine=pq.get_engine(0), host_staging=False); wp.fill_synthetic()
print("c5", wp.time_device(warmup=2, iters=10)); wp.close()
PY
