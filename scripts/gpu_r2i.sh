#!/bin/bash
TAG=${1:-r02i}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_full_$TAG.log 2>&1; grep -n "^E   \|passed\|failed" gpurun_out/pytest_gpu_full_$TAG.log | cut -c1-400 | head -60
timeout 300 python scripts/bench_wide.py --cpu 2>&1 | tee gpurun_out/wide_$TAG.log
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/c5_$TAG.log
import sys, json, os
sys.path.insert(0, ".")
import polars_quant_b200 as pq
eng = pq.get_engine(0)
from polars_quant_b200 import windows, longrows
lp = longrows.LongPanel(500, 1_000_000, engine=eng, host_staging=False); lp.fill_synthetic(); print("c3", lp.time_device()); lp.close()
def run(tag, **kw):
    wp = windows.WindowPanel(10_000, 5_040, engine=eng, host_staging=False, **kw)
    wp.fill_synthetic()
    print("c5", tag, wp.time_device())
    wp.close()
for G, U, SM in ((2, 5, 32), (3, 4, 32), (3, 5, 32), (4, 3, 32), (2, 5, 20)):
    os.environ["PQB_WIN_GROUPS"] = str(G); os.environ["PQB_WIN_UNITS"] = str(U); os.environ["PQB_WIN_SMEM_MAX"] = str(SM)
    run("G=%d U=%d smem_max=%d" % (G, U, SM), kdj=(5, 9, 14, 60, 250), ext=(5, 20, 55, 250), atr=14)
PY
timeout 600 python bench.py > gpurun_out/bench_c4_$TAG.json 2> gpurun_out/bench_c4_$TAG.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_$TAG.json')); print({k: d[k] for k in ('value','ms_per_step','scaling')}); print(d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline']['value']); print({k:(v.get('kernel_ms'), v.get('frac'), v.get('error')) for k,v in d['other_workloads'].items()})"
tail -3 gpurun_out/bench_c4_$TAG.err
