#!/bin/bash
# usage: gpurun -- bash scripts/gpu_r2k.sh <tag> <commit>
TAG=${1:-r02k}; COMMIT=${2:-unknown}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_full_$TAG.log 2>&1; grep -n "^E   \|passed\|failed" gpurun_out/pytest_gpu_full_$TAG.log | cut -c1-400 | head -40
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
run("default", kdj=(5, 9, 14, 60, 250), ext=(5, 20, 55, 250), atr=14)
run("no250", kdj=(5, 9, 14, 60), ext=(5, 20, 55), atr=14)
run("only250", kdj=(250,), ext=(250,), atr=0)
PY
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
timeout 600 python bench.py > gpurun_out/bench_c4_$TAG.json 2> gpurun_out/bench_c4_$TAG.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_$TAG.json')); print({k: d[k] for k in ('value','ms_per_step','scaling')}); print(d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline']['value']); print({k:(v.get('kernel_ms'), v.get('frac'), v.get('error')) for k,v in d['other_workloads'].items()})"
tail -3 gpurun_out/bench_c4_$TAG.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_bench_c4.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
python scripts/ncu_traffic.py gpurun_out/${TAG}_launches_bench_c4.csv $COMMIT gpurun_out/${TAG}_launch_summary.txt
