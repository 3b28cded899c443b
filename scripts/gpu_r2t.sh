#!/bin/bash
TAG=${1:-r02t}
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_c4_$TAG.json 2> gpurun_out/bench_c4_$TAG.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_$TAG.json')); print({k: d[k] for k in ('value','ms_per_step','scaling')}); print(d['roofline']['frac'], d['e2e']); print(d['cpu_baseline']['value']); print({k:(v.get('kernel_ms'), v.get('frac'), v.get('error')) for k,v in d['other_workloads'].items()})"
tail -3 gpurun_out/bench_c4_$TAG.err
timeout 300 python scripts/bench_config1.py 2>&1 | tee gpurun_out/config1_$TAG.log | tail -5
timeout 300 python scripts/bench_nulls_mode.py 2>&1 | tee gpurun_out/nulls_$TAG.log | tail -12
