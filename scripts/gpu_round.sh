#!/bin/bash
# One GPU session: parity tests, benches, ncu launch list + full capture of the fused kernel.
# Usage (under gpurun): bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_$TAG.log
python bench.py --workload c2 --steps 20 --warmup 3 > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err; tail -c 3000 gpurun_out/bench_c2_$TAG.json; tail -5 gpurun_out/bench_c2_$TAG.err
python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/bench_c4_$TAG.json 2> gpurun_out/bench_c4_$TAG.err; tail -c 3000 gpurun_out/bench_c4_$TAG.json; tail -5 gpurun_out/bench_c4_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --workload c2 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_list_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:suite_fused -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --workload c2 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out
for V in polars_quant_b200/libpqb200_*.so; do
  [ -f "$V" ] || continue
  echo "== variant: $V =="
  PQB_LIB=$PWD/$V python bench.py --workload c4 --steps 10 --warmup 3 --no-e2e --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('variant c4', d['value'], d['roofline']['frac'])"
  PQB_LIB=$PWD/$V python bench.py --workload c2 --steps 20 --warmup 3 --no-e2e --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('variant c2', d['value'], d['roofline']['frac'])"
done
