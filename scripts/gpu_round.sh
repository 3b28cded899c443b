#!/bin/bash
# One GPU session: parity tests, benches, ncu launch list + captures of the fused kernel.
# Usage (under gpurun): bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_$TAG.log
python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/bench_c4_$TAG.json 2> gpurun_out/bench_c4_$TAG.err; tail -c 3000 gpurun_out/bench_c4_$TAG.json; tail -5 gpurun_out/bench_c4_$TAG.err
python bench.py --workload c2 --steps 20 --warmup 3 > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err; tail -c 3000 gpurun_out/bench_c2_$TAG.json; tail -5 gpurun_out/bench_c2_$TAG.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>&1; tail -c 1500 gpurun_out/bench_ref_$TAG.json
# launch list of the default bench command (every launch with its device time)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --workload c4 --steps 2 --warmup 3 --e2e-steps 1 --e2e-symbols 1664 --no-cpu > gpurun_out/ncu_list_$TAG.log 2>&1
# DRAM traffic of the fused kernel at the full config-4 size (two light passes)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:suite_fused -s 3 -c 1 --csv \
    --log-file gpurun_out/traffic_c4_$TAG.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_traffic_$TAG.log 2>&1
# full capture at config-4 row length with 3 resident blocks per SM (14,208 symbols: same steady state, 1/3.5 the footprint)
ncu --set full --clock-control none --import-source on -k regex:suite_fused -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --workload c4 --symbols 14208 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out
