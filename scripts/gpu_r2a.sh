#!/bin/bash
# Round 2, GPU session A: parity tests (incl. the reference-executed golden vectors), default bench, reference arm.
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_c4_$TAG.json 2> gpurun_out/bench_c4_$TAG.err; tail -c 6000 gpurun_out/bench_c4_$TAG.json; tail -5 gpurun_out/bench_c4_$TAG.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>&1; tail -c 1500 gpurun_out/bench_ref_$TAG.json
