#!/bin/bash
TAG=r03j
timeout 900 python -m pytest tests/test_gpu_nulls.py -m gpu -q -x 2>&1 | tail -2
timeout 600 python scripts/bench_halted_symbols.py 2>&1 | tee -a gpurun_out/halted_$TAG.log
PQB_BENCH_SYMBOLS=50000 timeout 900 python scripts/bench_halted_symbols.py 2>&1 | tee -a gpurun_out/halted_$TAG.log
