#!/bin/bash
TAG=r03h
timeout 900 python -m pytest tests/test_gpu_nulls.py tests/test_gpu_columns.py tests/test_gpu_wide.py -m gpu -q -x > gpurun_out/pytest_nulls_$TAG.log 2>&1; grep -n "^E   \|passed\|failed" gpurun_out/pytest_nulls_$TAG.log | cut -c1-300 | head -20
for C in 0 1; do PQB_COMPACT_NULLS=$C timeout 600 python scripts/bench_halted_symbols.py 2>&1 | tee -a gpurun_out/halted_$TAG.log; done
timeout 300 python scripts/bench_nulls_mode.py 2>&1 | tee gpurun_out/nulls_$TAG.log | tail -5
