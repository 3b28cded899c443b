#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_nulls.py tests/test_gpu_columns.py tests/test_gpu_wide.py -m gpu -q -x 2>&1 | tail -3
PQB_BENCH_SYMBOLS=50000 python scripts/bench_halted_symbols.py 2>&1 | cut -c1-190
python scripts/bench_halted_symbols.py 2>&1 | cut -c1-190
