#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_nulls.py tests/test_gpu_columns.py tests/test_gpu_wide.py -q -m gpu -x 2>&1 | tail -2
PQB_BENCH_SYMBOLS=8192 python scripts/bench_halted_symbols.py 2>&1 | cut -c1-200
