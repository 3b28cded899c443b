#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_nulls.py tests/test_gpu_columns.py tests/test_gpu_wide.py -q -m gpu -x 2>&1 | tail -3
python scripts/bench_nulls_mode.py 2>&1 | cut -c1-200; python scripts/bench_halted_symbols.py 2>&1 | cut -c1-200
PQB_BENCH_SYMBOLS=50000 python scripts/bench_halted_symbols.py 2>&1 | cut -c1-200
