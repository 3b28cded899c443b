#!/bin/bash
# the compile-time-specialised null-aware kernel (<FULLS, NULLS>) with fast stages: parity, then timings against the general one
timeout 1200 python -m pytest tests/test_gpu_nulls.py tests/test_gpu_columns.py tests/test_gpu_wide.py tests/test_gpu_plugin.py -q -m gpu -x 2>&1 | tail -8
echo "== <true, true> (default)"; python scripts/bench_nulls_mode.py 2>&1 | cut -c1-220; python scripts/bench_halted_symbols.py 2>&1 | cut -c1-200
echo "== PQB_NULLS_FULLS=0"; PQB_NULLS_FULLS=0 python scripts/bench_nulls_mode.py 2>&1 | cut -c1-220
echo "== 50,000"; PQB_BENCH_SYMBOLS=50000 python scripts/bench_halted_symbols.py 2>&1 | cut -c1-200
