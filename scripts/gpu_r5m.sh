#!/bin/bash
# fast stages of the null-aware kernel: parity (the whole null suite + columns / wide, which carry nulls), then timings both ways
timeout 1200 python -m pytest tests/test_gpu_nulls.py tests/test_gpu_columns.py tests/test_gpu_wide.py tests/test_gpu_plugin.py -q -m gpu -x 2>&1 | tail -8
echo "== fast stages (default)"; python scripts/bench_nulls_mode.py 2>&1 | cut -c1-220; python scripts/bench_halted_symbols.py 2>&1 | cut -c1-200
echo "== PQB_NULLS_FAST=0"; PQB_NULLS_FAST=0 python scripts/bench_nulls_mode.py 2>&1 | cut -c1-220; PQB_NULLS_FAST=0 python scripts/bench_halted_symbols.py 2>&1 | cut -c1-200
