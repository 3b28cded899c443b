#!/bin/bash
python scripts/bench_halted_symbols.py 2>&1 | cut -c1-200
PQB_BENCH_SYMBOLS=50000 python scripts/bench_halted_symbols.py 2>&1 | cut -c1-200
