#!/bin/bash
for O in 0 1 2; do echo "== order $O"; PQB_COMPACT_ORDER=$O PQB_BENCH_SYMBOLS=50000 timeout 900 python scripts/bench_halted_symbols.py 2>&1 | grep "500 symbols\|50 symbols"; done
