#!/bin/bash
for D in 0 40 200; do echo "== delay $D us"; PQB_COMPACT_DELAY_US=$D PQB_BENCH_SYMBOLS=50000 timeout 900 python scripts/bench_halted_symbols.py 2>&1 | grep "symbols (\|no nulls" | cut -c1-200; done
echo "== 20000 symbols, delay 40"; python scripts/bench_halted_symbols.py 2>&1 | cut -c1-200
python -m pytest tests/test_gpu_nulls.py -m gpu -q 2>&1 | tail -1
