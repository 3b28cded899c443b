#!/bin/bash
for CFG in "1 40" "1 0" "0 0"; do set -- $CFG
echo "== exclusive $1 delay $2 (timeline build)"
PQB_COMPACT_EXCLUSIVE=$1 PQB_COMPACT_DELAY_US=$2 PQB_LIB=$PWD/build_variants/libpqb200_tl.so PQB_BENCH_SYMBOLS=50000 PQB_HALTED=500 python scripts/prof_halted.py 2>&1 | grep timeline | tail -1
done
for CFG in "1 40" "1 0"; do set -- $CFG
echo "== exclusive $1 delay $2"
PQB_COMPACT_EXCLUSIVE=$1 PQB_COMPACT_DELAY_US=$2 PQB_BENCH_SYMBOLS=50000 python scripts/bench_halted_symbols.py 2>&1 | grep "symbols (\|no nulls" | cut -c1-190
PQB_COMPACT_EXCLUSIVE=$1 PQB_COMPACT_DELAY_US=$2 python scripts/bench_halted_symbols.py 2>&1 | grep "symbols (\|no nulls" | cut -c1-190
done
