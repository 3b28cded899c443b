#!/bin/bash
TAG=r03i
for C in 1; do PQB_COMPACT_NULLS=$C timeout 600 python scripts/bench_halted_symbols.py 2>&1 | tee -a gpurun_out/halted_$TAG.log; done
PQB_BENCH_SYMBOLS=50000 timeout 900 python scripts/bench_halted_symbols.py 2>&1 | tee -a gpurun_out/halted_$TAG.log
PQB_COMPACT_NULLS=0 PQB_BENCH_SYMBOLS=50000 timeout 900 python scripts/bench_halted_symbols.py 2>&1 | tee -a gpurun_out/halted_$TAG.log
python scripts/shape_sweep.py 5000x2520 5500x2520 2>&1 | tee -a gpurun_out/halted_$TAG.log
