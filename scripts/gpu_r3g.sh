#!/bin/bash
TAG=r03g
timeout 600 ncu --set full --clock-control none --import-source on -k suite_fused_kernel -s 3 -c 1 -f -o gpurun_out/prof_nulls_$TAG python scripts/prof_nulls.py > gpurun_out/ncu_nulls_$TAG.log 2>&1
tail -3 gpurun_out/ncu_nulls_$TAG.log
ls -la gpurun_out/prof_nulls_$TAG.ncu-rep
