#!/bin/bash
TAG=${1:-r02u}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_columns.py tests/test_gpu_plugin.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_wide_$TAG.log 2>&1; grep -n "^E   \|passed\|failed" gpurun_out/pytest_wide_$TAG.log | cut -c1-300 | head -20
timeout 300 python scripts/bench_wide.py --cpu 2>&1 | tee gpurun_out/wide_$TAG.log
