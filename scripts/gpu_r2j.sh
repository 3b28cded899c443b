#!/bin/bash
TAG=${1:-r02j}
mkdir -p gpurun_out
timeout 300 python scripts/bench_wide.py --cpu 2>&1 | tee gpurun_out/wide_$TAG.log
timeout 900 python -m pytest tests/test_gpu_ref_golden.py tests/test_gpu_parity.py tests/test_gpu_wide.py tests/test_gpu_extras.py -m gpu -q > gpurun_out/pytest_gpu_part_$TAG.log 2>&1; grep -n "^E   \|passed\|failed" gpurun_out/pytest_gpu_part_$TAG.log | cut -c1-400 | head -40
