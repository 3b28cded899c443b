#!/bin/bash
# optional groups alone on the big panel: field-sized stages + CTA width by waves (the wide kernel), against the eight-warp launch
timeout 900 python -m pytest tests/test_gpu_extras.py tests/test_gpu_ref_golden.py tests/test_gpu_split.py tests/test_gpu_plugin.py -q -m gpu -x 2>&1 | tail -3
echo "== auto"; PQB_PRINT_OCC=1 python scripts/probe_groups.py 2>&1 | grep -v "^\[pqb\]"
PQB_PRINT_OCC=1 python scripts/probe_groups.py 2>&1 | grep "^\[pqb\]" | sort | uniq -c
echo "== eight warps, full stages"; PQB_BASE_WARPS=8 PQB_BASE_SLIM=0 python scripts/probe_groups.py 2>&1 | tail -13
echo "== auto, 10,000"; python scripts/probe_groups.py 10000 2>&1 | tail -13
