#!/bin/bash
# PQB_IND_FASTK opt-in (ABI 7): shims + extras + plugin parity, then the optional-group timings on the every-plane panel
timeout 900 python -m pytest tests/test_gpu_extras.py tests/test_gpu_plugin.py tests/test_gpu_ref_golden.py -q -m gpu 2>&1 | tail -8
timeout 600 python scripts/bench_next_rows.py groups 2>&1 | grep -v "alone" | cut -c1-260
