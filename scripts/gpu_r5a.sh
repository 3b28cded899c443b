#!/bin/bash
# AROON as block-decomposed window positions: parity (extras, reference-executed vectors, nulls), then the optional-group timings
timeout 900 python -m pytest tests/test_gpu_extras.py tests/test_gpu_ref_golden.py tests/test_gpu_nulls.py tests/test_gpu_split.py -q -m gpu 2>&1 | tail -8
timeout 600 python scripts/bench_next_rows.py groups 2>&1 | tail -20
