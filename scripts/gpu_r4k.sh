#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_plugin.py tests/test_gpu_extras.py tests/test_gpu_nulls.py tests/test_gpu_candles.py -q -m gpu 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "single_column or small_panel or config1 or config2" 2>&1 | tail -3
echo "== mapped (default)"; python scripts/bench_config1.py 2>&1 | tail -3
echo "== copy engine (PQB_SINGLE_MAPPED=0)"; PQB_SINGLE_MAPPED=0 python scripts/bench_config1.py 2>&1 | tail -3
