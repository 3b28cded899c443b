#!/bin/bash
# final shape policy (two-warp CTAs rotate role / producer in pairs): parity of every partial-suite / optional-group path, then timings
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_extras.py tests/test_gpu_split.py tests/test_gpu_plugin.py tests/test_gpu_ref_golden.py -q -m gpu -x 2>&1 | tail -3
python scripts/probe_occ.py ema rsi bbands kdj+atr sma macd obv+ad 2>&1 | tail -7
python scripts/probe_groups.py 2>&1 | tail -13
