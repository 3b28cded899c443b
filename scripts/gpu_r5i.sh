#!/bin/bash
# producer placement: KDJ + ATR (two slots) and a four-slot suite at fixed CTA widths, then the automatic choice
for w in 3 4 5 6 7 8; do echo "== PQB_BASE_WARPS=$w"; PQB_BASE_WARPS=$w PQB_PRINT_OCC=1 python scripts/probe_occ.py kdj+atr willr+midprice sma+ema+rsi+macd+bbands 2>&1 | sort | uniq | tail -6; done
echo "== auto"; PQB_PRINT_OCC=1 python scripts/probe_occ.py ema rsi bbands kdj+atr willr+midprice obv+ad sma+ema+rsi+macd+bbands 2>&1 | sort | uniq
