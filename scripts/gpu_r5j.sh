#!/bin/bash
# whole GPU suite on the final partial-suite launch policy + timings
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4
PQB_PRINT_OCC=1 python scripts/probe_partial.py 2>&1 | grep -v "^\[pqb\]" | tail -12
python scripts/bench_configs.py 2>&1 | tail -12 | cut -c1-400
