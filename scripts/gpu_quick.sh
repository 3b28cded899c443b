#!/bin/bash
# quick GPU check: bit-exactness report, parity tests, kernel-only benches
python scripts/bitexact_report.py 300 2520 2>&1 | awk "{print \$1, \$4, \$5}" | tr '\n' ';'; echo
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for w in c2 c4; do python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', 'ms %.3f'%d['roofline']['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], 'GB/s %.0f'%d['roofline']['achieved'])"; done
