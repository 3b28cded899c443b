#!/bin/bash
# kernel-only C2/C4 timing of alternative builds of libpqb200.so: bash scripts/variants.sh build/libA.so build/libB.so ...
for lib in "$@"; do
  for w in c2 c4; do
    PQB_LIB=$PWD/$lib python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib $w', 'ms %.3f'%d['roofline']['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], 'GB/s %.0f'%d['roofline']['achieved'])"
  done
done
