#!/bin/bash
# where do the compact tail CTAs land? (a -DPQB_DEBUG_SMID build) + timings of the 149..296-block regime
export PQB_LIB=$PWD/build/libsmid.so
for cfg in "PQB_SMALL_BLOCKS=200 PQB_TAIL_PARTS=3" "PQB_SMALL_BLOCKS=200 PQB_TAIL_PARTS=2" "PQB_SMALL_BLOCKS=230 PQB_TAIL_PARTS=2" "PQB_SMALL_BLOCKS=200 PQB_TAIL_PARTS=5"; do
  echo "== $cfg"
  env $cfg python scripts/shape_sweep.py 5504x5040 6272x5040 7104x5040 2>&1 | grep -v "^\[pqb\] SMs" | head -5
  env $cfg python scripts/shape_sweep.py 6272x5040 2>&1 | grep "SMs by" | sort | uniq -c | sort -rn | head -4
  env $cfg python scripts/shape_sweep.py 7104x5040 2>&1 | grep "SMs by" | sort | uniq -c | sort -rn | head -3
done
unset PQB_LIB
echo "== product lib"
python scripts/shape_sweep.py 5000x2520 5504x5040 6272x5040 7104x5040
