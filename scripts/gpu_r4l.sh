#!/bin/bash
for lib in polars_quant_b200/libpqb200.so build/libub4.so build/libub8.so; do
  echo "== $lib"
  PQB_LIB=$PWD/$lib python scripts/probe_partial.py 2>&1 | tail -12
  PQB_LIB=$PWD/$lib python scripts/probe_single_call.py 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k: v['c_abi_us'] for k,v in d.items()})"
done
