#!/bin/bash
# compute-sanitizer memcheck over the null-aware launches of r05 (fast stages, <FULLS, NULLS>, dispatch on the third stream, compaction of partial suites)
export PQB_HOST_POOL_MB=0
for t in "tests/test_gpu_nulls.py -k 'fast_stages or compaction or function_by_function'" "tests/test_gpu_columns.py" "tests/test_gpu_wide.py"; do
  tag=$(echo "$t" | tr -c 'a-zA-Z0-9\n' '_' | cut -c1-60)
  echo "== memcheck $t" | tee -a gpurun_out/r05u_memcheck.txt
  eval timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r05u_mc_$tag.log python -m pytest $t -x -q -m gpu 2>&1 | tail -3 | tee -a gpurun_out/r05u_memcheck.txt
  echo "exit ${PIPESTATUS[0]}" | tee -a gpurun_out/r05u_memcheck.txt
  grep -h "ERROR SUMMARY\|Invalid\|out of bounds" gpurun_out/r05u_mc_$tag.log | sort | uniq -c | head -5 | tee -a gpurun_out/r05u_memcheck.txt
done
