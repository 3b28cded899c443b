#!/bin/bash
# compute-sanitizer memcheck over the launches changed in r05: AROON's block arrays, slim-stage / narrow partial-suite CTAs, the mapped single-column path
export PQB_HOST_POOL_MB=0
for t in "tests/test_gpu_extras.py -k 'aroon or fastk or optional_groups'" "tests/test_gpu_parity.py -k 'many_wave or partial_suites or single_column'" \
         "tests/test_gpu_plugin.py" "tests/test_gpu_ref_golden.py"; do
  tag=$(echo "$t" | tr -c 'a-zA-Z0-9\n' '_' | cut -c1-60)
  echo "== memcheck $t" | tee -a gpurun_out/r05k_memcheck.txt
  eval timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r05k_mc_$tag.log python -m pytest $t -x -q -m gpu 2>&1 | tail -3 | tee -a gpurun_out/r05k_memcheck.txt
  echo "exit ${PIPESTATUS[0]}" | tee -a gpurun_out/r05k_memcheck.txt
  grep -h "ERROR SUMMARY\|Invalid\|out of bounds" gpurun_out/r05k_mc_$tag.log | sort | uniq -c | head -5 | tee -a gpurun_out/r05k_memcheck.txt
done
