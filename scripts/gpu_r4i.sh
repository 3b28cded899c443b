#!/bin/bash
# compute-sanitizer memcheck over this round's new launches (window kernel, symbol compaction, SPLIT0, compact tail, long rows),
# then ncu --set full of the nine-warp small-panel kernel on config 2 (5,000 x 2,520)
export PQB_HOST_POOL_MB=0
for t in "tests/test_gpu_windows.py -k small_panels" "tests/test_gpu_windows.py -k nan_highs" \
         "tests/test_gpu_nulls.py -k compaction" "tests/test_gpu_nulls.py -k 'reused or function_by_function or single_column'" \
         "tests/test_gpu_longrows.py -k 'ema_set or short_rows'" "tests/test_gpu_split.py" "tests/test_gpu_wide.py"; do
  tag=$(echo "$t" | tr -c 'a-zA-Z0-9\n' '_' | cut -c1-60)
  echo "== memcheck $t" | tee -a gpurun_out/r04i_memcheck.txt
  eval timeout 700 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r04i_mc_$tag.log python -m pytest $t -x -q -m gpu 2>&1 | tail -3 | tee -a gpurun_out/r04i_memcheck.txt
  echo "exit ${PIPESTATUS[0]}" | tee -a gpurun_out/r04i_memcheck.txt
  grep -h "ERROR SUMMARY\|Invalid\|out of bounds" gpurun_out/r04i_mc_$tag.log | sort | uniq -c | head -5 | tee -a gpurun_out/r04i_memcheck.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k suite_fused_kernel -s 2 -c 2 -f -o gpurun_out/prof_c2_r04i python - <<'PY' > gpurun_out/ncu_c2_r04i.log 2>&1
import sys
sys.path.insert(0, ".")
import polars_quant_b200 as pq
from polars_quant_b200 import _native as N
p = pq.Panel(5_000, 2_520, engine=pq.get_engine(0), host_staging=False)
p.fill_synthetic(seed=1, sigma=0.02)
prm = N.default_params()
p.run(prm); p.run(prm); p.sync()
PY
tail -2 gpurun_out/ncu_c2_r04i.log
