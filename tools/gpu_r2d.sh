#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_prologue.py tests/test_gpu_reference_cuda.py -m gpu -q -s --timeout 900 > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2d_tests.log
grep -n "full frame\] raw\|knife-edge\] full" gpurun_out/r2d_tests.log | cut -c1-400
for dbg in 0 1; do
  TH_CHAIN_DBG=$dbg TH_CHAIN_STATS=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-culled --no-extras > gpurun_out/r2d_stats_dbg$dbg.json 2> gpurun_out/r2d_stats_dbg$dbg.txt
  echo "== TH_CHAIN_DBG=$dbg"; python -c "
import json;d=json.loads(open('gpurun_out/r2d_stats_dbg$dbg.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['ms_per_step_by_category'])"
  grep "chain stats" gpurun_out/r2d_stats_dbg$dbg.txt | tail -5 | cut -c1-1500
done
