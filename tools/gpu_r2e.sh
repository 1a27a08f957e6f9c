#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -x > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2e_tests.log
for mode in 0 1; do
  TH_CHAIN_INPLACE_MIX=$mode TH_CHAIN_STATS=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-culled --no-extras > gpurun_out/r2e_stats_inplace$mode.json 2> gpurun_out/r2e_stats_inplace$mode.txt
  echo "== TH_CHAIN_INPLACE_MIX=$mode"; python -c "
import json;d=json.loads(open('gpurun_out/r2e_stats_inplace$mode.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['ms_per_step_by_category'], d['roofline']['frac'])"
  grep "chain stats" gpurun_out/r2e_stats_inplace$mode.txt | tail -5 | cut -c1-2200
done
