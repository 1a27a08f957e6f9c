#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_premapped.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q --timeout 1200 -x > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2h_tests.log
for d in 1 0; do
  TH_CHAIN_DEFER=$d TH_CHAIN_STATS=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-culled --no-extras > gpurun_out/r2h_stats_defer$d.json 2> gpurun_out/r2h_stats_defer$d.txt
  echo "== TH_CHAIN_DEFER=$d"; python -c "
import json;d=json.loads(open('gpurun_out/r2h_stats_defer$d.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['ms_per_step_by_category'], d['roofline']['frac'])"
  grep "chain stats" gpurun_out/r2h_stats_defer$d.txt | tail -5 | cut -c1-2200
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1]);print('BENCH', d['value'], d['ms_per_step'], d['ms_per_step_by_category'], d['roofline']['frac'], d['e2e'], d['culled'])"
