#!/bin/bash
# round 2, visit a: every GPU test (nothing skipped), bench default (pre-mapped) and plain maps, chain wait stats
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 900 > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; cat gpurun_out/r2a_bench.json | cut -c1-1800; tail -3 gpurun_out/r2a_bench.err
timeout 600 python bench.py --steps 3 --warmup 3 --plain-maps --no-cpu-baseline > gpurun_out/r2a_bench_plain.json 2> gpurun_out/r2a_bench_plain.err; echo "plain rc=$?"; cut -c1-600 gpurun_out/r2a_bench_plain.json
TH_CHAIN_STATS=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-culled > gpurun_out/r2a_stats.json 2> gpurun_out/r2a_chain_stats.txt; echo "stats rc=$?"; grep "chain stats" gpurun_out/r2a_chain_stats.txt | tail -6
