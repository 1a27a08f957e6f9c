#!/bin/bash
# round 2, visit b: whole GPU suite incl. the genuine reference in torch-CUDA, then the full bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s --timeout 1500 > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2b_tests.log
grep -n "full frame\|make_renderer\|knife-edge\] full" gpurun_out/r2b_tests.log | cut -c1-400
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; cut -c1-3000 gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err; echo "ref rc=$?"; cut -c1-800 gpurun_out/r2b_bench_ref.json
