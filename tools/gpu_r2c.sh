#!/bin/bash
# round 2, visit c: failing tests of visit b + the new prologue tests, then the whole suite
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_prologue.py tests/test_gpu_reference_cuda.py tests/test_gpu_renderer_plugin.py -m gpu -q -s --timeout 900 > gpurun_out/r2c_tests1.log 2>&1; echo "tests1 rc=$?"; tail -8 gpurun_out/r2c_tests1.log
grep -n "full frame\]\|make_renderer\]" gpurun_out/r2c_tests1.log | cut -c1-420
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 --deselect tests/test_gpu_reference_cuda.py > gpurun_out/r2c_tests2.log 2>&1; echo "tests2 rc=$?"; tail -8 gpurun_out/r2c_tests2.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c_bench.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["ms_per_step_by_category"])
print(json.dumps(d.get("torch_cuda_baseline"))[:900])
print(json.dumps(d.get("plugin"))[:2500])
PY
tail -3 gpurun_out/r2c_bench.err
