#!/bin/bash
# compute-sanitizer memcheck + racecheck on the final kernels (smoke: 24x24x16 dense + culled through the pre-mapped
# chain program), and on the TMEM-mix / deferred-tail variants
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  for cfg in "0 inplace" "0 tmem" "1 inplace"; do
    set -- $cfg
    echo "===== compute-sanitizer --tool $tool  TH_CHAIN_DEFER=$1 TH_CHAIN_MIX=$2" >> gpurun_out/$TAG_sanitizer_$tool.txt
    TH_CHAIN_DEFER=$1 TH_CHAIN_MIX=$2 timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^$" | tail -25 >> gpurun_out/${TAG}_sanitizer_$tool.txt
  done
done
tail -12 gpurun_out/${TAG}_sanitizer_memcheck.txt; tail -30 gpurun_out/${TAG}_sanitizer_racecheck.txt
