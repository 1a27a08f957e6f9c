#!/bin/bash
# One GPU-box visit: parity tests, the bench line, and (optionally) ncu captures.
# usage: tools/gpu_check.sh <tag> [tests] [bench] [launches] [ncu:<kernel regex>:<skip>:<count>]
tag=$1; shift
mkdir -p gpurun_out
for what in "$@"; do
  case $what in
    tests) timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${tag}_tests.log;;
    bench) timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_ -s 4 -c 8 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --profile-run > gpurun_out/${tag}_launches.log 2>&1; echo "launches rc=$?";;
    ncu:*) IFS=: read -r _ rx skip cnt <<< "$what"
       timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o gpurun_out/${tag}_${rx} python bench.py --profile-run > gpurun_out/${tag}_ncu_${rx}.log 2>&1; echo "ncu $rx rc=$?";;
  esac
done
