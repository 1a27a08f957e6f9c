#!/bin/bash
# Session-3 evidence pass on the final build: ncu launch list (time, DRAM bytes, L2 sectors) of three consecutive
# chunks of the dense frame, one `--set full` capture each of k_features and k_chain, compute-sanitizer on smoke().
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum
ncu --metrics $M --clock-control none -k regex:"k_chain|k_features" -s 6 -c 6 --csv --log-file gpurun_out/s3_launches_final.csv python bench.py --profile-run > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/s3_launches_final.csv
ncu --set full --import-source on --clock-control none -k regex:k_features -s 3 -c 1 -o gpurun_out/s3_k_features -f python bench.py --profile-run > gpurun_out/s3_ncu_k_features.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_chain -s 3 -c 1 -o gpurun_out/s3_k_chain -f python bench.py --profile-run > gpurun_out/s3_ncu_k_chain.log 2>&1
ls -la gpurun_out/s3_k_*.ncu-rep
for tool in memcheck racecheck; do
  echo "===== compute-sanitizer --tool $tool (final build, smoke(): 24x24x16 dense + culled)" > gpurun_out/s3_sanitizer_$tool.txt
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^$" | tail -25 >> gpurun_out/s3_sanitizer_$tool.txt
done
tail -6 gpurun_out/s3_sanitizer_memcheck.txt; tail -12 gpurun_out/s3_sanitizer_racecheck.txt
