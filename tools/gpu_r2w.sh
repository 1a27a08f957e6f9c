#!/bin/bash
# A/B of the L2 access-policy window over the chain kernel's scratch: live time and DRAM bytes (ncu launch list)
for w in 0 1; do
  TH_CHAIN_L2WIN=$w python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-culled > gpurun_out/r2w_bench_$w.json 2> gpurun_out/r2w_bench_$w.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2w_bench_$w.json').read().strip().splitlines()[-1])
print('L2WIN=$w', d['value'], d['roofline']['frac'], d['ms_per_step_by_category'], d['clocks'])
"
  grep "L2 window" gpurun_out/r2w_bench_$w.err | head -2
  TH_CHAIN_L2WIN=$w ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum --clock-control none -k regex:"k_chain|k_features" -s 6 -c 6 --csv --log-file gpurun_out/r2w_launches_$w.csv python bench.py --profile-run > /dev/null 2>&1
  python tools/ncu_summary.py gpurun_out/r2w_launches_$w.csv
done
