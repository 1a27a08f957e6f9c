#!/bin/bash
mkdir -p gpurun_out
for cfg in "1 inplace" "0 inplace" "0 tmem" "1 inplace" "0 inplace"; do
  set -- $cfg
  TH_CHAIN_DEFER=$1 TH_CHAIN_MIX=$2 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-culled --no-extras > gpurun_out/r2i_ab.json 2> gpurun_out/r2i_ab.err
  python -c "
import json;d=json.loads(open('gpurun_out/r2i_ab.json').read().strip().splitlines()[-1]);print('defer=$1 mix=$2', round(d['value']), round(d['ms_per_step'],2), d['ms_per_step_by_category'], round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"
done
