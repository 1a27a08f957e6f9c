#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
for cfg in "1 inplace" "0 inplace" "0 tmem"; do
  set -- $cfg
  TH_CHAIN_DEFER=$1 TH_CHAIN_MIX=$2 timeout 900 ncu --metrics $M --clock-control none -k regex:"k_chain|k_features|k_gemm_tc2|k_integrate" -s 8 -c 6 --csv --log-file gpurun_out/r2j_launches_defer$1_$2.csv python bench.py --profile-run > gpurun_out/r2j_ncu.log 2>&1
  echo "== defer=$1 mix=$2"; python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2j_launches_defer$1_$2.csv')) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
agg={}
for r in rows[1:]:
    k=(r[ix['ID']], r[ix['Kernel Name']][:40]); agg.setdefault(k,{})[r[ix['Metric Name']]]=r[ix['Metric Value']]
for k,v in agg.items(): print(k, {m.split('.')[0][-22:]:x for m,x in v.items()})
PY
done
