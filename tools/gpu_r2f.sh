#!/bin/bash
# ncu --set full + source of one k_chain launch, in-place mix variant (the faster one so far)
mkdir -p gpurun_out
TH_CHAIN_INPLACE_MIX=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 2 -c 1 -f -o gpurun_out/r2f_chain_inplace python bench.py --profile-run > gpurun_out/r2f_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2f_ncu.log
ls -la gpurun_out/r2f_chain_inplace.ncu-rep
