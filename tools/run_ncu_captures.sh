#!/bin/bash
# ncu captures of the hot kernels (one GPU): python tools are short targets, see tools/prof_target.py.  TAG names the round.
TAG=${1:-r2}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_msm_acc python tools/prof_target.py msm > gpurun_out/${TAG}_prof1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ntt_pass_kernel -s 3 -c 6 -f -o gpurun_out/${TAG}_prof_ntt python tools/prof_target.py ntt > gpurun_out/${TAG}_prof2.log 2>&1
ncu --set full --clock-control none -k regex:'msm_(recode|scan|scatter|bin_count|bin_scan|bin_scatter|fine_count|fine_scatter|combine_light|combine_heavy|rowcol|weighted|reduce)' -s 10 -c 10 -f -o gpurun_out/${TAG}_prof_msm_small python tools/prof_target.py msm > gpurun_out/${TAG}_prof3.log 2>&1
ncu --set full --clock-control none -k regex:quotient_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_quot python tools/prof_target.py quot > gpurun_out/${TAG}_prof4.log 2>&1
for f in 1 2 3 4; do tail -n 2 gpurun_out/${TAG}_prof$f.log; done
