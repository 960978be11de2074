#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_kernel -s 1 -c 1 -f -o gpurun_out/r1c_prof_msm_acc python tools/prof_target.py msm > gpurun_out/r1c_prof1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ntt_pass_kernel -s 3 -c 6 -f -o gpurun_out/r1c_prof_ntt python tools/prof_target.py ntt > gpurun_out/r1c_prof2.log 2>&1
ncu --set full --clock-control none -k regex:'msm_(recode|scan|scatter|combine_light|combine_heavy|rowcol|weighted)_kernel' -s 7 -c 7 -f -o gpurun_out/r1c_prof_msm_small python tools/prof_target.py msm > gpurun_out/r1c_prof3.log 2>&1
ncu --set full --clock-control none -k regex:quotient_kernel -s 1 -c 1 -f -o gpurun_out/r1c_prof_quot python tools/prof_target.py quot > gpurun_out/r1c_prof4.log 2>&1
for f in 1 2 3 4; do tail -n 2 gpurun_out/r1c_prof$f.log; done
