#!/bin/bash
# ncu evidence for round 1 (third pass): launch list of the bench command + full captures of the dominant kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --batch 0 > gpurun_out/r1c_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_kernel -s 1 -c 1 -o gpurun_out/r1c_prof_msm_acc python tools/prof_target.py > gpurun_out/r1c_prof1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ntt_pass_kernel -s 12 -c 3 -o gpurun_out/r1c_prof_ntt python tools/prof_target.py > gpurun_out/r1c_prof2.log 2>&1
ncu --set full --clock-control none -k regex:'msm_(recode|scan|scatter|combine_light|combine_heavy|rowcol|weighted)_kernel' -s 14 -c 7 -o gpurun_out/r1c_prof_msm_small python tools/prof_target.py > gpurun_out/r1c_prof3.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/r1c_prof1.log gpurun_out/r1c_prof2.log gpurun_out/r1c_prof3.log
