bash tools/run_final_1gpu.sh r2j > gpurun_out/r2j_campaign.txt 2>&1
bash tools/run_ncu_captures.sh r2j > gpurun_out/r2j_ncu.txt 2>&1
for f in msm_acc ntt msm_small quot; do python tools/ncu_summary.py gpurun_out/r2j_prof_$f.ncu-rep --raw-out gpurun_out/r2j_prof_${f}_raw.csv > gpurun_out/r2j_prof_${f}_summary.txt 2>&1; done
rm -f gpurun_out/r2j_prof_*.ncu-rep
python tools/timeline_e2e.py gpurun_out/r2j_timeline_e2e.csv > gpurun_out/r2j_timeline_e2e.txt 2>&1
python tools/sort_ab.py > gpurun_out/r2j_sort_ab.txt 2>&1
cat gpurun_out/r2j_campaign.txt | cut -c1-300; cat gpurun_out/r2j_prof_msm_acc_summary.txt
