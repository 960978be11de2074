timeout 900 python -m pytest tests/test_gpu_msm.py -x -q 2>&1 | tail -2 > gpurun_out/s19.txt
timeout 600 python tools/sort_ab.py 2>&1 | tail -3 >> gpurun_out/s19.txt
python tools/msm_ab.py >> gpurun_out/s19.txt 2>&1
python tools/msm_ab.py >> gpurun_out/s19.txt 2>&1
cat gpurun_out/s19.txt
