for c in 16 17 19 20; do ZKW_MSM_WINDOW_BITS=$c timeout 300 python tools/msm_ab.py 2>&1 | tail -1 | sed "s/^/c=$c /"; done > gpurun_out/s7_window.txt
cat gpurun_out/s7_window.txt
