ZKW_E2E_TRACE=1 python tools/timeline_e2e.py gpurun_out/s14_timeline_e2e.csv 2>&1 | grep "e2e\|launches" | tail -5 > gpurun_out/s14.txt
python tools/e2e_parts.py >> gpurun_out/s14.txt 2>&1
ZKW_SYNTH_NO_WRITEBACK=1 python tools/e2e_parts.py >> gpurun_out/s14.txt 2>&1
python tools/e2e_parts.py 17 >> gpurun_out/s14.txt 2>&1
ZKW_SYNTH_NO_WRITEBACK=1 python tools/e2e_parts.py 17 >> gpurun_out/s14.txt 2>&1
cat gpurun_out/s14.txt
