#!/bin/bash
# GPU call 3: stream priorities + merged challenge-independent commitments + single-instance tree kernels.
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/c3_pytest.txt 2>&1
ZKW_B200_LIB=$PWD/webauthn-halo2_b200/ab/libzkw_old.so python tools/msm_ab.py > gpurun_out/c3_ab.txt 2>&1
python tools/msm_ab.py >> gpurun_out/c3_ab.txt 2>&1
python tools/msm_ab.py >> gpurun_out/c3_ab.txt 2>&1
python tools/timeline.py gpurun_out/c3_timeline.csv > gpurun_out/c3_timeline.txt 2>&1
tail -3 gpurun_out/c3_pytest.txt; cat gpurun_out/c3_ab.txt gpurun_out/c3_timeline.txt
