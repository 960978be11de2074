#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/c6_pytest.txt 2>&1
for v in prev c4 c2; do ZKW_B200_LIB=$PWD/webauthn-halo2_b200/ab/libzkw_$v.so python tools/msm_ab.py >> gpurun_out/c6_ab.txt 2>&1; done
python tools/msm_ab.py >> gpurun_out/c6_ab.txt 2>&1
python tools/msm_ab.py >> gpurun_out/c6_ab.txt 2>&1
python tools/timeline.py gpurun_out/c6_timeline.csv > gpurun_out/c6_timeline.txt 2>&1
grep -E "passed|failed|error" gpurun_out/c6_pytest.txt | tail -3; cat gpurun_out/c6_ab.txt gpurun_out/c6_timeline.txt
