#!/bin/bash
# GPU call 2: warp-per-row bucket reduction (tests + A/B), inversion bench, NTT fake-twiddle A/B, proof timeline.
mkdir -p gpurun_out
tools/bin/inv_bench > gpurun_out/c2_inv.txt 2>&1
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/c2_pytest.txt 2>&1
ZKW_B200_LIB=$PWD/webauthn-halo2_b200/ab/libzkw_old.so python tools/msm_ab.py > gpurun_out/c2_ab.txt 2>&1
python tools/msm_ab.py >> gpurun_out/c2_ab.txt 2>&1
ZKW_B200_LIB=$PWD/webauthn-halo2_b200/ab/libzkw_old.so python tools/msm_ab.py >> gpurun_out/c2_ab.txt 2>&1
python tools/msm_ab.py >> gpurun_out/c2_ab.txt 2>&1
python tools/ntt_ab.py >> gpurun_out/c2_ab.txt 2>&1
ZKW_B200_LIB=$PWD/webauthn-halo2_b200/ab/libzkw_faketw.so python tools/ntt_ab.py >> gpurun_out/c2_ab.txt 2>&1
python tools/timeline.py gpurun_out/c2_timeline.csv > gpurun_out/c2_timeline.txt 2>&1
tail -3 gpurun_out/c2_pytest.txt; cat gpurun_out/c2_ab.txt gpurun_out/c2_inv.txt gpurun_out/c2_timeline.txt
