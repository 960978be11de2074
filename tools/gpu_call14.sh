#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/c14.txt
for i in 1 2 3; do
python tools/ntt_ab.py >> gpurun_out/c14.txt 2>&1
ZKW_B200_LIB=$PWD/webauthn-halo2_b200/ab/libzkw_nostage.so python tools/ntt_ab.py >> gpurun_out/c14.txt 2>&1
done
for i in 1 2; do
python tools/msm_ab.py >> gpurun_out/c14.txt 2>&1
ZKW_B200_LIB=$PWD/webauthn-halo2_b200/ab/libzkw_nostage.so python tools/msm_ab.py >> gpurun_out/c14.txt 2>&1
done
cat gpurun_out/c14.txt
