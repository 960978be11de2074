#!/bin/bash
mkdir -p gpurun_out
python tools/ntt_ab.py > gpurun_out/c11_ab.txt 2>&1
for v in t11 t12 t11s; do ZKW_B200_LIB=$PWD/webauthn-halo2_b200/ab/libzkw_$v.so python tools/ntt_ab.py >> gpurun_out/c11_ab.txt 2>&1; done
for v in t11 t12; do ZKW_B200_LIB=$PWD/webauthn-halo2_b200/ab/libzkw_$v.so python -m pytest tests/test_gpu_ntt.py tests/test_kat_fixture.py -m gpu -x -q 2>&1 | tail -2 >> gpurun_out/c11_ab.txt; done
for v in t11 t12; do ZKW_B200_LIB=$PWD/webauthn-halo2_b200/ab/libzkw_$v.so python tools/msm_ab.py >> gpurun_out/c11_ab.txt 2>&1; done
python tools/msm_ab.py >> gpurun_out/c11_ab.txt 2>&1
cat gpurun_out/c11_ab.txt
