#!/bin/bash
# GPU call 1: validate the row/column bucket reduction, A/B it against the previous library, pipe rates, bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
tools/bin/pipe_bench > gpurun_out/c1_pipe.txt 2>&1
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/c1_pytest.txt 2>&1
ZKW_B200_LIB=$PWD/webauthn-halo2_b200/ab/libzkw_old.so python tools/msm_ab.py > gpurun_out/c1_ab.txt 2>&1
python tools/msm_ab.py >> gpurun_out/c1_ab.txt 2>&1
ZKW_B200_LIB=$PWD/webauthn-halo2_b200/ab/libzkw_old.so python tools/msm_ab.py >> gpurun_out/c1_ab.txt 2>&1
python tools/msm_ab.py >> gpurun_out/c1_ab.txt 2>&1
python bench.py > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
tail -3 gpurun_out/c1_pytest.txt; cat gpurun_out/c1_ab.txt gpurun_out/c1_pipe.txt; cat gpurun_out/c1_bench.json | cut -c1-600
