#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ntt.py tests/test_kat_fixture.py tests/test_gpu_quotient.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/c13.txt
python tools/ntt_ab.py >> gpurun_out/c13.txt 2>&1
python tools/ntt_ab.py >> gpurun_out/c13.txt 2>&1
python tools/msm_ab.py >> gpurun_out/c13.txt 2>&1
( time python -m pytest tests -m gpu -x -q ) 2>&1 | grep -E "passed|failed|error|real" >> gpurun_out/c13.txt
cat gpurun_out/c13.txt
