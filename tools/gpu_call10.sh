#!/bin/bash
# 2-GPU call: NCCL split-MSM test, split-MSM sweep, bench at N=2
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/c10_pytest.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sweep_multi.py --ks 19 20 22 24 --out gpurun_out/r1c_sweep_split_msm_n2.json > gpurun_out/c10_sweep.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r1c_bench_n2.json 2> gpurun_out/c10_bench.err
grep -E "passed|failed|error" gpurun_out/c10_pytest.txt | tail -2; grep '"k"' gpurun_out/c10_sweep.log; cut -c1-300 gpurun_out/r1c_bench_n2.json; tail -3 gpurun_out/c10_bench.err
