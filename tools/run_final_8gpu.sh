#!/bin/bash
# 8-GPU box: bench at N=8 and N=4, split-MSM sweep at N=4 and N=8
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r1c_bench_n8.json 2> gpurun_out/f8_bench8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r1c_bench_n4.json 2> gpurun_out/f8_bench4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 tools/sweep_multi.py --ks 20 22 24 --out gpurun_out/r1c_sweep_split_msm_n8.json > gpurun_out/f8_sweep8.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 tools/sweep_multi.py --ks 20 22 24 --out gpurun_out/r1c_sweep_split_msm_n4.json > gpurun_out/f8_sweep4.log 2>&1
cut -c1-220 gpurun_out/r1c_bench_n8.json; cut -c1-220 gpurun_out/r1c_bench_n4.json; grep '"k"' gpurun_out/f8_sweep8.log gpurun_out/f8_sweep4.log | cut -c1-200
