#!/bin/bash
# 8-GPU box: bench.py at N = 8, 4, 2 (one process per GPU, torchrun), then the split-MSM sweep at N = 8 and 4
mkdir -p gpurun_out
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r1c_bench_n$n.json 2> gpurun_out/f8_bench$n.err
done
if [ "$1" = "sweep" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/sweep_multi.py --ks 20 22 24 --out gpurun_out/r1c_sweep_split_msm_n8.json > gpurun_out/f8_sweep8.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 tools/sweep_multi.py --ks 20 22 24 --out gpurun_out/r1c_sweep_split_msm_n4.json > gpurun_out/f8_sweep4.log 2>&1
fi
for n in 8 4 2; do cut -c1-160 gpurun_out/r1c_bench_n$n.json; done
