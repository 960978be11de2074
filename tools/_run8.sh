mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r2j_n8_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3 > gpurun_out/r2j_multi_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2j_bench_n8.json 2> gpurun_out/r2j_bench_n8.err
cat gpurun_out/r2j_n8_gpus.txt gpurun_out/r2j_multi_pytest.txt; cut -c1-220 gpurun_out/r2j_bench_n8.json; tail -3 gpurun_out/r2j_bench_n8.err
