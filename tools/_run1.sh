( time timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5 > gpurun_out/s16.txt
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/s16_bench_n1.json 2> gpurun_out/s16_bench.err
cat gpurun_out/s16.txt; cut -c1-200 gpurun_out/s16_bench_n1.json
