( time timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6 > gpurun_out/s3_pytest.txt
timeout 600 python tools/ntt_ab.py > gpurun_out/s3_ntt_ab.txt 2>&1
ZKW_NTT_NO_ZERO_SKIP=1 timeout 600 python tools/ntt_ab.py > gpurun_out/s3_ntt_ab_noskip.txt 2>&1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/s3_bench_n1.json 2> gpurun_out/s3_bench.err
cat gpurun_out/s3_pytest.txt gpurun_out/s3_ntt_ab.txt gpurun_out/s3_ntt_ab_noskip.txt; cut -c1-250 gpurun_out/s3_bench_n1.json
