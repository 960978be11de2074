#!/bin/bash
# Single-GPU campaign for a round: tests, smoke, bench (both arms, the driver's flags), ncu launch list, timeline, sweep.
TAG=${1:-r2}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-flavours --no-sweep --batch 0 > gpurun_out/${TAG}_launches_bench.log 2>&1
python tools/timeline.py gpurun_out/${TAG}_timeline_proof.csv > gpurun_out/${TAG}_timeline.txt 2>&1
python tools/sweep.py --out gpurun_out/${TAG}_sweep_msm_ntt.json > /dev/null 2> gpurun_out/${TAG}_sweep.err
grep -E "passed|failed|error" gpurun_out/${TAG}_pytest.txt | tail -2; tail -2 gpurun_out/${TAG}_smoke.txt; cut -c1-200 gpurun_out/${TAG}_bench_n1.json; cut -c1-300 gpurun_out/${TAG}_bench_reference.json; cat gpurun_out/${TAG}_timeline.txt; tail -2 gpurun_out/${TAG}_sweep.err
