#!/bin/bash
# Final single-GPU campaign for the round: tests, smoke, bench (both arms), ncu launch list + NTT capture, timeline.
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/f1_pytest.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f1_smoke.txt 2>&1
python bench.py > gpurun_out/r1c_bench_n1.json 2> gpurun_out/f1_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1c_bench_reference.json 2> gpurun_out/f1_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-flavours --batch 0 > gpurun_out/f1_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ntt_pass_kernel -s 3 -c 6 -f -o gpurun_out/r1c_prof_ntt python tools/prof_target.py ntt > gpurun_out/f1_prof2.log 2>&1
python tools/timeline.py gpurun_out/r1c_timeline_proof.csv > gpurun_out/f1_timeline.txt 2>&1
grep -E "passed|failed|error" gpurun_out/f1_pytest.txt | tail -2; cat gpurun_out/f1_smoke.txt | tail -2; cut -c1-200 gpurun_out/r1c_bench_n1.json; cut -c1-400 gpurun_out/r1c_bench_reference.json; cat gpurun_out/f1_timeline.txt
