#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/c7_pytest.txt 2>&1
python tools/msm_ab.py > gpurun_out/c7_ab.txt 2>&1
python tools/msm_ab.py >> gpurun_out/c7_ab.txt 2>&1
python bench.py > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
grep -E "passed|failed|error" gpurun_out/c7_pytest.txt | tail -3; cat gpurun_out/c7_ab.txt; python - <<'PY'
import json
d=json.load(open('gpurun_out/c7_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['batch']['value'], d['roofline']['frac'], d['roofline']['isolated'], d['cpu_baseline']['value'])
PY
tail -5 gpurun_out/c7_bench.err
