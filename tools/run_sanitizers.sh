#!/bin/bash
# compute-sanitizer passes over the GPU tests that cover every kernel (one GPU); TAG names the round.  Output: gpurun_out/${TAG}_sanitizers.txt
TAG=${1:-r2}
OUT=gpurun_out/${TAG}_sanitizers.txt
mkdir -p gpurun_out
{
echo "compute-sanitizer runs on a B200 (commands run from the repo root under gpurun)"
echo
echo "memcheck: tests/test_gpu_prover.py tests/test_gpu_ntt.py tests/test_gpu_msm.py -k 'bit_for_bit or lookup_rank or test_ntt_matches_oracle or skewed or dedicated_squaring or edge or both_entry_sorts'"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_prover.py tests/test_gpu_ntt.py tests/test_gpu_msm.py -x -q -k "bit_for_bit or lookup_rank or test_ntt_matches_oracle or skewed or dedicated_squaring or edge or both_entry_sorts" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | tail -4
echo
echo "racecheck: tests/test_gpu_prover.py tests/test_gpu_ntt.py tests/test_gpu_msm.py -k '(bit_for_bit and evm-gwc) or test_inverse_ntt_with_scale or lookup_rank or (both_entry_sorts and 4097) or (skewed and all_equal)'"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_prover.py tests/test_gpu_ntt.py tests/test_gpu_msm.py -x -q -k "(bit_for_bit and evm-gwc) or test_inverse_ntt_with_scale or lookup_rank or (both_entry_sorts and 4097) or (skewed and all_equal)" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|Race reported|Error" | sort | uniq -c | tail -8
echo
echo "synccheck: tests/test_gpu_prover.py -k 'bit_for_bit and evm-gwc-7'; tests/test_gpu_msm.py -k 'both_entry_sorts and 4097'"
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_prover.py tests/test_gpu_msm.py -x -q -k "(bit_for_bit and evm-gwc-7) or (both_entry_sorts and 4097)" 2>&1 | grep -E "passed|failed|ERROR SUMMARY" | tail -3
} > $OUT 2>&1
cat $OUT
