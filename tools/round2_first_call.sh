#!/bin/bash
# First GPU call of round 2: do the kernels staged at the end of round 1 (never run on hardware) work, and what do
# they buy?   gpurun --timeout 600 -- 'bash tools/round2_first_call.sh'
# Every step is bounded by `timeout`; a protocol bug in a staged kernel trips the ~2 s watchdog in tc::mbar_wait
# (a CUDA error, not a hung GPU).
OUT=gpurun_out
mkdir -p $OUT
( timeout 240 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -3 | tee $OUT/r02a_pytest_default.log
( M3B200_TEST_EXPERIMENTAL=1 timeout 240 python -m pytest tests/test_gpu_experimental.py -q -s ) 2>&1 | tail -15 | tee $OUT/r02a_pytest_experimental.log
bash tools/ab_env.sh "" M3B200_UPS_V2=1 M3B200_ROWGEMM_V2=1 "M3B200_UPS_V2=1 M3B200_ROWGEMM_V2=1" 2>&1 | tee $OUT/r02a_ab_staged.txt
