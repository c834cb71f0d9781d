#!/bin/bash
# r02f: flow issuer unrolled (descriptor templates), conv_post taps-in-N with the epilogue split over both column groups
OUT=gpurun_out; mkdir -p $OUT
( timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "low_voice or persistent or benchmarked or tensor_core_mrf" ) > $OUT/r02f_pytest.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|rror|worst RMS|RMS vs" $OUT/r02f_pytest.log | tail -14
bash tools/ab_env.sh "" "M3B200_DEC_POST_K7=1" "" 2>&1 | tee $OUT/r02f_ab.txt
