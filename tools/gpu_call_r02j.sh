#!/bin/bash
# r02j: 16 epilogue warps in the C=64 MRF kernel (M3B200_MRF64_WARPS=16) and the last-stage kernel (M3B200_DEC_WARPS2=16)
OUT=gpurun_out; mkdir -p $OUT
( M3B200_MRF64_WARPS=16 M3B200_DEC_WARPS2=16 timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "low_voice or persistent or tensor_core_mrf or benchmarked_config2" ) > $OUT/r02j_pytest.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|rror|worst RMS|RMS vs" $OUT/r02j_pytest.log | tail -8
bash tools/ab_env.sh "" "M3B200_MRF64_WARPS=16" "M3B200_DEC_WARPS2=16" "M3B200_MRF64_WARPS=16 M3B200_DEC_WARPS2=16" "" 2>&1 | tee $OUT/r02j_ab.txt
