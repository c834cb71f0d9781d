#!/bin/bash
# r02i: 16 epilogue warps in the C=64 MRF kernel (M3B200_MRF64_WARPS=16): parity, A/B, per-role cycle counters
OUT=gpurun_out; mkdir -p $OUT
( M3B200_MRF64_WARPS=16 timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "low_voice or persistent or tensor_core_mrf or benchmarked_config2" ) > $OUT/r02i_pytest.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|rror|worst RMS|RMS vs" $OUT/r02i_pytest.log | tail -8
bash tools/ab_env.sh "" "M3B200_MRF64_WARPS=16" "" "M3B200_MRF64_WARPS=16" 2>&1 | tee $OUT/r02i_ab.txt
M3B200_MRF64_WARPS=16 M3B200_MRF_PROFILE=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "mrf_ws profile" | tail -1 | tee $OUT/r02i_prof16.txt
