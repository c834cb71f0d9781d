#!/bin/bash
# r02k: TMEM loads batched per epilogue phase (one tcgen05.wait::ld per phase) in dec_fused_kernel and mrf_ws_kernel;
# commit-period micro-benchmark
OUT=gpurun_out; mkdir -p $OUT
( timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "low_voice or persistent or tensor_core_mrf or benchmarked_config2" ) > $OUT/r02k_pytest.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|rror|worst RMS|RMS vs" $OUT/r02k_pytest.log | tail -8
bash tools/ab_env.sh "" "M3B200_DEC_WARPS2=16" "M3B200_MRF64_WARPS=16" "" 2>&1 | tee $OUT/r02k_ab.txt
timeout 200 python tools/ubench.py 2>&1 | tail -16 | tee $OUT/r02k_ubench_commit.txt
