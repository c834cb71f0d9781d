#!/bin/bash
# r02n: second-generation flow kernel (kernels_tc_flow2.cu): parity vs first kernel / unfused / oracle, A/B
OUT=gpurun_out; mkdir -p $OUT
( timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "flow_second or low_voice or benchmarked_config2 or long_form" ) > $OUT/r02n_pytest.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|rror|flow2 vs|worst RMS" $OUT/r02n_pytest.log | tail -14
bash tools/ab_env.sh "" "M3B200_FLOW_V1=1" "" 2>&1 | tee $OUT/r02n_ab.txt
