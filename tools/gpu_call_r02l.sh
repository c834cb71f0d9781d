#!/bin/bash
# r02l: weight-ring consumers without the per-stage tcgen05.fence (M3B200_RING_FENCE=1 restores it); TMEM-read micro-benchmark
OUT=gpurun_out; mkdir -p $OUT
( timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "low_voice or persistent or tensor_core_mrf or benchmarked_config2" ) > $OUT/r02l_pytest.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|rror" $OUT/r02l_pytest.log | tail -3
bash tools/ab_env.sh "" "M3B200_RING_FENCE=1" "" 2>&1 | tee $OUT/r02l_ab.txt
timeout 200 python tools/ubench.py 2>&1 | tail -6 | tee $OUT/r02l_ubench_tmem.txt
