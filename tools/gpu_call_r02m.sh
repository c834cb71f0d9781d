#!/bin/bash
# r02m: C=64 MRF kernel with two taps per ring slot / issuer stage; full GPU suite on the tree to be committed
OUT=gpurun_out; mkdir -p $OUT
( time timeout 600 python -m pytest tests -m gpu -x -q -s ) > $OUT/r02m_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/r02m_pytest.log
grep -E "passed|failed|rror" $OUT/r02m_pytest.log | tail -4
bash tools/ab_env.sh "" "M3B200_DEC_WARPS2=16" "" 2>&1 | tee $OUT/r02m_ab.txt
