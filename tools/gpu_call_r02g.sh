#!/bin/bash
# r02g: micro-benchmarks of the MMA stream as the conv kernels issue it (shifted / odd-pitch / varying operands,
# concurrent LSU and bulk-copy traffic) + the full GPU suite on the current tree
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/ubench.py > $OUT/r02g_ubench.txt 2>&1; echo "ubench exit $?"; tail -45 $OUT/r02g_ubench.txt
( time timeout 600 python -m pytest tests -m gpu -x -q -s ) > $OUT/r02g_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/r02g_pytest.log
grep -E "passed|failed|rror" $OUT/r02g_pytest.log | tail -5
bash tools/ab_env.sh "" 2>&1 | tee $OUT/r02g_ab.txt
