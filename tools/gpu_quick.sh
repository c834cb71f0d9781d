#!/bin/bash
# quick kernel iteration:  bash tools/gpu_quick.sh "<pytest -k expression>" "<env A>" "<env B>" ...
OUT=gpurun_out; mkdir -p $OUT
K="$1"; shift
( timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "$K" ) 2>&1 | tail -25
bash tools/ab_env.sh "$@"
