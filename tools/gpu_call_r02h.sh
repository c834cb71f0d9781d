#!/bin/bash
# r02h: per-role cycle counters of the flow / C=64 MRF / last-stage kernels on the benchmark batch
OUT=gpurun_out; mkdir -p $OUT
M3B200_FLOW_PROFILE=1 M3B200_DEC_PROFILE=1 M3B200_MRF_PROFILE=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/r02h_prof.json 2> $OUT/r02h_prof.err
echo "exit $?"; grep -E "profile\]" $OUT/r02h_prof.err | tail -12
