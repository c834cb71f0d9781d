#!/bin/bash
# r02d: evidence refresh after the persistent C=128 MRF kernel + sanitizer passes over the fused kernels
OUT=gpurun_out; mkdir -p $OUT
export SKIP=""
export FULL_K="mrf_ws128_kernel|flow_tc_kernel|mrf_ws_kernel|dec_fused_kernel"
export FULL_C=4
bash tools/gpu_round.sh r02d
for tool in memcheck racecheck synccheck; do
  ( time timeout 300 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_run.py ) > $OUT/r02d_sanitizer_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run ok|Error:|hazard" $OUT/r02d_sanitizer_$tool.log | head -8
done
( timeout 200 python __graft_entry__.py --smoke ) > $OUT/r02d_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/r02d_smoke.log
