#!/bin/bash
# One gpurun call's worth of evidence: GPU parity tests, the 1-GPU bench line, the ncu launch list of
# the same command and one `ncu --set full` capture of the dominant kernels.  Outputs under gpurun_out/.
#   gpurun --timeout 900 -- 'bash tools/gpu_round.sh <tag>'
# Sections can be skipped with SKIP="tests bench launches full".
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
skip() { case " $SKIP " in *" $1 "*) return 0;; esac; return 1; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > $OUT/${TAG}_gpu.txt 2>&1

if ! skip tests; then
  ( time timeout 420 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -5 $OUT/${TAG}_pytest.log
fi
if ! skip bench; then
  timeout 300 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
  echo "bench exit $?"; cat $OUT/${TAG}_bench.json
fi
if ! skip launches; then
  timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --profile-only > $OUT/${TAG}_launches.log 2>&1
  echo "launch list exit $?"
fi
if ! skip full; then
  timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:"${FULL_K:-dec_fused_kernel|mrf_ws_kernel|mrf_tc_kernel}" -c ${FULL_C:-3} -f -o $OUT/${TAG}_full \
    python bench.py --batch 64 --steps 1 --warmup 0 --profile-only > $OUT/${TAG}_full.log 2>&1
  echo "ncu full exit $?"
  ncu -i $OUT/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2>/dev/null
fi
if [ -n "$EXTRA" ]; then bash -c "$EXTRA" > $OUT/${TAG}_extra.log 2>&1; echo "extra exit $?"; tail -20 $OUT/${TAG}_extra.log; fi
