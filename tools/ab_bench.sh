#!/bin/bash
# elimination experiments on the MRF kernel (results are wrong on purpose; timing only)
for dbg in 0 1 2 4 6 8 16 31; do
  echo -n "== MRF_DEBUG=$dbg  "
  M3B200_MRF_DEBUG=$dbg timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('ms/step %.2f  mrf %.2f'%(d['ms_per_step'], d['stage_ms_per_step']['mrf']))"
done
