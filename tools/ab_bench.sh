#!/bin/bash
for cfg in "16 16" "8 16" "16 8" "8 8"; do
  set -- $cfg
  echo -n "== DEC_WARPS=$1 MRF_WARPS=$2  "
  M3B200_DEC_WARPS=$1 M3B200_MRF_WARPS=$2 timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('ms/step %.2f  value %.1f'%(d['ms_per_step'], d['value']/1e6), {k:round(v,2) for k,v in d['stage_ms_per_step'].items() if k in ('mrf','upsample','flow','text_encoder')})"
done
