#!/bin/bash
# A/B of MRF tile configurations (stage times only)
for cfg in "4 2" "2 2" "3 2"; do
  set -- $cfg
  echo "== NT32=$1 NT64=$2"
  M3B200_MRF_NT32=$1 M3B200_MRF_NT64=$2 timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('value %.1f M/s  ms/step %.2f'%(d['value']/1e6,d['ms_per_step']), {k:round(v,2) for k,v in d['stage_ms_per_step'].items()})"
done
