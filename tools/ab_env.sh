#!/bin/bash
# A/B of environment switches on the 1-GPU bench: one line per configuration.
#   bash tools/ab_env.sh "NAME=VALUE ..." "NAME2=VALUE2" ...      ("" = defaults)
for cfg in "$@"; do
  echo -n "== [${cfg:-defaults}]  "
  env $cfg timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('ms/step %.2f  value %.1f M/s'%(d['ms_per_step'], d['value']/1e6), {k:round(v,2) for k,v in d['stage_ms_per_step'].items() if v>0.05})"
done
