#!/bin/bash
# multi-GPU bench lines on the box this runs on:  bash tools/gpu_call_scale.sh <tag> "<N list>" [extra bench args]
TAG=${1:-r02}; NS=${2:-"1 2"}; shift; shift
OUT=gpurun_out; mkdir -p $OUT
for N in $NS; do
  for SC in weak strong; do
    if [ "$N" = "1" ]; then
      [ "$SC" = "strong" ] && continue
      timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > $OUT/${TAG}_n${N}_${SC}.json 2> $OUT/${TAG}_n${N}_${SC}.err
    else
      timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
        bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --scaling $SC "$@" > $OUT/${TAG}_n${N}_${SC}.json 2> $OUT/${TAG}_n${N}_${SC}.err
    fi
    echo "N=$N $SC exit $?"; tail -3 $OUT/${TAG}_n${N}_${SC}.err | cut -c1-300
  done
done
python - "$TAG" <<'PY'
import json,glob,sys
tag=sys.argv[1]
base=None
for f in sorted(glob.glob(f'gpurun_out/{tag}_n*_*.json')):
    try:
        d=json.loads(open(f).readlines()[-1])
        if d['n_gpus']==1: base=d
        print(f, 'N=%d %s: ms/step %.2f value %.1f M/s | e2e %.1f M/s (%.2f ms/step)'%(d['n_gpus'], d['scaling'], d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d['e2e'].get('ms_per_step',0)),
              ('eff value %.2f e2e %.2f'%(d['value']/base['value']/d['n_gpus'], d['e2e']['value']/base['e2e']['value']/d['n_gpus'])) if base else '')
    except Exception as e:
        print(f,'FAILED',e)
PY
