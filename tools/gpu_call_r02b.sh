#!/bin/bash
# r02b: full GPU suite (new parity tests on the benchmarked batches) + 1-GPU bench lines for cfg3/cfg2/cfg4/cfg5
OUT=gpurun_out
mkdir -p $OUT
( time timeout 600 python -m pytest tests -m gpu -x -q -s ) > $OUT/r02b_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/r02b_pytest.log
tail -5 $OUT/r02b_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 > $OUT/r02b_bench_cfg3.json 2> $OUT/r02b_bench_cfg3.err; echo "cfg3 exit $?"
for w in cfg2 cfg4 cfg5; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r02b_bench_$w.json 2> $OUT/r02b_bench_$w.err; echo "$w exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02b_bench_*.json')):
    try:
        d=json.loads(open(f).readline())
        print(f, 'ms/step %.2f value %.1f M/s e2e %.1f M/s'%(d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6), {k:round(v,2) for k,v in d['stage_ms_per_step'].items() if v>0.05}, 'launches/step', d['gpu_launches']/d['steps'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
