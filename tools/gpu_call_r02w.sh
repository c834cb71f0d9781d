#!/bin/bash
# r02w: evidence after dec_planes_kernel: full GPU suite, bench line (cfg3 with CPU leg; cfg2/4/5), launch list, ncu --set full
# of the last-stage kernel, per-role cycle counters (profile build)
OUT=gpurun_out; mkdir -p $OUT; T=${TAG:-r03b}
( time timeout 700 python -m pytest tests -m gpu -x -q -s ) > $OUT/${T}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${T}_pytest.log; grep -E "passed|failed|rror" $OUT/${T}_pytest.log | tail -3
timeout 400 python bench.py --steps 20 --warmup 3 > $OUT/${T}_bench_cfg3.json 2> $OUT/${T}_bench_cfg3.err; echo "cfg3 exit $?"
for w in cfg2 cfg4 cfg5; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_$w.json 2> $OUT/${T}_bench_$w.err; echo "$w exit $?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_bench_*.json')):
    try:
        d=json.loads(open(f).readline())
        print(f, 'ms/step %.2f value %.1f M/s e2e %.1f M/s'%(d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6), {k:round(v,2) for k,v in d['stage_ms_per_step'].items() if v>0.05}, 'launches/step', d['gpu_launches']/d['steps'], 'roofline', round(d['roofline']['frac'],3))
    except Exception as e:
        print(f, 'FAILED', e)
PY
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${T}_launches.csv python bench.py --steps 1 --warmup 1 --profile-only > $OUT/${T}_launches.log 2>&1; echo "launch list exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"dec_planes_kernel|mrf_ws_kernel" -c 2 -f -o $OUT/${T}_full \
    python bench.py --batch 64 --steps 1 --warmup 0 --profile-only > $OUT/${T}_full.log 2>&1; echo "ncu full exit $?"
ncu -i $OUT/${T}_full.ncu-rep --page raw --csv > $OUT/${T}_full_raw.csv 2>/dev/null
M3B200_LIBRARY=$PWD/mimic3_b200/libm3b200_prof.so M3B200_FLOW_PROFILE=1 M3B200_DEC_PROFILE=1 M3B200_MRF_PROFILE=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "profile\]" | tail -4 | tee $OUT/${T}_role_cycles.txt
rm -f $OUT/${T}_full.ncu-rep
