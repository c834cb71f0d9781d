#!/bin/bash
# r02e: SFU gate math in the flow kernel (A/B = bench stage times vs r02d), full GPU suite incl. the weight-cache test,
# ncu --set full + source of the three persistent decoder kernels
OUT=gpurun_out; mkdir -p $OUT
( time timeout 600 python -m pytest tests -m gpu -x -q -s ) > $OUT/r02e_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/r02e_pytest.log
grep -E "passed|failed|error|plain load|worst RMS" $OUT/r02e_pytest.log | tail -12
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/r02e_bench.json 2> $OUT/r02e_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02e_bench.json').readline())
print('ms/step %.2f value %.1f M/s e2e %.1f M/s'%(d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6), {k:round(v,2) for k,v in d['stage_ms_per_step'].items() if v>0.05})
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"dec_fused_kernel|mrf_ws_kernel|mrf_ws128_kernel" -c 3 -f -o $OUT/r02e_full \
    python bench.py --batch 64 --steps 1 --warmup 0 --profile-only > $OUT/r02e_full.log 2>&1
echo "ncu full exit $?"
ncu -i $OUT/r02e_full.ncu-rep --page raw --csv > $OUT/r02e_full_raw.csv 2>/dev/null
ncu -i $OUT/r02e_full.ncu-rep --page source --print-source cuda,sass --csv > $OUT/r02e_full_source.csv 2>/dev/null
ls -la $OUT | tail -8
