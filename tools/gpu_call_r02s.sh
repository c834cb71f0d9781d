#!/bin/bash
# r02s: dec_planes_kernel evidence: per-role cycle counters (profile build), launch list, ncu --set full of the kernel
OUT=gpurun_out; mkdir -p $OUT
M3B200_LIBRARY=$PWD/mimic3_b200/libm3b200_prof.so M3B200_DEC_PROFILE=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "profile\]" | tail -2 | tee $OUT/r02s_role_cycles.txt
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02s_launches.csv python bench.py --steps 1 --warmup 1 --profile-only > $OUT/r02s_launches.log 2>&1; echo "launch list exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"dec_planes_kernel" -c 1 -f -o $OUT/r02s_full \
    python bench.py --batch 64 --steps 1 --warmup 0 --profile-only > $OUT/r02s_full.log 2>&1; echo "ncu full exit $?"
ncu -i $OUT/r02s_full.ncu-rep --page raw --csv > $OUT/r02s_full_raw.csv 2>/dev/null
ncu -i $OUT/r02s_full.ncu-rep --page source --csv > $OUT/r02s_full_source.csv 2>/dev/null
ls -la $OUT | tail -8
