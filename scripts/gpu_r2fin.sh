#!/bin/bash
# same-box A/B of the scalar walk rate before / after the value-channel hooks; ncu of one steady round of the final kernels; launch list of a short bench
tag=${1:-r2fin}
out=gpurun_out/$tag
mkdir -p $out
run() { echo "== $1"; shift; env "$@" ZB_REPS=2 timeout 200 python scripts/profile_case.py 8 16384 64 2>&1 | tail -1; }
{
run "current" ZB_DUMMY=1
run "before the channel hooks" ZOMBIE_B200_LIB=zombie_b200/variants/libprechannels.so
run "current" ZB_DUMMY=1
run "before the channel hooks" ZOMBIE_B200_LIB=zombie_b200/variants/libprechannels.so
} 2>&1 | tee $out/ab.log
echo "== ncu: one steady-state round, full set"
ZB_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:wave -s 50 -c 4 -o $out/wave_steady python scripts/profile_case.py 8 16384 64 > $out/ncu_wave.log 2>&1; echo "rc=$?"; tail -2 $out/ncu_wave.log
echo "== ncu launch list of a short bench"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --batch 4096 --no-cpu > $out/bench_under_ncu.log 2>&1; echo "rc=$?"; tail -c 300 $out/bench_under_ncu.log; wc -l $out/launches.csv; gzip -f $out/launches.csv
