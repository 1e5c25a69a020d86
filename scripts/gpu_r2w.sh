#!/bin/bash
# child-parallel node rounds of the pooled star query, step-kernel occupancy variants; parity suites; ncu launch list of the bench command
tag=${1:-r2w}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest (queries vs reference, queries, solve)"; timeout 1500 python -m pytest tests/test_gpu_queries_vs_reference.py tests/test_gpu_queries.py tests/test_gpu_solve.py -m gpu -q -p no:cacheprovider --timeout 900 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $out/pytest.log
run() { echo "== $1"; shift; env "$@" ZB_REPS=2 timeout 200 python scripts/profile_case.py 8 16384 64 2>&1 | tail -1; }
{
run "child-parallel node rounds (default)" ZB_DUMMY=1
run "node-parallel node rounds" ZOMBIE_B200_LIB=zombie_b200/variants/libnode_parallel.so
run "step kernel 4 CTAs/SM" ZOMBIE_B200_LIB=zombie_b200/variants/libstep4.so
run "step kernel 5 CTAs/SM" ZOMBIE_B200_LIB=zombie_b200/variants/libstep5.so
run "step kernel 6 CTAs/SM" ZOMBIE_B200_LIB=zombie_b200/variants/libstep6.so
run "default again" ZB_DUMMY=1
} 2>&1 | tee $out/ab.log
echo "== big batch default"; ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 32768 64 2>&1 | tail -1 | tee -a $out/ab.log
echo "== ncu: one steady-state round, full set"
ZB_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:wave -s 50 -c 4 -o $out/wave_steady python scripts/profile_case.py 8 16384 64 > $out/ncu_wave.log 2>&1; echo "rc=$?"; tail -2 $out/ncu_wave.log
echo "== ncu launch list of a short bench"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --batch 4096 --no-cpu > $out/bench_under_ncu.log 2>&1; echo "rc=$?"; tail -c 300 $out/bench_under_ncu.log; wc -l $out/launches.csv
