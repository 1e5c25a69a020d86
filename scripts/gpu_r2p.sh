#!/bin/bash
tag=${1:-r2p}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest (rws, queries vs reference)"; timeout 1200 python -m pytest tests/test_rws.py tests/test_gpu_queries_vs_reference.py tests/test_cpp_extras.py tests/test_bvc_pipeline.py -m gpu -q -p no:cacheprovider --timeout 600 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $out/pytest.log
run() { echo "== $1"; shift; env "$@" ZB_REPS=2 timeout 200 python scripts/profile_case.py 8 16384 64 2>&1 | tail -1; }
{
run "default (524288 slots)" ZB_DUMMY=1
run "slots 262144" ZB_WAVE_SLOTS=262144
run "slots 1048576" ZB_WAVE_SLOTS=1048576
run "tail 16384" ZB_WAVE_TAIL=16384
} 2>&1 | tee $out/ab.log
echo "== big batch"; ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 32768 64 2>&1 | tail -1 | tee -a $out/ab.log
echo "== big batch 1M slots"; ZB_WAVE_SLOTS=1048576 ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 32768 64 2>&1 | tail -1 | tee -a $out/ab.log
echo "== ncu: one steady-state round, full set"
ZB_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:wave -s 40 -c 3 -o $out/wave_steady python scripts/profile_case.py 8 16384 64 > $out/ncu_wave.log 2>&1; echo "rc=$?"; tail -2 $out/ncu_wave.log
echo "== ncu launch list"
ZB_REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file $out/launches.csv python scripts/profile_case.py 8 16384 64 > $out/ncu_launch.log 2>&1; echo "rc=$?"; tail -1 $out/ncu_launch.log
