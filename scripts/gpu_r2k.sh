#!/bin/bash
tag=${1:-r2k}
out=gpurun_out/$tag
mkdir -p $out
echo "== wave check"; timeout 200 python scripts/wave_check.py 2>&1 | grep -v "^escaped" | tail -4 | tee $out/wave_check.log
echo "== pytest (solve tests)"; timeout 600 python -m pytest tests/test_gpu_solve.py tests/test_rws.py tests/test_cpp_api.py -m gpu -q -p no:cacheprovider --timeout 300 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest.log
run() { echo "== $1"; shift; env "$@" ZB_REPS=2 timeout 200 python scripts/profile_case.py 8 8192 64 2>&1 | tail -1; }
{
run "tail 0" ZB_WAVE_TAIL=0
run "tail 1024" ZB_WAVE_TAIL=1024
run "tail 4096" ZB_WAVE_TAIL=4096
run "tail 8192 (default)" ZB_DUMMY=1
run "tail 16384" ZB_WAVE_TAIL=16384
run "tail 32768" ZB_WAVE_TAIL=32768
run "tail 65536" ZB_WAVE_TAIL=65536
} 2>&1 | tee $out/ab.log
echo "== big batch default"; ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 32768 64 2>&1 | tail -1 | tee -a $out/ab.log
echo "== big batch tail 32768"; ZB_WAVE_TAIL=32768 ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 32768 64 2>&1 | tail -1 | tee -a $out/ab.log
echo "== launch list (tail 0)"
ZB_WAVE_TAIL=0 ZB_REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file $out/launches_tail0.csv python scripts/profile_case.py 8 8192 64 > $out/ncu_launch.log 2>&1; echo "rc=$?"; tail -2 $out/ncu_launch.log
