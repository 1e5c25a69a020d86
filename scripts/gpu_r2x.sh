#!/bin/bash
# reflectance-only redo: parity suites, A/B of the redo triggers, batch size of the bench
tag=${1:-r2x}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest (queries vs reference, queries, solve)"; timeout 1500 python -m pytest tests/test_gpu_queries_vs_reference.py tests/test_gpu_queries.py tests/test_gpu_solve.py -m gpu -q -p no:cacheprovider --timeout 900 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $out/pytest.log
run() { echo "== $1"; shift; env "$@" ZB_REPS=2 timeout 200 python scripts/profile_case.py 8 16384 64 2>&1 | tail -1; }
{
run "triggers 0.02 / 0.1, reflectance-only redo (default)" ZB_DUMMY=1
run "triggers 0.05 / 0.3" ZOMBIE_B200_LIB=zombie_b200/variants/libthr_orig.so
run "triggers 0.01 / 0.05" ZOMBIE_B200_LIB=zombie_b200/variants/libthr_tight.so
run "default again" ZB_DUMMY=1
} 2>&1 | tee $out/ab.log
echo "== batch 32768"; ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 32768 64 2>&1 | tail -1 | tee -a $out/ab.log
echo "== batch 65536"; ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 65536 64 2>&1 | tail -1 | tee -a $out/ab.log
echo "== batch 65536, tail 16384"; ZB_WAVE_TAIL=16384 ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 65536 64 2>&1 | tail -1 | tee -a $out/ab.log
echo "== batch 65536, tail 4096"; ZB_WAVE_TAIL=4096 ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 65536 64 2>&1 | tail -1 | tee -a $out/ab.log
