#!/bin/bash
# deferred reference-ordered redo of star primitives + tighter redo triggers: parity suites, A/B, ncu of a steady round
tag=${1:-r2u}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest (queries vs reference, solve, new insideDomain tests, bvc pipeline, cpp)"; timeout 1500 python -m pytest tests/test_gpu_queries_vs_reference.py tests/test_gpu_queries.py tests/test_gpu_solve.py tests/test_bvc_pipeline.py tests/test_cpp_api.py -m gpu -q -p no:cacheprovider --timeout 900 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $out/pytest.log
run() { echo "== $1"; shift; env "$@" ZB_REPS=2 timeout 200 python scripts/profile_case.py 8 16384 64 2>&1 | tail -1; }
{
run "deferred redo, triggers 0.01 / 0.05 (default)" ZB_DUMMY=1
run "inline redo, triggers 0.01 / 0.05" ZOMBIE_B200_LIB=zombie_b200/variants/libredo_inline.so
run "deferred redo, triggers 0.05 / 0.3 (old)" ZOMBIE_B200_LIB=zombie_b200/variants/libredo_old_thresholds.so
run "default again" ZB_DUMMY=1
} 2>&1 | tee $out/ab.log
echo "== big batch default"; ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 32768 64 2>&1 | tail -1 | tee -a $out/ab.log
echo "== ncu: one steady-state round, full set"
ZB_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:wave -s 50 -c 4 -o $out/wave_steady python scripts/profile_case.py 8 16384 64 > $out/ncu_wave.log 2>&1; echo "rc=$?"; tail -2 $out/ncu_wave.log
