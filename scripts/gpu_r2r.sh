#!/bin/bash
# star-radius query split into two kernels: parity tests, A/B against the fused kernel, ncu of a steady round
tag=${1:-r2r}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest (solve, queries)"; timeout 900 python -m pytest tests/test_gpu_solve.py tests/test_gpu_queries.py tests/test_grid_pde.py -m gpu -q -p no:cacheprovider --timeout 600 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $out/pytest.log
run() { echo "== $1"; shift; env "$@" ZB_REPS=2 timeout 200 python scripts/profile_case.py 8 16384 64 2>&1 | tail -1; }
{
run "split (default)" ZB_DUMMY=1
run "fused" ZB_WAVE_STAR_SPLIT=0
run "split (again)" ZB_DUMMY=1
run "fused (again)" ZB_WAVE_STAR_SPLIT=0
} 2>&1 | tee $out/ab.log
echo "== big batch split"; ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 32768 64 2>&1 | tail -1 | tee -a $out/ab.log
echo "== big batch fused"; ZB_WAVE_STAR_SPLIT=0 ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 32768 64 2>&1 | tail -1 | tee -a $out/ab.log
echo "== ncu: one steady-state round, full set"
ZB_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:wave -s 50 -c 4 -o $out/wave_steady python scripts/profile_case.py 8 16384 64 > $out/ncu_wave.log 2>&1; echo "rc=$?"; tail -2 $out/ncu_wave.log
