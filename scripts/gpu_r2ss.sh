#!/bin/bash
# the step kernel's own order (sorted by star-radius class): A/B on one box, identical estimates, GPU solve tests
tag=${1:-r2ss}
out=gpurun_out/$tag
mkdir -p $out
run() { echo "== $1"; shift; env "$@" ZB_REPS=2 timeout 200 python scripts/profile_case.py 8 16384 64 2>&1 | tail -1; }
{
run "step sort (default)" ZB_DUMMY=1
run "no step sort" ZB_WAVE_STEP_SORT=0
run "step sort (default)" ZB_DUMMY=1
run "no step sort" ZB_WAVE_STEP_SORT=0
} 2>&1 | tee $out/ab.log
echo "== big batch"; ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 65536 64 2>&1 | tail -1 | tee -a $out/ab.log
echo "== big batch, no step sort"; ZB_WAVE_STEP_SORT=0 ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 65536 64 2>&1 | tail -1 | tee -a $out/ab.log
echo "== pytest solve"; timeout 900 python -m pytest tests/test_gpu_solve.py tests/test_channels.py -m gpu -q -x -p no:cacheprovider --timeout 600 > $out/pytest.log 2>&1; echo "rc=$?"; tail -3 $out/pytest.log
