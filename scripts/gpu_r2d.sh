#!/bin/bash
tag=${1:-r2d}
out=gpurun_out/$tag
mkdir -p $out
echo "== star debug"; timeout 120 tools/gpu_star_debug gpurun_in 2.0 0.6832276 -0.09269869 0.530396 0.631202459 0.19 2>&1 | tee $out/star_debug.log
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $out/pytest.log
run() { echo "== $1"; shift; env "$@" ZB_REPS=2 timeout 200 python scripts/profile_case.py 8 8192 64 2>&1 | tail -1; }
{
run "default (hash order, block 128)" ZB_DUMMY=1
run "block 32" ZB_WAVE_BLOCK=32
run "slots 524288" ZB_WAVE_SLOTS=524288
} 2>&1 | tee $out/ab.log
echo "== big batch"; ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 32768 64 2>&1 | tail -1 | tee -a $out/ab.log
echo "== big batch 524288 slots"; ZB_WAVE_SLOTS=524288 ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 32768 64 2>&1 | tail -1 | tee -a $out/ab.log
