#!/bin/bash
# Round-2 third GPU pass: wavefront scheduler A/B (CTA size of the query kernels, slot order, work-list size, slot count)
tag=${1:-r2c}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/pytest.log
run() { echo "== $1"; shift; env "$@" ZB_REPS=2 timeout 200 python scripts/profile_case.py 8 8192 64 2>&1 | tail -1; }
{
run "default (sort 0, block 128)" ZB_DUMMY=1
run "block 64" ZB_WAVE_BLOCK=64
run "block 32" ZB_WAVE_BLOCK=32
run "sort 2 (cost descending), block 32" ZB_WAVE_SORT=2 ZB_WAVE_BLOCK=32
run "sort 2 (cost descending), block 128" ZB_WAVE_SORT=2
run "sort 1 (morton), block 32" ZB_WAVE_SORT=1 ZB_WAVE_BLOCK=32
run "small lists, block 64" ZOMBIE_B200_LIB=$PWD/zombie_b200/variants/libsmall_lists.so ZB_WAVE_BLOCK=64
run "small lists, block 32" ZOMBIE_B200_LIB=$PWD/zombie_b200/variants/libsmall_lists.so ZB_WAVE_BLOCK=32
run "slots 524288, block 32" ZB_WAVE_SLOTS=524288 ZB_WAVE_BLOCK=32
run "slots 131072, block 32" ZB_WAVE_SLOTS=131072 ZB_WAVE_BLOCK=32
run "persistent" ZB_SCHEDULER_PERSISTENT=1
} 2>&1 | tee $out/ab.log
echo "== big batch, block 32"; ZB_WAVE_BLOCK=32 ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 32768 64 2>&1 | tail -1 | tee -a $out/ab.log
ls -la $out
