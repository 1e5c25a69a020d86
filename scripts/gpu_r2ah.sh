#!/bin/bash
# occlusion rays without the sorted child insertion: A/B on one box
tag=${1:-r2ah}
out=gpurun_out/$tag
mkdir -p $out
run() { echo "== $1"; shift; env "$@" ZB_REPS=2 timeout 200 python scripts/profile_case.py 8 16384 64 2>&1 | tail -1; }
{
run "sorted (HEAD)" ZB_DUMMY=1
run "any-hit unsorted" ZOMBIE_B200_LIB=zombie_b200/variants/libanyhit_unsorted.so
run "sorted (HEAD)" ZB_DUMMY=1
run "any-hit unsorted" ZOMBIE_B200_LIB=zombie_b200/variants/libanyhit_unsorted.so
} 2>&1 | tee $out/ab.log
echo "== big batch HEAD"; ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 65536 64 2>&1 | tail -1 | tee -a $out/ab.log
echo "== big batch any-hit unsorted"; ZOMBIE_B200_LIB=zombie_b200/variants/libanyhit_unsorted.so ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 65536 64 2>&1 | tail -1 | tee -a $out/ab.log
