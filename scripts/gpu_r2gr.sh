#!/bin/bash
tag=${1:-r2gr}
out=gpurun_out/$tag
mkdir -p $out
run() { echo "== $1"; shift; env "$@" ZB_REPS=2 timeout 300 python scripts/profile_case.py 8 $N 64 2>&1 | tail -1; }
{
N=16384
run "16K: 1 group" ZB_DUMMY=1
run "16K: 2 groups" ZB_WAVE_GROUPS=2
run "16K: 4 groups" ZB_WAVE_GROUPS=4
N=65536
run "64K: 1 group" ZB_DUMMY=1
run "64K: 2 groups" ZB_WAVE_GROUPS=2
run "64K: 4 groups" ZB_WAVE_GROUPS=4
} 2>&1 | tee $out/ab.log
