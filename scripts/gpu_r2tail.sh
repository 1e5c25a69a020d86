#!/bin/bash
# where the fixed ~0.8 s of a solve goes: tail threshold and slot count at a 16,384-point solve (1 M walks = one slot each)
tag=${1:-r2tail}
out=gpurun_out/$tag
mkdir -p $out
run() { echo "== $1"; shift; env "$@" ZB_REPS=2 timeout 200 python scripts/profile_case.py 8 16384 64 2>&1 | tail -1; }
{
run "default (tail 8192, 1 M slots)" ZB_DUMMY=1
run "tail 0 (rounds to the end)" ZB_WAVE_TAIL=0
run "tail 2048" ZB_WAVE_TAIL=2048
run "tail 32768" ZB_WAVE_TAIL=32768
run "tail 131072" ZB_WAVE_TAIL=131072
run "slots 524288" ZB_WAVE_SLOTS=524288
run "slots 262144" ZB_WAVE_SLOTS=262144
run "slots 262144, tail 32768" ZB_WAVE_SLOTS=262144 ZB_WAVE_TAIL=32768
} 2>&1 | tee $out/ab.log
