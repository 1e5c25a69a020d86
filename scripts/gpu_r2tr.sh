#!/bin/bash
tag=${1:-r2tr}
out=gpurun_out/$tag
mkdir -p $out
ZB_WAVE_TRACE=1 ZB_REPS=1 timeout 200 python scripts/profile_case.py 8 16384 64 > $out/t16.log 2> $out/t16.err; tail -1 $out/t16.log
ZB_WAVE_TRACE=1 ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 65536 64 > $out/t64.log 2> $out/t64.err; tail -1 $out/t64.log
grep -c "wave round" $out/t16.err $out/t64.err
