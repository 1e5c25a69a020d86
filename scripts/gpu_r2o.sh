#!/bin/bash
tag=${1:-r2o}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $out/pytest.log
run() { echo "== $1"; shift; env "$@" ZB_REPS=2 timeout 200 python scripts/profile_case.py 8 16384 64 2>&1 | tail -1; }
{
run "default (tickets, 6 CTAs/SM)" ZB_DUMMY=1
run "no tickets" ZB_WAVE_PERSIST=0
run "tickets x2 CTAs" ZB_WAVE_PERSIST=2
run "lists1024 (4 CTAs/SM), tickets" ZOMBIE_B200_LIB=$PWD/zombie_b200/variants/liblists1024.so
run "slots 524288" ZB_WAVE_SLOTS=524288
run "tail 4096" ZB_WAVE_TAIL=4096
} 2>&1 | tee $out/ab.log
echo "== big batch"; ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 32768 64 2>&1 | tail -1 | tee -a $out/ab.log
echo "== bench"; timeout 900 python bench.py --steps 2 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "rc=$?"; cat $out/bench.json; tail -3 $out/bench.err
