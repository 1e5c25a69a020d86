#!/bin/bash
# Walk-kernel diagnosis: timing by enabled contribution (scripts/walk_timing.py) and one ncu --set full capture of
# walkKernel on a small batch of the headline workload. Usage under gpurun: bash scripts/gpu_walkprof.sh <tag> [points]
tag=${1:-w1}
pts=${2:-512}
out=gpurun_out/$tag
mkdir -p $out
echo "== timing"; timeout 400 python scripts/walk_timing.py 8 1024 > $out/walk_timing.log 2>&1; cat $out/walk_timing.log
echo "== profile case"; timeout 200 python scripts/profile_case.py 8 4096 64 > $out/profile_case.log 2>&1; cat $out/profile_case.log
echo "== ncu"; timeout 500 ncu --set full --clock-control none --import-source on -k regex:walkKernel -s 1 -c 1 -o $out/walk \
    python scripts/profile_case.py 8 $pts 64 > $out/ncu_walk.log 2>&1; echo "rc=$?"; tail -3 $out/ncu_walk.log
ls -la $out
