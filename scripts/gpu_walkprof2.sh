#!/bin/bash
# Steady-state profile of walkKernel (several walks per lane): source counters + scheduler / warp-state sections only,
# so that ncu needs few replays of the multi-second launch. Usage under gpurun: bash scripts/gpu_walkprof2.sh <tag> [points]
tag=${1:-w2}
pts=${2:-8192}
out=gpurun_out/$tag
mkdir -p $out
timeout 800 ncu --section SourceCounters --section SchedulerStats --section WarpStateStats --section SpeedOfLight --section Occupancy --section LaunchStats \
    --clock-control none --import-source on -k regex:walkKernel -s 1 -c 1 -o $out/walk \
    python scripts/profile_case.py 8 $pts 64 > $out/ncu_walk.log 2>&1; echo "rc=$?"; tail -4 $out/ncu_walk.log
ls -la $out
