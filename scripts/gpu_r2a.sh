#!/bin/bash
# Round-2 first GPU pass: every GPU test (no -x), steady-state rate, one cheap ncu capture of the shipped walk kernel
# (sections only, one launch of ~1 s), A/B variants. Usage under gpurun: bash scripts/gpu_r2a.sh <tag>
tag=${1:-r2a}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
nproc > $out/nproc.txt
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -40 $out/pytest.log
echo "== steady"; ZB_REPS=2 timeout 200 python scripts/profile_case.py 8 8192 64 > $out/steady.log 2>&1; cat $out/steady.log | tail -3
echo "== ncu sections"
ZB_REPS=1 timeout 400 ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis \
    --section LaunchStats --section Occupancy --section SourceCounters --section InstructionStats --clock-control none --import-source on \
    -k regex:walkKernel -c 1 -o $out/walk_steady python scripts/profile_case.py 8 2048 64 > $out/ncu_walk.log 2>&1; echo "rc=$?"; tail -3 $out/ncu_walk.log
echo "== variants"; bash scripts/gpu_variants.sh 8192 > $out/variants.log 2>&1; cat $out/variants.log
ls -la $out
