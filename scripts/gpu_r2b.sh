#!/bin/bash
# Round-2 second GPU pass: wavefront scheduler. GPU tests, steady-state A/B (sort on/off, slot counts, persistent kernel),
# ncu sections of the three wave kernels at a steady-state round. Usage under gpurun: bash scripts/gpu_r2b.sh <tag>
tag=${1:-r2b}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -30 $out/pytest.log
run() { echo "== $1"; shift; env "$@" ZB_REPS=2 timeout 200 python scripts/profile_case.py 8 8192 64 2>&1 | tail -2; }
run "wavefront default" ZB_DUMMY=1 | tee $out/ab.log
run "wavefront nosort" ZB_WAVE_SORT=0 | tee -a $out/ab.log
run "wavefront slots 131072" ZB_WAVE_SLOTS=131072 | tee -a $out/ab.log
run "wavefront slots 524288" ZB_WAVE_SLOTS=524288 | tee -a $out/ab.log
run "persistent" ZB_SCHEDULER_PERSISTENT=1 | tee -a $out/ab.log
echo "== big batch"; ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 32768 64 2>&1 | tail -1 | tee -a $out/ab.log
echo "== ncu sections (wave kernels, rounds 20..21)"
ZB_REPS=1 timeout 600 ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis \
    --section LaunchStats --section Occupancy --section SourceCounters --section InstructionStats --clock-control none --import-source on \
    -k regex:wave -s 60 -c 6 -o $out/wave_steady python scripts/profile_case.py 8 8192 64 > $out/ncu_wave.log 2>&1; echo "rc=$?"; tail -3 $out/ncu_wave.log
ls -la $out
