#!/bin/bash
tag=${1:-r2i}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $out/pytest.log
echo "== wave check"; timeout 120 python scripts/wave_check.py 2>&1 | grep -v "^escaped" | tail -4 | tee $out/wave_check.log
run() { echo "== $1"; shift; env "$@" ZB_REPS=2 timeout 200 python scripts/profile_case.py 8 8192 64 2>&1 | tail -1; }
{
run "default" ZB_DUMMY=1
run "slots 524288" ZB_WAVE_SLOTS=524288
} 2>&1 | tee $out/ab.log
echo "== big batch 524288 slots"; ZB_WAVE_SLOTS=524288 ZB_REPS=1 timeout 300 python scripts/profile_case.py 8 32768 64 2>&1 | tail -1 | tee -a $out/ab.log
echo "== ncu sections (wave kernels at a steady-state round)"
ZB_REPS=1 timeout 600 ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis \
    --section LaunchStats --section Occupancy --section SourceCounters --section InstructionStats --clock-control none --import-source on \
    -k regex:wave -s 60 -c 3 -o $out/wave_steady python scripts/profile_case.py 8 8192 64 > $out/ncu_wave.log 2>&1; echo "rc=$?"; tail -3 $out/ncu_wave.log
