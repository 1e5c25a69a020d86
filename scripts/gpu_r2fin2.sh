#!/bin/bash
# same-box A/B of the scalar walk rate before / after the value-channel kernels (now separate instantiations); channel tests; ncu of one steady round
tag=${1:-r2fin2}
out=gpurun_out/$tag
mkdir -p $out
run() { echo "== $1"; shift; env "$@" ZB_REPS=2 timeout 200 python scripts/profile_case.py 8 16384 64 2>&1 | tail -1; }
{
run "current" ZB_DUMMY=1
run "before the channel hooks" ZOMBIE_B200_LIB=zombie_b200/variants/libprechannels.so
run "current" ZB_DUMMY=1
run "before the channel hooks" ZOMBIE_B200_LIB=zombie_b200/variants/libprechannels.so
} 2>&1 | tee $out/ab.log
echo "== pytest channels + surface"; timeout 900 python -m pytest tests/test_channels.py tests/test_zombie_surface.py -m gpu -q -x -p no:cacheprovider --timeout 600 > $out/pytest.log 2>&1; echo "rc=$?"; tail -5 $out/pytest.log
echo "== channel cost"; timeout 600 python scripts/channel_bench.py 8 16384 64 2>&1 | tail -1 | tee $out/channels.log
echo "== ncu: one steady-state round, full set"
ZB_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:wave -s 50 -c 4 -o $out/wave_steady python scripts/profile_case.py 8 16384 64 > $out/ncu_wave.log 2>&1; echo "rc=$?"; tail -2 $out/ncu_wave.log
