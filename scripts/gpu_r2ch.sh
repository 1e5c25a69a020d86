#!/bin/bash
# vector-valued solves: their tests; cost of extra channels on the headline workload; regression check of the scalar rate
tag=${1:-r2ch}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest channels + surface"; timeout 900 python -m pytest tests/test_channels.py tests/test_zombie_surface.py -m gpu -q -x -p no:cacheprovider --timeout 600 > $out/pytest.log 2>&1; echo "rc=$?"; tail -30 $out/pytest.log
echo "== scalar rate (regression check)"; ZB_REPS=2 timeout 300 python scripts/profile_case.py 8 16384 64 2>&1 | tail -1 | tee $out/scalar.log
echo "== channel cost"; timeout 600 python scripts/channel_bench.py 8 16384 64 2>&1 | tail -3 | tee $out/channels.log
