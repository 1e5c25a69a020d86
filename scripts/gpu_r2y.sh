#!/bin/bash
# on-device hierarchy builder: its GPU tests, the whole GPU suite, walk rate on a device-built hierarchy vs the host SAH one
tag=${1:-r2y}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest device builder"; timeout 900 python -m pytest tests/test_device_builder.py -m gpu -q -s -p no:cacheprovider --timeout 600 > $out/pytest_builder.log 2>&1; echo "rc=$?"; tail -25 $out/pytest_builder.log
echo "== A/B walk rate"; {
echo "host SAH:"; ZB_REPS=2 timeout 300 python scripts/profile_case.py 8 16384 64 2>&1 | tail -1
echo "device:"; ZOMBIE_B200_BVH_BUILDER=device ZB_REPS=2 timeout 300 python scripts/profile_case.py 8 16384 64 2>&1 | tail -1
} 2>&1 | tee $out/ab.log
echo "== pytest all"; timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 900 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $out/pytest.log
