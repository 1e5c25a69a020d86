#!/bin/bash
# where a first full-size device build spends its time inside a long-lived process; compute-sanitizer on the new kernels; rest of the GPU suite
tag=${1:-r2san}
out=gpurun_out/$tag
mkdir -p $out
echo "== builder tests with the stage trace"; ZB_BUILD_TRACE=1 timeout 900 python -m pytest tests/test_device_builder.py -m gpu -q -s -p no:cacheprovider --timeout 600 > $out/pytest_builder.log 2>&1; echo "rc=$?"; grep -n "build of\|passed\|failed" $out/pytest_builder.log | tail -5; grep -n "bvh build" $out/pytest_builder.log | awk '{ if ($(NF-1) > 5.0) print }' | tail -20
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_case.py > $out/memcheck.log 2>&1; echo "rc=$?"; tail -4 $out/memcheck.log
echo "== racecheck"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_case.py > $out/racecheck.log 2>&1; echo "rc=$?"; tail -4 $out/racecheck.log
echo "== pytest, the files after test_device_builder"; timeout 1800 python -m pytest tests/test_g*.py tests/test_r*.py tests/test_z*.py tests/test_cpp_extras.py -m gpu -q -p no:cacheprovider --timeout 900 > $out/pytest_rest.log 2>&1; echo "rc=$?"; tail -4 $out/pytest_rest.log
