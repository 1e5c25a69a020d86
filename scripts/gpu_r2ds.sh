#!/bin/bash
tag=${1:-r2ds}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest double-sided + related"; timeout 900 python -m pytest tests/test_double_sided_dirichlet.py tests/test_gpu_solve.py tests/test_channels.py tests/test_cpp_api.py tests/test_cpp_extras.py tests/test_grid_pde.py -m gpu -q -x -p no:cacheprovider --timeout 600 > $out/pytest.log 2>&1; echo "rc=$?"; tail -25 $out/pytest.log
run() { echo "== $1"; shift; env "$@" ZB_REPS=2 timeout 200 python scripts/profile_case.py 8 16384 64 2>&1 | tail -1; }
{
run "current" ZB_DUMMY=1
run "before per-side data" ZOMBIE_B200_LIB=zombie_b200/variants/libanyhit_unsorted.so
run "current" ZB_DUMMY=1
} 2>&1 | tee $out/ab.log
