#!/bin/bash
tag=${1:-r2n}
out=gpurun_out/$tag
mkdir -p $out
echo "== star debug"; timeout 600 python tests/diagnostics/star_fullsize_debug.py $out/star_full.npz 2>&1 | tail -14 | tee $out/star_debug.log
echo "== pytest (grid pde, cpp api)"; timeout 900 python -m pytest tests/test_grid_pde.py tests/test_cpp_api.py tests/test_cpp_extras.py -m gpu -q -p no:cacheprovider --timeout 600 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest.log
run() { echo "== $1"; shift; env "$@" ZB_REPS=2 timeout 200 python scripts/profile_case.py 8 16384 64 2>&1 | tail -1; }
{
run "default" ZB_DUMMY=1
for v in occ6 occ6c7 occ6b4 occ6b16; do
run "$v" ZOMBIE_B200_LIB=$PWD/zombie_b200/variants/lib$v.so
done
} 2>&1 | tee $out/ab.log
