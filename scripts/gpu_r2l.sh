#!/bin/bash
tag=${1:-r2l}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest (queries + solve)"; timeout 900 python -m pytest tests/test_gpu_queries.py tests/test_gpu_queries_vs_reference.py tests/test_gpu_solve.py -m gpu -q -p no:cacheprovider --timeout 300 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest.log
run() { echo "== $1"; shift; env "$@" ZB_REPS=2 timeout 200 python scripts/profile_case.py 8 16384 64 2>&1 | tail -1; }
{
run "default" ZB_DUMMY=1
for v in occ5 occ6 occ7 occ8; do
run "$v" ZOMBIE_B200_LIB=$PWD/zombie_b200/variants/lib$v.so
done
run "occ6 slots 524288" ZOMBIE_B200_LIB=$PWD/zombie_b200/variants/libocc6.so ZB_WAVE_SLOTS=524288
run "occ7 slots 524288" ZOMBIE_B200_LIB=$PWD/zombie_b200/variants/libocc7.so ZB_WAVE_SLOTS=524288
} 2>&1 | tee $out/ab.log
