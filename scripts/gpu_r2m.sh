#!/bin/bash
tag=${1:-r2m}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest (full-size parity)"; timeout 900 python -m pytest tests/test_gpu_queries_vs_reference.py -m gpu -q -p no:cacheprovider --timeout 600 -k full_size > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest.log
run() { echo "== $1"; shift; env "$@" ZB_REPS=2 timeout 200 python scripts/profile_case.py 8 16384 64 2>&1 | tail -1; }
{
run "default" ZB_DUMMY=1
for v in occ5 occ6 occ7 occ8; do
run "$v" ZOMBIE_B200_LIB=$PWD/zombie_b200/variants/lib$v.so
done
} 2>&1 | tee $out/ab.log
