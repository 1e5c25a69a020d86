#!/bin/bash
# Profiles the cooperative queries in isolation (scripts/query_bench.py) — timings, then one ncu --set full capture
# of closestPointKernel and starRadiusKernel. Usage under gpurun: bash scripts/gpu_queries.sh <tag>
tag=${1:-q1}
out=gpurun_out/$tag
mkdir -p $out
echo "== timings"; timeout 200 python scripts/query_bench.py 262144 8 > $out/query_bench.log 2>&1; cat $out/query_bench.log
echo "== ncu"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:'closestPointKernel|starRadiusKernel' -s 2 -c 3 -o $out/queries \
    python scripts/query_bench.py 65536 8 > $out/ncu_queries.log 2>&1; echo "rc=$?"; tail -3 $out/ncu_queries.log
ls -la $out
