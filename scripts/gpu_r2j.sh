#!/bin/bash
tag=${1:-r2j}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/smi.txt; nproc >> $out/smi.txt
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $out/pytest.log
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "rc=$?"; cat $out/bench.json; tail -3 $out/bench.err
