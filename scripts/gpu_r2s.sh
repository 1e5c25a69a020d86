#!/bin/bash
# tie-rule parity test, then BASELINE configs C1/C2/C4/C5 at (a fraction of, then) their stated sizes
tag=${1:-r2s}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest (closest-point index vs Baseline, query suites)"; timeout 900 python -m pytest tests/test_gpu_queries_vs_reference.py tests/test_gpu_queries.py -m gpu -q -p no:cacheprovider --timeout 600 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $out/pytest.log
echo "== configs at 5 %"; timeout 600 python scripts/run_configs.py c1 c2 c4 c5 --scale 0.05 > $out/configs_5pct.jsonl 2> $out/configs_5pct.err; echo "rc=$?"; cat $out/configs_5pct.jsonl; tail -3 $out/configs_5pct.err
echo "== C1 full"; timeout 120 python scripts/run_configs.py c1 > $out/c1.json 2> $out/c1.err; echo "rc=$?"; cat $out/c1.json; tail -2 $out/c1.err
echo "== C5 full"; timeout 400 python scripts/run_configs.py c5 > $out/c5.json 2> $out/c5.err; echo "rc=$?"; cat $out/c5.json; tail -2 $out/c5.err
echo "== C4 full"; timeout 480 python scripts/run_configs.py c4 > $out/c4.json 2> $out/c4.err; echo "rc=$?"; cat $out/c4.json; tail -2 $out/c4.err
echo "== C2 full"; timeout 480 python scripts/run_configs.py c2 > $out/c2.json 2> $out/c2.err; echo "rc=$?"; cat $out/c2.json; tail -2 $out/c2.err
