#!/bin/bash
# full GPU suite in natural order (no -x), smoke, default bench (both arms)
tag=${1:-r2q}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest (all, no -x)"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $out/pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke.log
echo "== bench"; timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; echo "rc=$?"; cat $out/bench.json; tail -3 $out/bench.err
echo "== bench reference"; timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "rc=$?"; cat $out/bench_ref.json
