#!/bin/bash
# device builder after the CTA-per-slot cone pass: build bench; then the whole GPU suite and the bench (both arms)
tag=${1:-r2z2}
out=gpurun_out/$tag
mkdir -p $out
echo "== build bench (trace)"; ZB_BUILD_TRACE=1 timeout 300 python scripts/build_bench.py 8 200000 > $out/build_bench.json 2> $out/build_trace.txt; echo "rc=$?"; cat $out/build_bench.json; tail -16 $out/build_trace.txt
echo "== build bench C5 size (5.2 M triangles)"; timeout 300 python scripts/build_bench.py 9 100000 > $out/build_bench_9.json 2> $out/build_trace_9.txt; echo "rc=$?"; cat $out/build_bench_9.json
echo "== pytest all"; timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 900 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $out/pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; echo "rc=$?"; cat $out/bench.json; tail -3 $out/bench.err
echo "== bench reference"; timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "rc=$?"; cat $out/bench_ref.json
