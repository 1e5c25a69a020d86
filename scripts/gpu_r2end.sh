#!/bin/bash
# end of round 2: the whole GPU suite, smoke, both bench arms, the other configs at size
tag=${1:-r2end}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi > $out/smi.txt 2>&1
echo "== pytest all"; timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 900 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $out/pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; echo "rc=$?"; cat $out/bench.json; tail -3 $out/bench.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "rc=$?"; cat $out/bench_ref.json
