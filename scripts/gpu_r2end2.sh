#!/bin/bash
# end of round 2, final code: the whole GPU suite in one go, smoke, the BASELINE configs at size
tag=${1:-r2end2}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest all"; timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 900 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== configs at size"; timeout 900 python scripts/run_configs.py > $out/configs.jsonl 2> $out/configs.err; echo "rc=$?"; cat $out/configs.jsonl | cut -c1-600; tail -2 $out/configs.err
