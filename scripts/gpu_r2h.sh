#!/bin/bash
tag=${1:-r2h}
out=gpurun_out/$tag
mkdir -p $out
echo "== wave check"; ZB_WAVE_GROUPS=1 timeout 120 python scripts/wave_check.py 2>&1 | tail -4 | tee $out/wave_check.log
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $out/pytest.log
