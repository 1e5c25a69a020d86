#!/bin/bash
# device builder: where its time goes at BASELINE size; kernel list by ncu; surface tests
tag=${1:-r2z}
out=gpurun_out/$tag
mkdir -p $out
echo "== build bench (trace)"; ZB_BUILD_TRACE=1 timeout 300 python scripts/build_bench.py 8 200000 > $out/build_bench.json 2> $out/build_trace.txt; echo "rc=$?"; cat $out/build_bench.json; tail -60 $out/build_trace.txt
echo "== ncu launch list of the builds"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/build_launches.csv python scripts/build_bench.py 8 1000 > $out/ncu_build.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, collections, sys
rows = [r for r in csv.reader(open("gpurun_out/%s/build_launches.csv" % (sys.argv[1] if len(sys.argv) > 1 else "r2z"))) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    a = agg.setdefault(r[ki][:70], [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]: print("%-72s %5d launches %10.3f ms" % (k, n, t/1e6))
PY
echo "== pytest surface + builder"; timeout 900 python -m pytest tests/test_zombie_surface.py tests/test_device_builder.py -m gpu -q -x -p no:cacheprovider --timeout 600 > $out/pytest.log 2>&1; echo "rc=$?"; tail -8 $out/pytest.log
