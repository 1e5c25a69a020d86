#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (both arms), ncu launch list and one full capture of the
# walk kernel. Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh <tag>
tag=${1:-r1}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi > $out/smi.txt 2>&1
nproc > $out/nproc.txt
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/pytest.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $out/smoke.log
echo "== bench"; timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; cat $out/bench.json; tail -3 $out/bench.err
echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "rc=$?"; cat $out/bench_ref.json
echo "== ncu launches"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $out/ncu_launch_bench.log 2>&1; echo "rc=$?"
echo "== ncu full"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:walkKernel -s 1 -c 1 -o $out/walk_full \
    python bench.py --steps 1 --warmup 1 --no-cpu > $out/ncu_full_bench.log 2>&1; echo "rc=$?"
ls -la $out
