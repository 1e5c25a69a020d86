#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (both arms), ncu launch list and one full capture of the
# walk kernel. Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh <tag> [stages]
tag=${1:-r1}
stages=${2:-pytest,smoke,bench,ref,launches,full}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi > $out/smi.txt 2>&1
nproc > $out/nproc.txt
has() { [[ ",$stages," == *",$1,"* ]]; }
if has pytest; then echo "== pytest"; timeout 420 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $out/pytest.log; fi
if has smoke; then echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $out/smoke.log; fi
if has bench; then echo "== bench"; timeout 420 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; cat $out/bench.json; tail -3 $out/bench.err; fi
if has ref; then echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "rc=$?"; cat $out/bench_ref.json; fi
if has launches; then echo "== ncu launches"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 1 --batch 4096 --no-cpu > $out/ncu_launch_bench.log 2>&1; echo "rc=$?"; fi
if has full; then echo "== ncu full"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:walkKernel -s 1 -c 1 -o $out/walk_full \
    python bench.py --steps 1 --warmup 1 --batch 4096 --no-cpu > $out/ncu_full_bench.log 2>&1; echo "rc=$?"; fi
ls -la $out
