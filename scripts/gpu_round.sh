#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (both arms), ncu launch list and one full capture of the
# walk kernel. Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh <tag> [stages]
# (extra stage, not in the default list: variants = steady-state rate of every zombie_b200/variants/lib*.so)
tag=${1:-r1}
stages=${2:-pytest,smoke,bench,launches,full,ref,bvc}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi > $out/smi.txt 2>&1
nproc > $out/nproc.txt
has() { [[ ",$stages," == *",$1,"* ]]; }
if has pytest; then echo "== pytest"; timeout 300 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $out/pytest.log; fi
if has smoke; then echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $out/smoke.log; fi
if has bench; then echo "== bench"; timeout 330 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; cat $out/bench.json; tail -3 $out/bench.err; fi
if has launches; then echo "== ncu launches"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 1 --batch 2048 --no-cpu > $out/ncu_launch_bench.log 2>&1; echo "rc=$?"; fi
if has full; then echo "== ncu full"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:walkKernel -c 1 -o $out/walk_full \
    env ZB_REPS=1 python scripts/profile_case.py 8 512 64 > $out/ncu_full.log 2>&1; echo "rc=$?"; tail -2 $out/ncu_full.log; fi
if has ref; then echo "== bench reference"; timeout 150 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "rc=$?"; cat $out/bench_ref.json; fi
if has bvc; then echo "== bvc"; timeout 120 python scripts/bench_bvc.py --eval 2000000 --steps 2 > $out/bench_bvc.json 2> $out/bench_bvc.err; echo "rc=$?"; cat $out/bench_bvc.json; tail -2 $out/bench_bvc.err; fi
if has variants; then echo "== variants (build them first, here: bash scripts/build_variants.sh)"; bash scripts/gpu_variants.sh 8192 > $out/variants.log 2>&1; cat $out/variants.log; fi
ls -la $out
