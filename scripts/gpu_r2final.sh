#!/bin/bash
# final code: both bench arms
out=gpurun_out/r2final
mkdir -p $out
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; echo "rc=$?"; tail -1 $out/bench.json | cut -c1-400
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "rc=$?"; tail -1 $out/bench_ref.json | cut -c1-200
