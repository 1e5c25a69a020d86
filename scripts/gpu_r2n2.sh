#!/bin/bash
# two GPUs: the bench line with the NCCL gather in the timed region (weak and strong), and the sharding tests
tag=${1:-r2n2}
out=gpurun_out/$tag
mkdir -p $out
echo "== bench 2 GPUs weak"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $out/bench_2gpu_weak.json 2> $out/bench_2gpu_weak.err; echo "rc=$?"; tail -1 $out/bench_2gpu_weak.json; tail -3 $out/bench_2gpu_weak.err
echo "== bench 2 GPUs reference arm"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $out/bench_2gpu_ref.json 2> $out/bench_2gpu_ref.err; echo "rc=$?"; tail -1 $out/bench_2gpu_ref.json
