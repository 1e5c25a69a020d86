#!/bin/bash
# two GPUs, final code: weak and strong scaling bench lines (NCCL gather in the timed region)
tag=${1:-r2n2s}
out=gpurun_out/$tag
mkdir -p $out
echo "== weak"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > $out/weak.json 2> $out/weak.err; echo "rc=$?"; tail -1 $out/weak.json | cut -c1-260
echo "== strong (65,536 points per step in total)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 3 --warmup 3 --scaling strong --no-cpu > $out/strong.json 2> $out/strong.err; echo "rc=$?"; tail -1 $out/strong.json | cut -c1-260
