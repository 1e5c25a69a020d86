#!/bin/bash
# two-GPU bench: NCCL all-gather of the statistics inside the timed region (weak and strong scaling), then the N=2 reference arm contract
tag=${1:-r2t}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi -L > $out/smi.txt
echo "== weak, 2 GPUs"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --batch 16384 --no-cpu > $out/weak2.json 2> $out/weak2.err; echo "rc=$?"; cat $out/weak2.json; tail -3 $out/weak2.err
echo "== strong, 2 GPUs"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --batch 32768 --scaling strong --no-cpu > $out/strong2.json 2> $out/strong2.err; echo "rc=$?"; cat $out/strong2.json; tail -3 $out/strong2.err
echo "== 1 GPU same batch"; timeout 600 python bench.py --steps 3 --warmup 3 --batch 16384 --no-cpu > $out/one.json 2> $out/one.err; echo "rc=$?"; cat $out/one.json; tail -3 $out/one.err
echo "== sharding test on 2 GPUs (nccl)"; timeout 300 python -m pytest tests/test_sharding.py -q -p no:cacheprovider -m gpu > $out/pytest.log 2>&1; tail -3 $out/pytest.log
