#!/bin/bash
# usage: call_scale.sh N   (under gpurun --gpus N)
set -u
N=$1; O=gpurun_out/r02s; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > $O/bench_n$N.json 2> $O/bench_n$N.err
echo "rc=$?"; tail -c 600 $O/bench_n$N.json
