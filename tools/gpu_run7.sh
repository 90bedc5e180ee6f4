#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary7.txt; tail -3 gpurun_out/$name.log; }
run bench2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3
run ref2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1
