#!/bin/bash
# Round-2 call J (2 GPUs): CUDA-graph replay of the data-parallel step (NCCL all-reduces captured) against the eager step.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
AB="--steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-encode --no-secondary"
for mode in "" "--no-graph" "" "--no-graph"; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N $AB $mode > gpurun_out/ddp_graph$mode.log 2>&1
  echo "== N=$N $mode exit $?"; python tools/print_bench.py gpurun_out/ddp_graph$mode.log 2>/dev/null | grep -E "^img/s|ddp_check" || tail -n 15 gpurun_out/ddp_graph$mode.log
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $N --config 4 $AB > gpurun_out/ddp_graph_c4.log 2>&1
echo "== N=$N config 4 graph exit $?"; python tools/print_bench.py gpurun_out/ddp_graph_c4.log 2>/dev/null | grep -E "^img/s|ddp_check" || tail -n 15 gpurun_out/ddp_graph_c4.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus $N --config 4 $AB --no-graph > gpurun_out/ddp_eager_c4.log 2>&1
echo "== N=$N config 4 eager exit $?"; python tools/print_bench.py gpurun_out/ddp_eager_c4.log 2>/dev/null | grep -E "^img/s|ddp_check" || tail -n 15 gpurun_out/ddp_eager_c4.log
