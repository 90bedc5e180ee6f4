#!/bin/bash
# Round-2 call K (2 GPUs): NCCL parity tests incl. graph replay under data parallelism; 2-rank default line with all configs.
mkdir -p gpurun_out
timeout 800 python -m pytest tests/test_gpu_train_loop.py -m gpu -q --timeout 700 -p no:cacheprovider > gpurun_out/ddp_tests.log 2>&1
echo "ddp tests exit $?"; tail -n 3 gpurun_out/ddp_tests.log; grep -E "FAIL|Error" gpurun_out/ddp_tests.log | head -20
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench2_full.log 2>&1
echo "== N=$N full line exit $?"; python tools/print_bench.py gpurun_out/bench2_full.log 2>/dev/null | grep -E "^img/s|secondary|ddp_check|encode|gemm TF" || tail -n 25 gpurun_out/bench2_full.log
