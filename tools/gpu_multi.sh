#!/bin/bash
# Multi-GPU runs (N = the GPUs of the box):
#   gpurun --gpus 2 --timeout 600  -- 'bash tools/gpu_multi.sh test'          NCCL data-parallel parity tests + a short N-rank line
#   gpurun --gpus 8 --timeout 900  -- 'bash tools/gpu_multi.sh bench [args]'  the default N-rank line (add --no-secondary for a short run)
# Every step runs under its own short timeout: a rank stuck in a collective must not hold the box.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
line() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N "${@:2}"; }
case "${1:-test}" in
  test)
    timeout 250 python -m pytest tests/test_gpu_train_loop.py -m gpu -q -k ddp --timeout 200 -p no:cacheprovider > gpurun_out/ddp_tests.log 2>&1
    echo "ddp tests exit $?: $(tail -n 1 gpurun_out/ddp_tests.log)"
    line 250 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-secondary > gpurun_out/bench$N.log 2>&1; echo "bench N=$N exit $?"
    python tools/print_bench.py gpurun_out/bench$N.log | grep -E "^img/s|ddp_check|encode|gemm TF" ;;
  bench)
    line 800 --no-cpu-baseline --no-gpu-reference "${@:2}" > gpurun_out/bench$N.log 2>&1; echo "bench N=$N exit $?"
    python tools/print_bench.py gpurun_out/bench$N.log ;;
esac
