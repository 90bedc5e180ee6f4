#!/bin/bash
# Round-2 multi-GPU call: (N=2) NCCL data-parallel parity tests; (N=8) scaling A/B of NCCL CTA caps and the full line.
#   gpurun --gpus 2 --timeout 900  -- 'bash tools/gpu_r2d_multi.sh test2'
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_r2d_multi.sh ab8'
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
trun() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
summ() { python tools/print_bench.py $1 2>/dev/null | grep -E "^img/s|gemm TF|f32_acc|layernorm_bwd|secondary|ddp_check|encode" ; }
case "$1" in
  test2)
    timeout 800 python -m pytest tests/test_gpu_train_loop.py -m gpu -q --timeout 700 -p no:cacheprovider > gpurun_out/ddp_tests.log 2>&1
    echo "ddp tests exit $?"; tail -n 15 gpurun_out/ddp_tests.log
    AB="--steps 6 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-encode"
    timeout 600 bash -c "$(declare -f trun); N=$N; trun 29541 $AB" > gpurun_out/bench2.log 2>&1; echo "bench N=$N exit $?"; summ gpurun_out/bench2.log
    ;;
  ab8)
    AB="--steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-encode --no-secondary"
    i=0
    for ctas in default 8 16 4; do
      i=$((i+1))
      if [ "$ctas" = default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$ctas; fi
      timeout 400 bash -c "$(declare -f trun); N=$N; trun $((29550+i)) $AB" > gpurun_out/bench8_ctas_$ctas.log 2>&1
      echo "== N=$N NCCL_MAX_CTAS=$ctas exit $?"; summ gpurun_out/bench8_ctas_$ctas.log
    done
    unset NCCL_MAX_CTAS
    [ -n "$2" ] && export NCCL_MAX_CTAS=$2
    timeout 900 bash -c "$(declare -f trun); N=$N; trun 29560 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference" > gpurun_out/bench8_full.log 2>&1
    echo "== N=$N full line (NCCL_MAX_CTAS=${NCCL_MAX_CTAS:-default}) exit $?"; python tools/print_bench.py gpurun_out/bench8_full.log
    ;;
esac
