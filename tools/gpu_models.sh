#!/bin/bash
# 1-GPU lines of the other BASELINE configs with the per-kernel-family detail: 3 (patch32 training), 4 (patch128 training),
# patch64 training, 5 (patch64 encode).
mkdir -p gpurun_out
F="--steps 6 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-secondary --no-encode"
for c in 3 4; do
  timeout 600 python bench.py --config $c $F > gpurun_out/bench_c$c.log 2>&1; echo "config $c exit $?"; python tools/print_bench.py gpurun_out/bench_c$c.log
done
timeout 600 python bench.py --model tae_patch64_vocab4096_px256 $F > gpurun_out/bench_p64.log 2>&1; echo "patch64 train exit $?"; python tools/print_bench.py gpurun_out/bench_p64.log
timeout 600 python bench.py --config 5 --steps 20 --warmup 5 > gpurun_out/bench_c5.log 2>&1; echo "config 5 exit $?"; python tools/print_bench.py gpurun_out/bench_c5.log
