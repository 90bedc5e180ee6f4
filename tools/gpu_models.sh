#!/bin/bash
# 1-GPU training-step throughput of the other model sizes (BASELINE.json configs 3-5), with the per-kernel-family detail.
mkdir -p gpurun_out
for m in tae_patch32_vocab1024_px256 tae_patch64_vocab4096_px256 tae_patch128_vocab16384_px256; do
  timeout 600 python bench.py --model $m --steps 3 --warmup 3 --no-cpu-baseline --no-encode > gpurun_out/bench_$m.log 2>&1; echo "$m exit $?"
  python tools/print_bench.py gpurun_out/bench_$m.log
done
