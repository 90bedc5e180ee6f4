#!/bin/bash
# Round-2 call C (1 GPU): side-stream weight gradients A/B on the short-grid configs, CUDA graph on config 4, tests.
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?: $(tail -n 1 gpurun_out/$name.log | cut -c1-200)"; }
summ() { python tools/print_bench.py $1 2>/dev/null | grep -E "^img/s|gemm TF|layernorm|attention|adamw|f32_acc|bf16 " ; }
run gpu_tests python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider
grep -E "^FAILED|^ERROR" gpurun_out/gpu_tests.log | head
AB="--steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-secondary --no-encode"
for cfg in 4 3; do
  for ov in 0 8192 65536; do
    timeout 600 python bench.py --config $cfg $AB --wgrad-overlap-rows $ov > gpurun_out/c${cfg}_ov$ov.log 2>&1; echo "== config $cfg overlap rows $ov"; summ gpurun_out/c${cfg}_ov$ov.log
  done
done
for ov in 0 8192; do
  timeout 600 python bench.py --model tae_patch64_vocab4096_px256 $AB --wgrad-overlap-rows $ov > gpurun_out/p64_ov$ov.log 2>&1; echo "== patch64 train overlap rows $ov"; summ gpurun_out/p64_ov$ov.log
done
timeout 600 python bench.py --config 4 $AB --graph > gpurun_out/c4_graph.log 2>&1; echo "== config 4 graph"; summ gpurun_out/c4_graph.log
timeout 600 python bench.py --config 4 $AB --graph --wgrad-overlap-rows 0 > gpurun_out/c4_graph_ov0.log 2>&1; echo "== config 4 graph, no overlap"; summ gpurun_out/c4_graph_ov0.log
