#!/bin/bash
# Round-2 call G: per-problem tile width of the forward GEMMs — parity suite, then A/B against 256-wide tiles everywhere.
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?: $(tail -n 1 gpurun_out/$name.log | cut -c1-200)"; }
summ() { python tools/print_bench.py $1 2>/dev/null | grep -E "^img/s|gemm TF|bf16 |f32_resid|bf16_gelu" ; }
run gpu_tests python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider
grep -E "^FAILED|^ERROR" gpurun_out/gpu_tests.log | head
AB="--steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-secondary --no-encode"
for i in 1 2; do
  for v in default tile256; do
    [ "$v" = default ] && unset TAE_B200_LIB || export TAE_B200_LIB=tae_b200/libtae_b200.$v.so
    timeout 600 python bench.py --config 5 --steps 20 --warmup 5 > gpurun_out/tn_c5_${v}_$i.log 2>&1; echo "== encode patch64 $v"; summ gpurun_out/tn_c5_${v}_$i.log
    timeout 600 python bench.py --config 4 $AB > gpurun_out/tn_c4_${v}_$i.log 2>&1; echo "== config 4 $v"; summ gpurun_out/tn_c4_${v}_$i.log
    timeout 600 python bench.py --model tae_patch64_vocab4096_px256 $AB > gpurun_out/tn_p64_${v}_$i.log 2>&1; echo "== patch64 train $v"; summ gpurun_out/tn_p64_${v}_$i.log
  done
done
unset TAE_B200_LIB
timeout 600 python bench.py --config 2 $AB > gpurun_out/tn_c2_default.log 2>&1; echo "== config 2 default"; summ gpurun_out/tn_c2_default.log
