#!/bin/bash
# Round-2 call E: programmatic dependent launch — parity suite, then A/B against a TAE_PDL=0 build (graph replay and eager).
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?: $(tail -n 1 gpurun_out/$name.log | cut -c1-200)"; }
summ() { python tools/print_bench.py $1 2>/dev/null | grep -E "^img/s" ; }
run gpu_tests python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider
grep -E "^FAILED|^ERROR" gpurun_out/gpu_tests.log | head
run smoke python __graft_entry__.py smoke
AB="--steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-secondary --no-encode --no-roofline"
for i in 1 2; do
  for cfg in 2 4; do
    for mode in "" "--no-graph"; do
      timeout 600 python bench.py --config $cfg $AB $mode > gpurun_out/pdl_c${cfg}${mode}_$i.log 2>&1; echo "== config $cfg PDL $mode"; summ gpurun_out/pdl_c${cfg}${mode}_$i.log
      TAE_B200_LIB=tae_b200/libtae_b200.nopdl.so timeout 600 python bench.py --config $cfg $AB $mode > gpurun_out/nopdl_c${cfg}${mode}_$i.log 2>&1; echo "== config $cfg no PDL $mode"; summ gpurun_out/nopdl_c${cfg}${mode}_$i.log
    done
  done
done
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/pdl_c5.log 2>&1; echo "== encode PDL"; summ gpurun_out/pdl_c5.log
TAE_B200_LIB=tae_b200/libtae_b200.nopdl.so timeout 600 python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/nopdl_c5.log 2>&1; echo "== encode no PDL"; summ gpurun_out/nopdl_c5.log
