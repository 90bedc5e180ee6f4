#!/bin/bash
# Round-2 call B: full GPU suite on HEAD (ring forward / row-dot TMA epilogue defaults, new N=64 backward, row-group LayerNorm
# backward), short-grid probes, exp2-polynomial variants of the ring forward, and the complete default bench line.
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?: $(tail -n 1 gpurun_out/$name.log | cut -c1-200)"; }
rm -f gpurun_out/grad_parity.txt
run gpu_tests python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider
grep -E "^FAILED|^ERROR" gpurun_out/gpu_tests.log | head
run smoke python __graft_entry__.py smoke
run probe_small python tools/gpu_probe.py --small-grids; cat gpurun_out/probe_small.log
run probe_attn python tools/gpu_probe.py --attn-only; grep attention gpurun_out/probe_attn.log
for v in poly1 poly2; do
  TAE_B200_LIB=tae_b200/libtae_b200.$v.so run attn_tests_$v python -m pytest tests/test_gpu_kernels.py -m gpu -q -k attention --timeout 300 -p no:cacheprovider
  TAE_B200_LIB=tae_b200/libtae_b200.$v.so run probe_attn_$v python tools/gpu_probe.py --attn-only; grep "attention fwd" gpurun_out/probe_attn_$v.log
done
start=$(date +%s)
timeout 1500 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench_full exit $? in $(( $(date +%s) - start )) s"
python tools/print_bench.py gpurun_out/bench_full.log 2>/dev/null || tail -c 3000 gpurun_out/bench_full.log
tail -n 5 gpurun_out/bench_full.err
