#!/bin/bash
# Round-2 call L: final single-GPU validation (parity suite, smoke, default bench line) and the GEMM DRAM-traffic capture on
# the final sources.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; echo "gpu_tests exit $?: $(tail -n 1 gpurun_out/gpu_tests.log)"
grep -E "^FAILED|^ERROR" gpurun_out/gpu_tests.log | head
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?: $(tail -n 1 gpurun_out/smoke.log)"
start=$(date +%s)
timeout 600 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench_full exit $? in $(( $(date +%s) - start )) s"
python tools/print_bench.py gpurun_out/bench_full.log 2>/dev/null || tail -c 3000 gpurun_out/bench_full.log
timeout 120 python tools/ncu_gemm.py > gpurun_out/ncu_plain_gemm.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm -s 12 -c 12 -f -o gpurun_out/gemm_r2 python tools/ncu_gemm.py > gpurun_out/ncu_gemm.log 2>&1
echo "gemm ncu rc $?"
