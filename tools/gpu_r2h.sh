#!/bin/bash
# Round-2 call H: what the round-end driver runs on one GPU — parity suite, smoke, the default bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; echo "gpu_tests exit $?: $(tail -n 1 gpurun_out/gpu_tests.log)"
grep -E "^FAILED|^ERROR" gpurun_out/gpu_tests.log | head
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?: $(tail -n 1 gpurun_out/smoke.log)"
start=$(date +%s)
timeout 1500 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench_full exit $? in $(( $(date +%s) - start )) s"
python tools/print_bench.py gpurun_out/bench_full.log 2>/dev/null || tail -c 3000 gpurun_out/bench_full.log
tail -n 3 gpurun_out/bench_full.err
start=$(date +%s)
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.log 2>&1; echo "reference arm exit $? in $(( $(date +%s) - start )) s"; tail -c 700 gpurun_out/bench_ref.log
