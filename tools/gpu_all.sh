#!/bin/bash
# Everything the round-end driver runs on a fresh B200, in one call:  gpurun --timeout 1800 -- 'bash tools/gpu_all.sh'
# (parity suite, smoke, the default bench line with its secondary configs and reference arms, the CPU reference arm)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; echo "gpu_tests exit $?: $(tail -n 1 gpurun_out/gpu_tests.log)"
grep -E "^FAILED|^ERROR" gpurun_out/gpu_tests.log | head
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?: $(tail -n 1 gpurun_out/smoke.log)"
start=$(date +%s)
timeout 900 python bench.py "$@" > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench exit $? in $(( $(date +%s) - start )) s"
python tools/print_bench.py gpurun_out/bench_full.log 2>/dev/null || tail -c 3000 gpurun_out/bench_full.log
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.log 2>&1; echo "reference arm exit $?"; tail -c 400 gpurun_out/bench_ref.log
