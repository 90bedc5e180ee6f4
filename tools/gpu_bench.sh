#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench.log 2>&1; echo "bench exit $?"
python tools/print_bench.py gpurun_out/bench.log || tail -n 20 gpurun_out/bench.log
