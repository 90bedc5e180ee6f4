#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary6.txt; tail -3 gpurun_out/$name.log; }
run k_ln python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -k layernorm -p no:cacheprovider
run bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline
run probe_full python tools/gpu_probe.py
grep -E "layernorm|attention" gpurun_out/probe_full.log
