#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary2.txt; tail -4 gpurun_out/$name.log; }
run smoke python __graft_entry__.py smoke
run k_ln python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -k "layernorm" -p no:cacheprovider
run bench python bench.py --steps 5 --warmup 3
run probe_full python tools/gpu_probe.py
