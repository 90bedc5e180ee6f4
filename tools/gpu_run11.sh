#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary11.txt; tail -4 gpurun_out/$name.log; }
run k_sel python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 200 -k "attention or layernorm or wgrad" -p no:cacheprovider
run probe_full python tools/gpu_probe.py
grep -E "attention|layernorm|wgrad" gpurun_out/probe_full.log
