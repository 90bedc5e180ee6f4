#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary8.txt; tail -12 gpurun_out/$name.log; }
run k_attn python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 200 -k "attention" -p no:cacheprovider
run probe_full python tools/gpu_probe.py
grep -E "attention|layernorm" gpurun_out/probe_full.log
