#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary26.txt; tail -n 6 gpurun_out/$name.log; }
run k_attn python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -p no:cacheprovider -x -k attention
timeout 120 python tools/gpu_probe.py --attn-only
timeout 120 python tools/attn_trace.py
