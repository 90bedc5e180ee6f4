#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary19.txt; tail -4 gpurun_out/$name.log; }
run k_attn python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -p no:cacheprovider -x -k attention
python tools/gpu_probe.py --attn-only
python tools/attn_trace.py
