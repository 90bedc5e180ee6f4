#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary18.txt; tail -4 gpurun_out/$name.log; }
run k_kernels python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -p no:cacheprovider -x
python tools/gpu_probe.py --attn-only
TAE_ATTN_BWD_V1=1 python tools/gpu_probe.py --attn-only
python tools/attn_trace.py
run probe_full python tools/gpu_probe.py
cat gpurun_out/probe_full.log
