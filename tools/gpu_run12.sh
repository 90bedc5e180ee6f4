#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary12.txt; tail -6 gpurun_out/$name.log; }
run k_gemm python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 200 -k "gemm" -p no:cacheprovider
run model python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 300 -p no:cacheprovider
run probe_full python tools/gpu_probe.py
grep -E "gemm|cuBLAS" gpurun_out/probe_full.log
