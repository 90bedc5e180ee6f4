#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary16.txt; tail -4 gpurun_out/$name.log; }
run k_gemm python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 200 -k "gemm" -p no:cacheprovider
run model python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider -k "not full_size"
bash tools/gpu_bench.sh
