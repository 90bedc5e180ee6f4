#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary21.txt; tail -n 25 gpurun_out/$name.log; }
run k_fp32 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -p no:cacheprovider -k "fp32 or gemm or layernorm"
run m_fp32 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 900 -p no:cacheprovider -k "fp32 or golden"
