#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary27.txt; tail -n 12 gpurun_out/$name.log; }
run k_new python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -p no:cacheprovider -k "rowdot or delta or attention or gemm"
run m_all python -m pytest tests/test_gpu_model.py tests/test_gpu_vit.py -m gpu -q --timeout 900 -p no:cacheprovider
bash tools/gpu_bench.sh --no-encode
