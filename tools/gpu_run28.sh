#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary28.txt; tail -n 6 gpurun_out/$name.log; }
run k_new python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -p no:cacheprovider -k "rowdot or mse or im2col or patch"
python tools/gpu_probe.py 2>&1 | grep -E "mse|im2col"
bash tools/gpu_bench.sh --no-encode
