#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary25.txt; tail -n 4 gpurun_out/$name.log; }
run k_ln python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "layernorm"
python tools/gpu_probe.py 2>&1 | grep -E "layernorm|colsum|mse|im2col|adamw"
