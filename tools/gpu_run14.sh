#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary14.txt; tail -25 gpurun_out/$name.log; }
run model_full python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 900 -p no:cacheprovider -k "full_size or sharding or fused_adamw"
