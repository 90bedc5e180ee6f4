#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary24.txt; tail -n 5 gpurun_out/$name.log; }
run gpu_all python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider
bash tools/gpu_bench.sh
