#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary3.txt; tail -4 gpurun_out/$name.log; }
run k_all python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -p no:cacheprovider
run model python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider
run bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline
run probe_full python tools/gpu_probe.py
