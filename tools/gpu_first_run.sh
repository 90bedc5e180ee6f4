#!/bin/bash
# First-contact GPU run: each group in its own process so one faulting kernel cannot hide the others.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary.txt; tail -5 gpurun_out/$name.log; }
run k_misc python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -k "not gemm" -p no:cacheprovider
run k_gemm_kk python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -k "gemm_majors and False-False" -p no:cacheprovider
run k_gemm_kmn python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -k "gemm_majors and False-True" -p no:cacheprovider
run k_gemm_mnmn python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -k "gemm_majors and True-True" -p no:cacheprovider
run k_gemm_mnk python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -k "gemm_majors and True-False" -p no:cacheprovider
run k_gemm_epi python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -k "gemm and not majors" -p no:cacheprovider
run model python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider
run probe python tools/gpu_probe.py --quick
cat gpurun_out/summary.txt
