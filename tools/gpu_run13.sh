#!/bin/bash
mkdir -p gpurun_out
python tools/ncu_gemm.py > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 6 -c 6 -o gpurun_out/gemm_r1b python tools/ncu_gemm.py > gpurun_out/ncu_gemm.log 2>&1
echo "ncu exit $?"
