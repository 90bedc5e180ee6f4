#!/bin/bash
mkdir -p gpurun_out
python tools/ncu_gemm.py > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 6 -c 6 -o gpurun_out/gemm_r1 python tools/ncu_gemm.py > gpurun_out/ncu_gemm.log 2>&1
echo "ncu exit $?"
timeout 600 python tools/gpu_probe.py --quick > gpurun_out/probe_q.log 2>&1; head -20 gpurun_out/probe_q.log
