#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary20.txt; tail -n 4 gpurun_out/$name.log; }
run k_kernels python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -p no:cacheprovider -x
run probe_full python tools/gpu_probe.py
cat gpurun_out/probe_full.log
STEP="python tools/ncu_step.py"
$STEP > gpurun_out/ncu_plain_step.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1.csv $STEP > gpurun_out/ncu_step.log 2>&1
echo "launch list rc $?"; tail -n 2 gpurun_out/ncu_step.log; wc -l gpurun_out/launches_r1.csv
