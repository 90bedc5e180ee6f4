#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary17.txt; tail -4 gpurun_out/$name.log; }
run gpu_all python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider
run probe_pipe python tools/gpu_probe.py
TAE_ATTN_BWD_V1=1 run probe_v1 python tools/gpu_probe.py --attn-only
run smoke python __graft_entry__.py smoke
