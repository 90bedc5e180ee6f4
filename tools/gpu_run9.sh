#!/bin/bash
mkdir -p gpurun_out
python tools/ncu_attn.py > gpurun_out/ncu_attn_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 2 -c 2 -o gpurun_out/attn_r1 python tools/ncu_attn.py > gpurun_out/ncu_attn.log 2>&1
echo "ncu exit $?"
