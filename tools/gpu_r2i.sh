#!/bin/bash
# Round-2 call I (2 GPUs): whole GPU suite incl. the NCCL data-parallel parity tests, short-grid probe, 2-rank bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; echo "gpu_tests exit $?: $(tail -n 1 gpurun_out/gpu_tests.log)"
grep -E "^FAILED|^ERROR" gpurun_out/gpu_tests.log | head
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/gpu_probe.py --small-grids > gpurun_out/probe_small.log 2>&1; grep attention gpurun_out/probe_small.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --config 5 --steps 20 --warmup 5 > gpurun_out/enc_graph.log 2>&1; python tools/print_bench.py gpurun_out/enc_graph.log | head -1
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --config 5 --steps 20 --warmup 5 --no-graph > gpurun_out/enc_eager.log 2>&1; python tools/print_bench.py gpurun_out/enc_eager.log | head -1
bash tools/gpu_r2d_multi.sh test2
