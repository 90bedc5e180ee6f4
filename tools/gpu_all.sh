#!/bin/bash
# Everything the round-end driver runs on a fresh B200, in one call:  gpurun --timeout 1800 -- 'bash tools/gpu_all.sh'
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; tail -n 4 gpurun_out/$name.log; }
run gpu_tests python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider
run smoke python __graft_entry__.py smoke
bash tools/gpu_bench.sh "$@"
