#!/bin/bash
# Round-2 call F: full suite on HEAD (incl. the mid-size reference fixture), then the evidence pass (tools/gpu_profile.sh r2).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; echo "gpu_tests exit $?: $(tail -n 1 gpurun_out/gpu_tests.log)"
grep -E "^FAILED|^ERROR" gpurun_out/gpu_tests.log | head
bash tools/gpu_profile.sh r2
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r2.csv
