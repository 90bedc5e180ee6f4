#!/bin/bash
# full parity suite + GEMM probe + one bench A/B for a variant build:  gpurun -- 'bash tools/gpu_diag2.sh tmabf'
v=$1
mkdir -p gpurun_out
timeout 200 python tools/gpu_probe.py --gemm-only > gpurun_out/probe_default.log 2>&1
TAE_B200_LIB=tae_b200/libtae_b200.$v.so timeout 600 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/gpu_tests_$v.log 2>&1
echo "$v gpu_tests exit $?: $(tail -n 1 gpurun_out/gpu_tests_$v.log)"
TAE_B200_LIB=tae_b200/libtae_b200.$v.so timeout 200 python tools/gpu_probe.py --gemm-only > gpurun_out/probe_$v.log 2>&1
paste -d'\n' <(grep -h "^gemm" gpurun_out/probe_default.log | sed 's/^/default /') <(grep -h "^gemm" gpurun_out/probe_$v.log | sed "s/^/$v /")
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ab_default_1.log 2>&1
TAE_B200_LIB=tae_b200/libtae_b200.$v.so timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ab_${v}_1.log 2>&1
python tools/print_bench.py gpurun_out/ab_default_1.log 2>/dev/null | head -6; python tools/print_bench.py gpurun_out/ab_${v}_1.log 2>/dev/null | head -6
