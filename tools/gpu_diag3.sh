#!/bin/bash
# last-minutes check of one variant: new tests on the default library, full suite on the variant, probe + bench A/B
v=$1; only=${2:-fc2}
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "dgelu" --timeout 120 -p no:cacheprovider > gpurun_out/gpu_tests_default_k.log 2>&1
echo "default dgelu tests exit $?: $(tail -n 1 gpurun_out/gpu_tests_default_k.log)"
TAE_B200_LIB=tae_b200/libtae_b200.$v.so timeout 300 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/gpu_tests_$v.log 2>&1
echo "$v gpu_tests exit $?: $(tail -n 1 gpurun_out/gpu_tests_$v.log)"
timeout 100 python tools/gpu_probe.py --gemm-only --only $only > gpurun_out/probe_default.log 2>&1
TAE_B200_LIB=tae_b200/libtae_b200.$v.so timeout 100 python tools/gpu_probe.py --gemm-only --only $only > gpurun_out/probe_$v.log 2>&1
paste -d'\n' <(grep -h "^gemm" gpurun_out/probe_default.log | sed 's/^/default /') <(grep -h "^gemm" gpurun_out/probe_$v.log | sed "s/^/$v /")
TAE_B200_LIB=tae_b200/libtae_b200.$v.so timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab_${v}_1.log 2>&1
python tools/print_bench.py gpurun_out/ab_${v}_1.log 2>/dev/null | head -8
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab_default_1.log 2>&1
python tools/print_bench.py gpurun_out/ab_default_1.log 2>/dev/null | head -8
